"""The batch entry point with the prepass on the GPU (l3b_raw_*): frame walk, side-info parse, reservoir
recurrence and main-data gathering as kernels for well-formed streams, host prepass for everything else -- so the PCM must be
the host route's (and the oracle's) bit for bit on clean streams AND on the fault-injection / header-fuzz corpus."""
import sys
from dataclasses import replace
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_clean_streams_take_the_device_prepass(ctx):
    import oracle
    from audio_formats_b200 import synth
    params = [synth.config1_params(1), synth.config3_params(4, 5.0), synth.config5_params(2, 4.0)] + \
             [synth.config4_params(s, 2.0) for s in range(16)] + \
             [synth.SynthParams(seed=61, hz=44100, nframes=150, vbr=1, block_mode=1, stereo_mode=2, reservoir=2, scfsi=1),
              synth.SynthParams(seed=62, hz=48000, nch=1, bitrate_kbps=96, nframes=150, crc=1, id3v2_bytes=777, id3v1=1),
              synth.SynthParams(seed=63, hz=11025, nch=1, bitrate_kbps=24, nframes=120),
              synth.SynthParams(seed=64, hz=8000, nch=2, bitrate_kbps=64, nframes=120, stereo_mode=2, block_mode=1),
              synth.SynthParams(seed=65, nframes=200, private_bits=1, scfsi=1, reservoir=2),
              synth.SynthParams(seed=66, nframes=2500, reservoir=2, block_mode=1)]                     # longer than one tile chain
    datas = [synth.generate(p).data for p in params]
    outs, info = ctx.decode_raw(datas)
    assert info["device_streams"] == len(datas), info
    assert info["prepass_ms"] > 0
    for d, o in zip(datas, outs):
        ref, _ = oracle.decode_all(d)
        assert o.shape == ref.shape and np.array_equal(bits(o), bits(ref))


def test_everything_else_takes_the_host_prepass_with_the_same_result(ctx):
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import api, synth
    import test_gpu_parity as tp
    clean = synth.generate(synth.config2_params(9, 3.0)).data
    fixed = synth.generate(replace(synth.config2_params(10, 3.0), no_padding=1)).data      # every frame 417 bytes
    datas = [clean,
             synth.generate(synth.SynthParams(seed=71, hz=44100, nch=2, bitrate_kbps=150, nframes=60, free_format=1, no_padding=1)).data,
             tp.with_info_tag(synth.generate(replace(synth.config1_params(4), nframes=60, no_padding=1)).data, 417, 32, 60, 576, 1000),
             synth.generate_l12(synth.L12Params(seed=72, layer=2, nframes=40)),
             clean[:len(clean) - 200],                      # cut inside the last frame
             b"\\x00" * 333 + clean,                          # leading garbage
             b"definitely not an mp3" * 100,
             fixed[:417 * 5] + fixed[417 * 9:]]             # four frames missing: reservoir underruns, still a clean chain
    outs, info = ctx.decode_raw(datas)
    assert info["device_streams"] == 2, info               # the clean stream and the one with whole frames missing
    assert info["status"][6] == api.E_USER and outs[6] is None
    for i, d in enumerate(datas):
        if i == 6:
            continue
        ref, _ = oracle.decode_all(d)
        assert outs[i].shape == ref.shape and np.array_equal(bits(outs[i]), bits(ref)), i
    # 16-bit delivery through the same route
    outs16, _ = ctx.decode_raw(datas[:3], flags=api.OUT_S16)
    for d, o in zip(datas[:3], outs16):
        ref, _ = oracle.decode_all(d)
        assert np.array_equal(o, np.clip(np.rint(ref.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16))


def test_fuzz_corpus_matches_the_host_route(ctx):
    """tests/test_host_prepass.py's fault-injection and header-fuzz streams, all in ONE raw batch: whichever route a stream
    takes, its PCM equals the host-prepass route's."""
    import audio_formats_b200 as af
    import test_host_prepass as hp
    datas = [hp.header_fuzz_stream(seed)[0] for seed in range(24)]
    outs, info = ctx.decode_raw(datas)
    host = []
    for d in datas:
        try:
            host.append(ctx.decode_scans([af.Scan(d)])[0])
        except af.L3BError:
            host.append(None)
    n_dev = info["device_streams"]
    assert 0 < n_dev < len(datas), info       # the duplicated-frame streams stay clean chains, the rest is damaged
    for i, (o, h) in enumerate(zip(outs, host)):
        if h is None:
            assert o is None or o.size == 0, i
        else:
            assert o is not None and o.shape == h.shape and np.array_equal(bits(o), bits(h)), i
