#!/usr/bin/env python3
"""A small but varied decode for compute-sanitizer (memcheck / racecheck), checked against the oracle:
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import audio_formats_b200 as af
import oracle
from audio_formats_b200 import synth

S = synth.SynthParams
cases = [S(seed=1, nframes=20), S(seed=2, nch=1, bitrate_kbps=64, nframes=20, mode_ext_any=1),
         S(seed=3, hz=22050, bitrate_kbps=64, nframes=24, block_mode=1, stereo_mode=2, small_scalefactors=0),
         S(seed=4, hz=16000, nch=1, bitrate_kbps=32, nframes=24, block_mode=2),
         S(seed=5, nframes=24, block_mode=1, stereo_mode=2, istereo_untied=1, scfsi=1, private_bits=1, reservoir=2),
         S(seed=6, bitrate_kbps=200, nframes=20, free_format=1, block_mode=1),
         S(seed=7, bitrate_kbps=160, nframes=30, vbr=1, block_mode=1, stereo_mode=1, table_cycle=1, small_scalefactors=0),
         S(seed=8, hz=8000, bitrate_kbps=24, nframes=24, stereo_mode=1), S(seed=9, hz=48000, bitrate_kbps=320, nframes=16, level=12.0)]
streams = [synth.generate(p) for p in cases]
ctx = af.Context(0)
outs = ctx.decode([s.data for s in streams])
for p, s, o in zip(cases, streams, outs):
    ref, _ = oracle.decode_all(s.data)
    assert o.shape == ref.shape and np.array_equal(o.view(np.uint32), ref.view(np.uint32)), p
# round 2: float taps, 16-bit delivery, tolerance mode, Layer I / II, the device prepass and the library's wave pipeline
from audio_formats_b200 import api
scans = [af.Scan(s.data) for s in streams[:5]]
af.decode_batch_with_taps(ctx, scans[:3], float_taps=True)
for flags in (api.OUT_S16, api.MATH_FUSED, api.OUT_S16 | api.MATH_FUSED):
    api.decode_mode(ctx, scans, flags)
l12 = [synth.generate_l12(synth.L12Params(seed=3, layer=2, nframes=30, joint=1)),
       synth.generate_l12(synth.L12Params(seed=4, layer=1, hz=32000, nch=1, bitrate_kbps=128, nframes=40)),
       synth.generate_l12(synth.L12Params(seed=5, layer=2, hz=22050, nch=1, bitrate_kbps=48, nframes=30, ref_syntax=0))]
for d, o in zip(l12, ctx.decode(l12)):
    ref, _ = oracle.decode_all(d)
    assert np.array_equal(o.view(np.uint32), ref.view(np.uint32))
raw_out, raw_info = ctx.decode_raw([s.data for s in streams] + l12[:1])
assert raw_info["device_streams"] >= 6, raw_info
for s, o in zip(streams, raw_out):
    ref, _ = oracle.decode_all(s.data)
    assert np.array_equal(o.view(np.uint32), ref.view(np.uint32))
total = sum(o.size for o in outs)
pin = api.PinnedBuffer(2 * (total + 256))
pipe = af.BatchPipeline(device=[0, 0], lanes=2, wave_streams=2, prepass_threads=2, s16=True)
pipe.decode_into([s.data for s in streams], pin.view(np.int16))
pipe.close()
pin.free()
st = af.AudioStream(ctx).openFromMemory(streams[4].data)
st.seekPosition(3000)
assert len(st.readSamplesFloat(5000)) == 5000
st.close()
ctx.close()
print("sanitize probe ok:", len(cases), "streams")
