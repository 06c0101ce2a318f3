"""Layer I / II (minimp3.d:284-484, 1557-1579).

CPU part: the oracle's Layer I / II restatement against an INDEPENDENT decoder (FFmpeg's mp1float / mp2float inside the image's
libavcodec) wherever the D reference and ISO 11172-3 read the same syntax -- all of Layer I, and Layer II frames in which every
band-channel entry is allocated.  The reference evaluates get_bits(2) for the scfsi field of EVERY entry (minimp3.d:417-421;
the upstream C and ISO only for allocated entries), so it reads other Layer II frames differently from any ISO decoder; the
drop-in reproduces the reference, and the generator can write either syntax.
GPU part (-m gpu): bit-exact PCM against the oracle through the batch entry point, the AudioStream surface and 16-bit delivery.
"""
import sys
from dataclasses import replace
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))

CASES = [  # layer, hz, nch, kbps, joint
    (2, 44100, 2, 192, 0), (2, 48000, 2, 128, 1), (2, 32000, 1, 64, 0), (2, 44100, 2, 64, 0), (2, 44100, 1, 48, 0),
    (2, 22050, 2, 64, 1), (2, 24000, 1, 32, 0), (2, 16000, 2, 160, 0), (2, 48000, 2, 384, 1), (2, 32000, 2, 96, 1),
    (1, 44100, 2, 256, 0), (1, 32000, 1, 128, 0), (1, 48000, 2, 384, 1), (1, 22050, 2, 128, 1), (1, 44100, 1, 64, 0),
]


def frames_of(d, layer, hz, rate):
    offs, sizes, i = [], [], 0
    while i + 4 <= len(d):
        fb = (384 if layer == 1 else 1152) * rate * 125 // hz
        if layer == 1:
            fb &= ~3
        fb += ((4 if layer == 1 else 1) if (d[i + 2] >> 1) & 1 else 0)
        offs.append(i); sizes.append(fb); i += fb
    return offs, sizes


@pytest.mark.parametrize("layer,hz,nch,rate,joint", [c for c in CASES if c[0] == 1] +
                         [(2, 44100, 2, 384, 0), (2, 48000, 2, 256, 0), (2, 32000, 2, 192, 0), (2, 22050, 2, 160, 0), (2, 16000, 2, 128, 0)])
def test_oracle_agrees_with_ffmpeg_where_the_reference_reads_iso_syntax(layer, hz, nch, rate, joint):
    import ffmpeg_mp3
    import oracle
    from audio_formats_b200 import synth
    if not ffmpeg_mp3.available():
        pytest.skip("libavcodec not found")
    d = synth.generate_l12(synth.L12Params(seed=layer * 100 + hz // 1000, layer=layer, hz=hz, nch=nch, bitrate_kbps=rate, nframes=40,
                                           joint=joint, ref_syntax=0, all_alloc=1 if layer == 2 else 0))
    pcm, _ = oracle.decode_all(d)
    ff = ffmpeg_mp3.decode_frames(d, *frames_of(d, layer, hz, rate), decoder=b"mp2float" if layer == 2 else b"mp1float")
    assert len(ff) == len(pcm) == 40 * (384 if layer == 1 else 1152)
    assert np.abs(pcm).max() > 0.05
    assert np.abs(ff[:, :nch] - pcm).max() <= 5e-6


def test_reference_reads_a_scfsi_field_for_every_layer2_entry():
    """The documented departure of the D reference from ISO syntax: an ISO-syntax frame with an unallocated entry is read
    differently from FFmpeg, the same frame written with a scfsi field for every entry decodes sanely."""
    import oracle
    from audio_formats_b200 import synth
    p = synth.L12Params(seed=7, layer=2, hz=44100, nch=2, bitrate_kbps=128, nframes=30)
    sane, _ = oracle.decode_all(synth.generate_l12(replace(p, ref_syntax=1)))
    iso, _ = oracle.decode_all(synth.generate_l12(replace(p, ref_syntax=0)))
    assert 0.05 < np.abs(sane).max() < 1.0
    assert np.abs(iso).max() > 1.0        # mis-read scalefactors and samples: far above full scale


def test_host_prepass_counts_what_the_oracle_decodes():
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import synth
    for layer, hz, nch, rate, joint in CASES:
        d = synth.generate_l12(synth.L12Params(seed=11, layer=layer, hz=hz, nch=nch, bitrate_kbps=rate, nframes=30, joint=joint))
        ref, _ = oracle.decode_all(d)
        sc = af.Scan(d)
        assert (sc.channels, sc.samplerate) == (nch, hz)
        assert sc.delivered_samples == ref.size and sc.granules == 30 * (1 if layer == 1 else 3)
        assert sc.stream_desc().layer == layer


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("layer,hz,nch,rate,joint", CASES)
@pytest.mark.parametrize("ref_syntax", [1, 0])
def test_gpu_pcm_is_bit_identical(ctx, layer, hz, nch, rate, joint, ref_syntax):
    import oracle
    from audio_formats_b200 import synth
    d = synth.generate_l12(synth.L12Params(seed=21 + layer, layer=layer, hz=hz, nch=nch, bitrate_kbps=rate, nframes=90, joint=joint,
                                           ref_syntax=ref_syntax, crc=int(hz == 48000)))
    ref, _ = oracle.decode_all(d)
    (got,) = ctx.decode([d])
    assert got.shape == ref.shape
    assert np.array_equal(bits(got), bits(ref)), np.argwhere(bits(got) != bits(ref))[:4]


@pytest.mark.gpu
def test_gpu_mixed_batch_s16_and_tiles(ctx):
    """Layer I, II and III streams in one batch; a Layer II stream longer than one 64-granule tile; 16-bit delivery."""
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import api, synth
    datas = [synth.generate_l12(synth.L12Params(seed=31, layer=2, hz=44100, nch=2, bitrate_kbps=192, nframes=400)),
             synth.generate(synth.config3_params(5, 2.0)).data,
             synth.generate_l12(synth.L12Params(seed=32, layer=1, hz=32000, nch=1, bitrate_kbps=128, nframes=300)),
             synth.generate(synth.config4_params(3, 2.0)).data,
             synth.generate_l12(synth.L12Params(seed=33, layer=2, hz=24000, nch=1, bitrate_kbps=48, nframes=200, joint=0))]
    refs = [oracle.decode_all(d)[0] for d in datas]
    outs = ctx.decode(datas)
    for o, r in zip(outs, refs):
        assert o.shape == r.shape and np.array_equal(bits(o), bits(r))
    outs16 = api.decode_mode(ctx, [af.Scan(d) for d in datas], api.OUT_S16)
    for o, r in zip(outs16, refs):
        want = np.clip(np.rint(r.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
        assert np.array_equal(o, want)


@pytest.mark.gpu
def test_gpu_audiostream_seek_and_damage(ctx):
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import synth
    d = synth.generate_l12(synth.L12Params(seed=41, layer=2, hz=44100, nch=2, bitrate_kbps=160, nframes=120, joint=1))
    s = af.AudioStream(ctx).openFromMemory(d)
    o = oracle.OracleStream(d)
    assert (s.getNumChannels(), s.getLengthInFrames()) == (2, o.length_frames)
    for pos in (0, 1, 383, 384, 5000, 100000, o.length_frames - 7):
        assert s.seekPosition(pos) and o.seek(pos)
        a, b = s.readSamplesFloat(2000), o.read_float(2000)
        assert a.shape == b.shape and np.array_equal(bits(a), bits(b)), pos
    s.close(); o.close()
    # damage: a frame whose bits run past its end is dropped together with the decoder state (minimp3.d:1571-1575);
    # garbage in the middle forces a resync
    b = bytearray(d)
    rng = np.random.default_rng(3)
    b[30000:30700] = rng.integers(0, 256, 700, dtype=np.uint8).tobytes()
    del b[50000:50411]
    for blob in (bytes(b), d[:len(d) - 123]):
        ref, _ = oracle.decode_all(blob)
        (got,) = ctx.decode([blob])
        assert got.shape == ref.shape and np.array_equal(bits(got), bits(ref))
