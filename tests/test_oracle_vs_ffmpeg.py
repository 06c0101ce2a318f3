"""Cross-check of the oracle against an INDEPENDENT decoder (CPU only).

The D reference cannot run in this image (no D compiler) and ships no MP3 golden vectors, so the oracle cannot be
pinned against the reference itself.  What the image does have is FFmpeg's libavcodec (inside
opencv_python_headless.libs), whose `mp3float` decoder shares no code with minimp3.  Two independent float
implementations of ISO 11172-3 / 13818-3 Layer III agree to a few 1e-7 of full scale when both are right, and
disagree at 1e-3 .. 1e-1 when either misreads a field -- so agreement here checks the oracle's Huffman books,
scalefactor decoding, requantisation, stereo processing, reorder, alias reduction, IMDCT windows, overlap and the
polyphase synthesis (and the synthetic generator's legality) on every format the generator can write.

Known, explained divergences that are excluded (both follow from the reference source, not from the restatement):
 * intensity stereo, MPEG-1, only the last band (sfb 21) intensity-coded: the reference gives it the default
   position (minimp3.d:974-980), FFmpeg re-uses scalefactor 20 -- those granules (and the two after, which hear
   them through the IMDCT overlap and the synthesis history) are skipped by rule;
 * intensity stereo in MPEG-2 LSF: FFmpeg supports positions < 16 only and has no "illegal position" marking, the
   reference implements both (minimp3.d:629-637, 981) -- not compared;
 * mixed_block_flag signalled on the short blocks only (`mixed_only_short=1`): the reference windows the tail of the
   START block by the flag of the block that FOLLOWS it (deferred windowing, minimp3.d:1062-1100), ISO decoders by the
   START block's own flag -- the generator's default signals the flag on start/short/stop alike, where both agree.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tools import ffmpeg_mp3 as ff  # noqa: E402

pytestmark = pytest.mark.skipif(not ff.available(), reason="libavcodec 62 (opencv_python_headless.libs) not present")

TOL = 5e-6   # of full scale; measured agreement is 0.7e-6 .. 2.2e-6 with |peak| ~ 0.65


def decode_both(p):
    import oracle
    from audio_formats_b200 import synth
    st = synth.generate(p, want_quantised=True)
    offs, sizes = ff.split_frames(st.data)
    assert len(offs) == st.frames
    got = ff.decode_frames(st.data, offs, sizes)
    ref, _ = oracle.decode_all(st.data)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return st, offs, got, ref


LONG = [dict(hz=hz, nch=nch, bitrate_kbps=rate) for hz, nch, rate in
        [(44100, 2, 128), (48000, 2, 192), (32000, 2, 96), (44100, 1, 64), (44100, 2, 320), (48000, 1, 160),
         (22050, 2, 64), (24000, 2, 96), (16000, 2, 48), (22050, 1, 32), (16000, 1, 160 // 2),
         (11025, 2, 32), (12000, 1, 16), (8000, 2, 24)]]


@pytest.mark.parametrize("fmt", LONG, ids=lambda f: f"{f['hz']}-{f['nch']}ch-{f['bitrate_kbps']}k")
def test_long_blocks_all_formats(built, fmt):
    """Every sample rate of MPEG-1 / MPEG-2 LSF / MPEG-2.5, mono and stereo, MS stereo, scfsi, heavy reservoir,
    every Huffman book (table_cycle), escapes, large scalefactors, CRC."""
    from audio_formats_b200 import synth
    for variant, kw in enumerate([dict(reservoir=1), dict(reservoir=2, scfsi=1, table_cycle=1, small_scalefactors=0, crc=1,
                                                          stereo_mode=1 if fmt["nch"] == 2 else 0)]):
        p = synth.SynthParams(seed=40 + variant, nframes=40, **fmt, **kw)
        _, _, got, ref = decode_both(p)
        d = np.abs(got.astype(np.float64) - ref).max()
        assert d <= TOL, f"{p}: max |oracle - ffmpeg| = {d:.3e}"


@pytest.mark.parametrize("hz,nch,rate", [(44100, 2, 128), (48000, 1, 96), (32000, 2, 160), (22050, 2, 64), (24000, 1, 48),
                                         (16000, 2, 80), (11025, 2, 32), (12000, 1, 24)])
def test_short_and_mixed_blocks(built, hz, nch, rate):
    """long -> start -> short / mixed -> stop sequences, block types independent per channel, subblock_gain,
    MS stereo; mixed_block_flag signalled consistently on the start / short / stop blocks of a mixed run."""
    from audio_formats_b200 import synth
    # LSF mixed blocks are a decoder-convention divergence, not compared: the reference (like libmad) ends Huffman
    # region 0 after region0_count + 1 = 8 entries of the MIXED band table = 48 coefficients (minimp3.d:551-578,
    # 778-786); FFmpeg (like mpg123) hard-codes 36 for every block_type 2 granule.  MPEG-1 mixed tables give 36 both ways.
    p = synth.SynthParams(seed=7 + hz, hz=hz, nch=nch, bitrate_kbps=rate, nframes=120, block_mode=1 if hz >= 32000 else 2,
                          stereo_mode=1 if nch == 2 else 0, reservoir=2, small_scalefactors=0, table_cycle=1)
    _, _, got, ref = decode_both(p)
    d = np.abs(got.astype(np.float64) - ref).max()
    assert d <= TOL, f"{p}: max |oracle - ffmpeg| = {d:.3e}"


@pytest.mark.parametrize("hz,nch,rate", [(44100, 2, 128), (48000, 1, 96), (22050, 2, 64), (32000, 2, 224), (11025, 1, 24)])
def test_vbr(built, hz, nch, rate):
    """Bitrate index and padding bit changing from frame to frame, heavy reservoir across frames of different size."""
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=11, hz=hz, nch=nch, bitrate_kbps=rate, nframes=120, vbr=1, block_mode=1 if hz >= 32000 else 2,
                          scfsi=1, small_scalefactors=0, stereo_mode=1 if nch == 2 else 0, reservoir=2)
    _, _, got, ref = decode_both(p)
    d = np.abs(got.astype(np.float64) - ref).max()
    assert d <= TOL, f"{p}: max |oracle - ffmpeg| = {d:.3e}"


@pytest.mark.parametrize("hz,rate", [(44100, 128), (48000, 160), (32000, 96)])
def test_intensity_stereo_mpeg1(built, hz, rate):
    """MS + intensity joint stereo on long blocks (pan table, illegal position 7, MS fallback).  Granules where only
    sfb 21 is intensity-coded are skipped (see the module docstring), together with the two granules that follow."""
    import oracle
    from audio_formats_b200 import synth
    sfb20, sfb21 = {44100: (342, 418), 48000: (330, 384), 32000: (448, 550)}[hz]   # ISO 11172-3 table B.8, long blocks
    p = synth.SynthParams(seed=300 + hz // 100, hz=hz, nch=2, bitrate_kbps=rate, nframes=150, stereo_mode=2, reservoir=1,
                          small_scalefactors=0)
    st, offs, got, ref = decode_both(p)
    _, taps = oracle.decode_all(st.data, taps=st.granules)
    diff = np.abs(got.astype(np.float64) - ref).reshape(st.granules, 576, 2).max(axis=(1, 2))
    skip = np.zeros(st.granules + 2, bool)
    n_intensity = 0
    for g in range(st.granules):
        intensity = bool(st.data[offs[g // 2] + 3] & 0x10)
        n_intensity += intensity
        nz = np.nonzero(taps["is"][g][1])[0]
        last = int(nz.max()) if len(nz) else -1
        if intensity and sfb20 <= last < sfb21:
            skip[g:g + 3] = True   # g itself, g+1 through the IMDCT overlap, g+2 through the synthesis history
    skip = skip[:-2]
    assert n_intensity >= 40 and skip.mean() < 0.3, (n_intensity, skip.mean())
    bad = np.nonzero((diff > TOL) & ~skip)[0]
    assert len(bad) == 0, f"{p}: granules {bad[:8]} differ from ffmpeg by up to {diff[bad].max():.3e}"
