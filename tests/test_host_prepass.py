"""Host prepass (frame sync, side info, reservoir slicing) against the oracle's control flow (CPU only):
which frames decode, how many samples are delivered, stream properties -- including damaged input."""
from dataclasses import replace

import numpy as np
import pytest


def scan_vs_oracle(data, label=""):
    import audio_formats_b200 as af
    import oracle
    try:
        ref = oracle.OracleStream(data)
    except ValueError:
        with pytest.raises(af.L3BError):
            af.Scan(data)
        return None
    sc = af.Scan(data)
    assert (sc.channels, sc.samplerate, sc.length_frames) == (ref.channels, ref.samplerate, ref.length_frames), label
    ref.close()
    pcm, taps = oracle.decode_all(data, taps=1 << 14)
    assert sc.delivered_samples == pcm.size, label
    return sc, pcm, taps


@pytest.mark.parametrize("cfg", ["c1", "c3", "c5"] + [f"c4_{i}" for i in range(12)])
def test_clean_streams(built, cfg):
    from audio_formats_b200 import synth
    p = {"c1": synth.config1_params(1), "c3": synth.config3_params(3, 4.0), "c5": synth.config5_params(5, 4.0)}.get(cfg)
    if p is None:
        p = synth.config4_params(int(cfg.split("_")[1]), 3.0)
    st = synth.generate(p)
    sc, pcm, taps = scan_vs_oracle(st.data, cfg)
    assert sc.granules == len(taps) == st.granules
    d = sc.descs.reshape(sc.granules, sc.channels)
    assert (d["w3"][0] >> 31).all() and not (d["w3"][1:] >> 31).any()   # state is zero only before the first granule
    assert (np.diff(d["bit_start"].reshape(-1).astype(np.int64)) >= 0).all()   # bit offsets grow monotonically


@pytest.mark.parametrize("hz,nch,kbps,nopad", [(44100, 2, 150, 1), (48000, 1, 100, 0), (22050, 2, 70, 0), (44100, 2, 400, 0),
                                                (32000, 2, 500, 1), (11025, 1, 20, 0)])
def test_free_format(built, hz, nch, kbps, nopad):
    """Bitrate index 0: the frame size comes from searching for the next matching header (minimp3.d:1460-1472,
    hdr_frame_bytes :270-278).  The index pass counts two frames fewer than the read loop decodes (it needs two
    following headers); the prepass has to reproduce both numbers."""
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=60 + nch, hz=hz, nch=nch, bitrate_kbps=kbps, nframes=40, free_format=1, no_padding=nopad,
                          block_mode=1 if hz >= 32000 else 2, stereo_mode=1 if nch == 2 else 0, reservoir=2)
    st = synth.generate(p)
    sc, pcm, taps = scan_vs_oracle(st.data, f"free format {hz} {nch} {kbps}")
    assert sc.granules == len(taps) == st.granules
    assert pcm.shape[0] == st.frames * st.samples_per_frame


@pytest.mark.parametrize("hz,nch,rate", [(44100, 2, 128), (48000, 1, 96), (22050, 2, 64), (11025, 1, 24)])
def test_vbr(built, hz, nch, rate):
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=90 + nch, hz=hz, nch=nch, bitrate_kbps=rate, nframes=80, vbr=1, reservoir=2,
                          block_mode=1 if hz >= 32000 else 2)
    st = synth.generate(p)
    sc, pcm, taps = scan_vs_oracle(st.data, f"vbr {hz} {nch}")
    assert sc.granules == len(taps) == st.granules and pcm.shape[0] == st.frames * st.samples_per_frame


@pytest.mark.parametrize("nch", [1, 2])
def test_private_bits(built, nch):
    """Private bits set (they become granule 0's scfsi nibble in the reference): same frames, same lengths."""
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=80 + nch, nch=nch, bitrate_kbps=128 // (3 - nch), nframes=50, private_bits=1, scfsi=1, block_mode=1)
    st = synth.generate(p)
    sc, pcm, taps = scan_vs_oracle(st.data, f"private bits {nch}ch")
    assert sc.granules == len(taps) == st.granules
    d = sc.descs.reshape(sc.granules, nch)
    assert ((d["w2"][0::2] >> 27) & 15).any()        # some granule 0 carries a leaked nibble


def test_tags_and_crc(built):
    from audio_formats_b200 import synth
    p = replace(synth.config3_params(12, 2.0), crc=1, id3v2_bytes=4096, id3v1=1)
    scan_vs_oracle(synth.generate(p).data, "tags")


def _ape_tag(item_bytes: int) -> bytes:
    """APEv2 tag: 32-byte header, items, 32-byte footer; the size field counts items + footer, and the reference removes
    32 + size bytes (minimp3_ex.d:102-108), i.e. exactly a tag that has its header."""
    size = item_bytes + 32
    def block(flags):
        return b"APETAGEX" + (2000).to_bytes(4, "little") + size.to_bytes(4, "little") + (1).to_bytes(4, "little") + \
               flags.to_bytes(4, "little") + bytes(8)
    return block(0xA0000000) + bytes((i * 37) & 0xFF for i in range(item_bytes)) + block(0x80000000)


@pytest.mark.parametrize("variant", ["ape", "ape+id3v1", "id3v1+ext", "id3v2-footer", "all"])
def test_trailing_and_leading_tags(built, variant):
    """ID3v1, its 227-byte "TAG+" extension, APEv2 footers and an ID3v2 tag with a footer are all skipped the way the
    reference skips them (minimp3_ex.d:93-142), including the order in which the trailing ones are tested."""
    from audio_formats_b200 import synth
    st = synth.generate(replace(synth.config3_params(31, 1.5), crc=0))
    body = st.data
    id3v1 = b"TAG" + bytes(125)
    ext = b"TAG+" + bytes(223)
    id3v2f = b"ID3\x04\x00\x10" + bytes([0, 0, 2, 0]) + bytes(256) + b"3DI\x04\x00\x10" + bytes([0, 0, 2, 0])
    data = {"ape": body + _ape_tag(300), "ape+id3v1": body + _ape_tag(77) + id3v1, "id3v1+ext": body + ext + id3v1,
            "id3v2-footer": id3v2f + body, "all": id3v2f + body + _ape_tag(500) + ext + id3v1}[variant]
    sc, pcm, taps = scan_vs_oracle(data, variant)
    assert sc.granules == st.granules == len(taps)          # every frame still decodes: nothing of a tag was taken for audio
    assert pcm.shape[0] == st.frames * st.samples_per_frame


def test_first_frames_without_reservoir_are_skipped(built):
    """Cutting the head off a stream leaves frames whose main_data_begin points before the cut: they emit no PCM
    and do not count in the length (minimp3.d:1546-1556, minimp3_ex.d:613-619)."""
    from audio_formats_b200 import synth
    from audio_formats_b200.api import Scan
    st = synth.generate(replace(synth.config1_params(3), reservoir=2, nframes=60))
    whole = Scan(st.data)
    # cut at the start of frame 5 (frames are 417/418 bytes at 44.1 kHz / 128 kbps)
    import oracle
    L = oracle.lib()
    pos, k = 0, 0
    b = st.data
    while k < 5:
        pos += L.l3o_hdr_frame_bytes(b[pos:pos + 4], 0) + L.l3o_hdr_padding(b[pos:pos + 4])
        k += 1
    cut = b[pos:]
    res = scan_vs_oracle(cut, "cut")
    assert res is not None
    sc, pcm, taps = res
    assert sc.granules < whole.granules - 2 * 5 + 1   # at least one more frame lost to the missing reservoir


@pytest.mark.parametrize("seed", range(10))
def test_fault_injection(built, seed):
    """Truncation, garbage, corrupted side info: the prepass must follow the oracle's resync/drop decisions."""
    from audio_formats_b200 import synth
    rng = np.random.default_rng(seed)
    st = synth.generate(replace(synth.config3_params(100 + seed, 1.5), nframes=50))
    b = bytearray(st.data)
    kind = seed % 5
    if kind == 0:      # truncate mid-frame
        b = b[: len(b) - int(rng.integers(1, 400))]
    elif kind == 1:    # garbage prefix
        b = bytearray(rng.integers(0, 255, 777, dtype=np.uint8).tobytes()) + b
    elif kind == 2:    # overwrite a run of bytes in the middle (kills headers and side info)
        at = len(b) // 2
        b[at:at + 900] = rng.integers(0, 256, 900, dtype=np.uint8).tobytes()
    elif kind == 3:    # flip bits in side infos of several frames
        for _ in range(6):
            at = int(rng.integers(0, len(b) - 40))
            b[at] ^= 1 << int(rng.integers(0, 8))
    else:              # drop a whole chunk (lost sync)
        at = len(b) // 3
        del b[at:at + 1000]
    scan_vs_oracle(bytes(b), f"fault{kind}")


def _frame_offsets(data: bytes):
    import oracle
    L = oracle.lib()
    pos, offs = 0, []
    while pos + 4 <= len(data) and data[pos] == 0xFF and (data[pos + 1] & 0xE0) == 0xE0:
        offs.append(pos)
        pos += L.l3o_hdr_frame_bytes(data[pos:pos + 4], 0) + L.l3o_hdr_padding(data[pos:pos + 4])
    return offs


def header_fuzz_stream(seed):
    """The damaged stream of test_header_level_fuzz (also decoded by tests/test_raw_gpu.py); returns (bytes, kind)."""
    from audio_formats_b200 import synth
    rng = np.random.default_rng(4000 + seed)
    p = synth.SynthParams(seed=300 + seed, nframes=int(rng.integers(330, 420)), bitrate_kbps=int(rng.choice([128, 192, 320])),
                          vbr=int(seed % 3 == 0), block_mode=1, stereo_mode=1, reservoir=int(rng.integers(0, 3)))
    st = synth.generate(p)
    b = bytearray(st.data)
    offs = _frame_offsets(st.data)
    assert len(offs) == st.frames
    kind = seed % 6
    if kind == 0:      # flip header bits of several frames
        for k in rng.choice(len(offs) - 2, 8, replace=False):
            b[offs[k] + int(rng.integers(1, 4))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 1:    # a few garbage bytes between frames, at several places (back to front: offsets stay valid)
        for k in sorted(rng.choice(len(offs) - 2, 5, replace=False), reverse=True):
            b[offs[k]:offs[k]] = rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8).tobytes()
    elif kind == 2:    # duplicate a frame and drop another
        k, j = sorted(int(x) for x in rng.choice(range(20, len(offs) - 20), 2, replace=False))
        frame = bytes(b[offs[k]:offs[k + 1]])
        del b[offs[j]:offs[j + 1]]
        b[offs[k]:offs[k]] = frame
    elif kind == 3:    # frames of another sample rate spliced in (the sync chain must refuse them)
        other = synth.generate(synth.SynthParams(seed=9, hz=32000, nframes=30, bitrate_kbps=96)).data
        k = int(rng.integers(50, len(offs) - 50))
        b[offs[k]:offs[k]] = other[: len(other) // 2]
    elif kind == 4:    # kill one header out of every ~40 so that chains end on a bad header again and again
        for k in range(17, len(offs) - 2, 41):
            b[offs[k]] = 0x00
    else:              # cut the stream inside the last 16 KiB in several ways (end-of-buffer rule of the chain)
        b = b[: offs[len(offs) - int(rng.integers(1, 12))] + int(rng.integers(0, 300))]
    return bytes(b), kind


@pytest.mark.parametrize("seed", range(24))
def test_header_level_fuzz(built, seed):
    """Damage aimed at the frame headers of streams long enough for the 128 KiB window to slide: flipped header bits
    (version, layer, bitrate, sample rate, padding, mode), a few garbage bytes between frames, a duplicated frame, frames
    of another format spliced in.  The index pass verifies its ten-header sync chain incrementally; whatever it decides
    has to be what the reference's frame-by-frame search decides (same length, same delivered samples, same granules)."""
    data, kind = header_fuzz_stream(seed)
    sc, pcm, taps = scan_vs_oracle(data, f"header fuzz {kind}/{seed}")
    assert sc.granules == len(taps)


def test_not_mp3(built):
    import audio_formats_b200 as af
    with pytest.raises(af.L3BError):
        af.Scan(b"RIFF" + bytes(5000))
    with pytest.raises(af.L3BError):
        af.Scan(b"")


def test_scans_assemble_matches_python_assembly(built):
    """l3b_scans_assemble (the library builds a wave's blob, descriptors and stream table) == the numpy assembly."""
    import audio_formats_b200 as af
    from audio_formats_b200 import api, synth
    streams = [synth.generate(synth.config4_params(s, 0.7)) for s in range(30, 41)]
    scans = [af.Scan(st.data) for st in streams]
    fast = api.HostBatch(scans)                                   # C path
    slow = api.HostBatch(scans + [], replicate=2)                 # numpy path (two copies of the same set)
    n = len(scans)
    assert fast.n_grch * 2 == slow.n_grch and fast.blob.size * 2 == slow.blob.size
    assert np.array_equal(fast.blob, slow.blob[:fast.blob.size])
    assert np.array_equal(fast.descs, slow.descs[:fast.n_grch])
    for name in fast.streams.dtype.names:
        assert np.array_equal(fast.streams[name], slow.streams[name][:n]), name
    assert fast.pcm_floats == int(slow.streams["pcm_off"][n - 1] + slow.streams["pcm_count"][n - 1])
    # every stream is followed by >= 16 zero bytes and starts 16-byte aligned
    for sd in fast.streams:
        off, nb = int(sd["maindata_off"]), int(sd["maindata_bytes"])
        assert off % 16 == 0 and not fast.blob[off + nb: off + nb + 16].any()
