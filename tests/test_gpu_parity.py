"""GPU <-> oracle parity (-m gpu).  Integer intermediates bit-exact, PCM bit-exact (stronger than the
north star's 1e-5 FS / 99.99 % identical-after-int16 bar, which is also asserted explicitly)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_FS = 1e-5          # north_star: max |delta| <= 1e-5 of full scale
MIN_IDENTICAL = 0.9999  # north_star: >= 99.99 % identical samples after 16-bit quantisation


def q16(x):
    return np.clip(np.rint(x.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int32)


def check_stream(ctx, stream, label):
    import audio_formats_b200 as af
    import oracle

    data = stream.data if hasattr(stream, "data") else stream
    quantised = getattr(stream, "quantised", None)
    sc = af.Scan(data)
    (pcm,), is_, iscf, ist, ftaps, sdesc = af.decode_batch_with_taps(ctx, [sc], float_taps=True)
    ref, taps = oracle.decode_all(data, taps=sc.granules + 8)
    nch = sc.channels
    assert len(taps) == sc.granules, label
    # integer intermediates: bit-exact
    ref_is = taps["is"][:, :nch].reshape(-1, 576)
    assert np.array_equal(is_, ref_is), f"{label}: quantised spectra differ at {np.argwhere(is_ != ref_is)[:5]}"
    ref_iscf = taps["iscf"][:, :nch].reshape(-1, 40)
    assert np.array_equal(iscf, ref_iscf), f"{label}: scalefactors differ at {np.argwhere(iscf != ref_iscf)[:5]}"
    if quantised is not None:
        assert np.array_equal(is_.reshape(-1, nch, 576), quantised), f"{label}: spectra differ from the encoder's"
    # float stage snapshots: bit-exact (signed zeros included) on every granule whose PCM is delivered
    per = 576 * nch
    skip, count = int(sdesc[0]["pcm_skip"]), int(sdesc[0]["pcm_count"])
    g0, g1 = skip // per, -(-(skip + count) // per) if count else skip // per
    for name in ("xr", "st", "im", "dct"):
        want = np.ascontiguousarray(taps[name][g0:g1, :nch]).reshape(-1, 576)
        got = ftaps[name][g0 * nch:g1 * nch]
        if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
            bad = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
            r, c = bad[0]
            raise AssertionError(f"{label}: float tap '{name}' differs at {len(bad)} places, first granule-channel {g0 * nch + r} "
                                 f"index {c}: got {got[r, c]!r} want {want[r, c]!r}")
    # PCM
    assert pcm.shape == ref.shape, label
    delta = np.abs(pcm.astype(np.float64) - ref.astype(np.float64)).max() if pcm.size else 0.0
    assert delta <= TOL_FS, f"{label}: max |delta| = {delta:.3e}"
    ident = (q16(pcm) == q16(ref)).mean() if pcm.size else 1.0
    assert ident >= MIN_IDENTICAL, f"{label}: only {ident:.6f} identical after int16 quantisation"
    bitexact = np.array_equal(pcm.view(np.uint32), ref.view(np.uint32))
    assert bitexact, f"{label}: PCM within tolerance (max delta {delta:.3e}) but not bit-identical"
    return pcm


def with_info_tag(data: bytes, frame_bytes: int, side_info_bytes: int, n_audio_frames: int, delay: int, padding: int) -> bytes:
    """Put a LAME-style Info tag frame (minimp3_ex.d:144-190) in front of a CBR stream: the first frame's header, zeroed
    side info and payload, then the tag with the encoder delay / padding fields."""
    b = bytearray(data[:4]) + bytearray(frame_bytes - 4)
    tag = bytearray(b"Info" + bytes([0, 0, 0, 1]) + n_audio_frames.to_bytes(4, "big"))
    tag += b"LAME3.100" + bytes(12)
    tag += bytes([(delay >> 4) & 0xFF, ((delay & 0xF) << 4) | ((padding >> 8) & 0xF), padding & 0xFF])
    b[4 + side_info_bytes:4 + side_info_bytes + len(tag)] = tag
    return bytes(b) + data


def test_config1_long_blocks(ctx):
    from audio_formats_b200 import synth
    check_stream(ctx, synth.generate(synth.config1_params(1), want_quantised=True), "config1")


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_config3_mixed_blocks_joint_stereo_reservoir(ctx, seed):
    from audio_formats_b200 import synth
    check_stream(ctx, synth.generate(synth.config3_params(seed, 6.0), want_quantised=True), f"config3[{seed}]")


@pytest.mark.parametrize("seed", list(range(24)))
def test_config4_heterogeneous(ctx, seed):
    from audio_formats_b200 import synth
    p = synth.config4_params(seed, 3.0)
    check_stream(ctx, synth.generate(p, want_quantised=True), f"config4[{seed}] {p.hz} Hz {p.nch} ch {p.bitrate_kbps} kbps")


@pytest.mark.parametrize("only_short", [0, 1])
@pytest.mark.parametrize("hz,rate", [(44100, 128), (22050, 64), (48000, 192)])
def test_mixed_block_flag_on_start_and_stop_blocks(ctx, hz, rate, only_short):
    """The reference derives n_long_bands from mixed_block_flag on EVERY block type (minimp3.d:1212): a STOP block
    carrying the flag keeps the normal window in its lowest bands.  only_short=1 is the inconsistent signalling
    (flag on the short blocks only) where the reference departs from ISO decoders; we follow the reference."""
    from audio_formats_b200 import synth
    for nch in (1, 2):
        p = synth.SynthParams.for_seconds(4.0, hz=hz, seed=900 + hz // 100 + nch, nch=nch, bitrate_kbps=rate if nch == 2 else rate // 2,
                                          block_mode=1, stereo_mode=1 if nch == 2 else 0, small_scalefactors=0,
                                          mixed_only_short=only_short)
        check_stream(ctx, synth.generate(p, want_quantised=True), f"mixed flag {hz} {nch}ch only_short={only_short}")


@pytest.mark.parametrize("hz,nch,kbps,nopad", [(44100, 2, 150, 1), (48000, 1, 100, 0), (22050, 2, 70, 0), (32000, 2, 500, 1)])
def test_free_format(ctx, hz, nch, kbps, nopad):
    """Free-format streams (bitrate index 0, frame size found by header search, up to 2,304-byte frames)."""
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=70 + nch, hz=hz, nch=nch, bitrate_kbps=kbps, nframes=60, free_format=1, no_padding=nopad,
                          block_mode=1 if hz >= 32000 else 2, stereo_mode=2 if nch == 2 else 0, reservoir=2, scfsi=1,
                          small_scalefactors=0)
    check_stream(ctx, synth.generate(p, want_quantised=True), f"free format {hz} {nch}ch {kbps} kbps")


@pytest.mark.parametrize("hz,nch,rate", [(44100, 2, 128), (48000, 1, 96), (32000, 2, 160), (22050, 2, 64), (16000, 1, 32)])
def test_private_bits_leak_into_scfsi(ctx, hz, nch, rate):
    """The reference reads the private bits together with scfsi, so in MPEG-1 they act as granule 0's scfsi nibble
    (minimp3.d:530-540, 600-601): a flagged partition is copied from the frame's zeroed scratch and its bits are not
    read.  Files with private bits set decode that way in the reference, so they must here (MPEG-2 drops the bits)."""
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=500 + hz // 100 + nch, hz=hz, nch=nch, bitrate_kbps=rate, nframes=120, private_bits=1, scfsi=1,
                          block_mode=1 if hz >= 32000 else 2, stereo_mode=2 if nch == 2 else 0, reservoir=2,
                          small_scalefactors=0)
    check_stream(ctx, synth.generate(p, want_quantised=True), f"private bits {hz} {nch}ch")


@pytest.mark.parametrize("hz,nch,rate", [(44100, 2, 128), (48000, 1, 96), (22050, 2, 64), (32000, 2, 224), (11025, 1, 24)])
def test_vbr(ctx, hz, nch, rate):
    """Bitrate index and padding bit change from frame to frame (allowed by hdr_compare, minimp3.d:241-247)."""
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=600 + hz // 100 + nch, hz=hz, nch=nch, bitrate_kbps=rate, nframes=160, vbr=1, scfsi=1,
                          block_mode=1 if hz >= 32000 else 2, stereo_mode=2 if nch == 2 else 0, reservoir=2,
                          small_scalefactors=0)
    check_stream(ctx, synth.generate(p, want_quantised=True), f"vbr {hz} {nch}ch")


@pytest.mark.parametrize("hz,nch,rate,sm", [(44100, 1, 64, 0), (22050, 1, 32, 0), (48000, 1, 96, 0), (44100, 2, 128, 2), (22050, 2, 64, 2)])
def test_mode_extension_bits_outside_joint_stereo(ctx, hz, nch, rate, sm):
    """mode_extension is "don't care" outside joint stereo, but the reference tests its bits in every mode
    (minimp3.d:100-103): an MPEG-1 mono frame with the intensity bit goes silent, an MPEG-2 one is scaled by sqrt(2)
    when the MS bit is set as well, plain-stereo frames get intensity processing.  Same here."""
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=700 + hz // 100 + nch, hz=hz, nch=nch, bitrate_kbps=rate, nframes=120, mode_ext_any=1,
                          stereo_mode=sm, block_mode=1 if hz >= 32000 else 2, small_scalefactors=0, reservoir=1)
    check_stream(ctx, synth.generate(p, want_quantised=True), f"mode_ext {hz} {nch}ch")


@pytest.mark.parametrize("hz,rate,seed", [(44100, 128, 1), (48000, 160, 2), (32000, 96, 3), (22050, 64, 4), (44100, 192, 5)])
def test_intensity_stereo_with_untied_block_types(ctx, hz, rate, seed):
    """Intensity stereo walks channel 0's band layout over channel 1's ist_pos array.  When the channels use different
    block types it reads entries channel 1 never transmitted: zeros in granule 0, granule 0's leftovers (and the top-band
    entries granule 0's intensity pass wrote) in granule 1 -- the array is per-frame scratch in the reference."""
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=800 + seed, hz=hz, nch=2, bitrate_kbps=rate, nframes=200, stereo_mode=2, istereo_untied=1,
                          block_mode=1, small_scalefactors=0, reservoir=1, scfsi=1)
    check_stream(ctx, synth.generate(p, want_quantised=True), f"untied intensity {hz}")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_untied_intensity_with_encoder_delay_across_tiles(ctx, seed):
    """An encoder delay makes every tile after the first start on an odd granule, so its recompute halo starts on the
    SECOND granule of a frame; under intensity stereo with different block types per channel that granule reads what
    granule 0 left in the per-frame ist_pos scratch (minimp3.d:1497, 974-980), so the halo has to reach back to granule 0."""
    from audio_formats_b200 import synth
    p = synth.SynthParams(seed=820 + seed, hz=44100, nch=2, bitrate_kbps=128, nframes=200, stereo_mode=2, istereo_untied=1,
                          block_mode=1, small_scalefactors=0, reservoir=1, scfsi=1, no_padding=1)
    st = synth.generate(p)
    data = with_info_tag(st.data, 417, 32, 200, delay=576, padding=600)
    import audio_formats_b200 as af
    assert (af.Scan(data).stream_desc().pcm_skip // 1152) % 2 == 1      # tiles start on odd granules
    check_stream(ctx, data, f"untied intensity + delay [{seed}]")


def test_s16_output_is_the_quantised_float_output(ctx):
    """L3B_OUT_S16: the device delivers q = clamp(lrintf(x * 32768), -32768, 32767) of the float sample (bit-exact)."""
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import api, synth
    streams = [synth.generate(synth.config3_params(31, 4.0)), synth.generate(synth.config4_params(7, 3.0)),
               synth.generate(synth.config4_params(8, 3.0)), synth.generate(synth.config2_params(5, 5.0)),
               synth.generate(synth.SynthParams(seed=9, nframes=40, level=200.0, gain_base=215))]   # loud: exercises the clamp
    scans = [af.Scan(s.data) for s in streams]
    outs = api.decode_mode(ctx, scans, api.OUT_S16)
    clipped = 0
    for st, o in zip(streams, outs):
        ref, _ = oracle.decode_all(st.data)
        assert o.dtype == np.int16 and o.shape == ref.shape
        want = q16(ref)
        clipped += int((np.abs(ref) > 1.0).sum())
        assert np.array_equal(o.astype(np.int32), want), np.argwhere(o.astype(np.int32) != want)[:4]
    assert clipped > 0, "the loud stream was meant to exceed full scale"


FUSED_MIN_IDENTICAL = 0.995   # measured 0.9990-0.9996: below the north star's 99.99 %, hence not the default mode


@pytest.mark.parametrize("cfg", ["config2", "config3", "config4m", "config4s", "config5"])
def test_fused_math_mode_is_within_tolerance(ctx, cfg):
    """L3B_MATH_FUSED (multiply-adds contracted into FMAs): PCM within 1e-5 of full scale of the reference; the share of
    samples identical after 16-bit quantisation is reported and bounded from below."""
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import api, synth
    p = {"config2": synth.config2_params(3, 8.0), "config3": synth.config3_params(3, 8.0), "config4m": synth.config4_params(2, 6.0),
         "config4s": synth.config4_params(45, 6.0), "config5": synth.config5_params(3, 6.0)}[cfg]
    st = synth.generate(p)
    (got,) = api.decode_mode(ctx, [af.Scan(st.data)], api.MATH_FUSED)
    ref, _ = oracle.decode_all(st.data)
    assert got.shape == ref.shape
    delta = np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()
    ident = (q16(got) == q16(ref)).mean()
    print(f"fused mode {cfg}: max |delta| = {delta:.3e} FS, identical after q16 = {ident:.6f}")
    assert delta <= TOL_FS, delta
    assert ident >= FUSED_MIN_IDENTICAL, ident


def test_documented_deviation_huffman_overrun_past_the_frame(ctx):
    """DESIGN.md 6, deviation (i), shown AS a deviation.  A (malformed) granule-channel whose big_values region runs past
    the end of its frame's main data: the reference's per-frame buffer is zero there (D default initialisation of the
    scratch), the linear blob of the GPU path holds the next frame's bytes.  Both decode the stream without complaint and
    agree everywhere except around the damaged granule -- and there they do differ, which is what the document says."""
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import synth
    st = synth.generate(synth.SynthParams(seed=77, nframes=60, reservoir=0, no_padding=1))     # 417-byte frames, main_data_begin = 0
    b = bytearray(st.data)
    k = 30                                       # frame to damage: big_values of its LAST granule-channel -> 288 pairs
    bit0 = (417 * k + 4) * 8 + 20 + 3 * 59 + 12  # MPEG-1 stereo side info: 20 bits, then 59 per granule-channel; part2_3_length is 12 bits
    for i in range(9):
        byte, bit = (bit0 + i) >> 3, 7 - ((bit0 + i) & 7)
        want = (288 >> (8 - i)) & 1
        b[byte] = (b[byte] & ~(1 << bit)) | (want << bit)
    data = bytes(b)
    ref, _ = oracle.decode_all(data)
    (got,) = ctx.decode([data])
    assert got.shape == ref.shape == (60 * 1152, 2)
    g = 2 * k + 1                                # the damaged granule
    lo, hi = 576 * g, 576 * (g + 3)              # its own PCM, the next granule (IMDCT overlap) and the one after (synthesis history)
    same = got.view(np.uint32) == ref.view(np.uint32)
    assert same[:lo].all() and same[hi:].all()
    assert not same[lo:hi].all(), "the over-read no longer differs from the reference: update DESIGN.md section 6 (i)"


def test_config5_320kbps(ctx):
    from audio_formats_b200 import synth
    check_stream(ctx, synth.generate(synth.config5_params(5, 6.0), want_quantised=True), "config5")


@pytest.mark.parametrize("hz", [8000, 11025, 12000])
def test_mpeg25(ctx, hz):
    from audio_formats_b200 import synth
    for nch in (1, 2):
        p = synth.SynthParams.for_seconds(3.0, hz=hz, seed=hz + nch, nch=nch, bitrate_kbps=32 if nch == 1 else 64,
                                          block_mode=0, stereo_mode=2 if nch == 2 else 0, scfsi=0, small_scalefactors=0)
        check_stream(ctx, synth.generate(p, want_quantised=True), f"mpeg2.5 {hz} {nch}ch")


def test_crc_and_tags(ctx):
    from audio_formats_b200 import synth
    from dataclasses import replace
    p = replace(synth.config3_params(11, 3.0), crc=1, id3v2_bytes=1000, id3v1=1)
    check_stream(ctx, synth.generate(p, want_quantised=True), "crc+id3")


def test_batch_of_streams_matches_single(ctx):
    """Batch entry point: many heterogeneous streams in one launch == each decoded alone."""
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import synth
    streams = [synth.generate(synth.config4_params(s, 2.0)) for s in range(40, 72)]
    outs = ctx.decode([s.data for s in streams])
    for s, o in zip(streams, outs):
        ref, _ = oracle.decode_all(s.data)
        assert o.shape == ref.shape
        assert np.array_equal(o.view(np.uint32), ref.view(np.uint32))


def test_tile_boundaries_long_stream(ctx):
    """A stream much longer than one CTA tile: every tile recomputes its 2-granule halo correctly."""
    from audio_formats_b200 import synth
    check_stream(ctx, synth.generate(synth.config3_params(21, 30.0), want_quantised=True), "long")


def test_wave_pipeline_matches_oracle(ctx):
    """BatchPipeline (waves over several contexts, recycled workspaces, pinned staging) delivers the same PCM."""
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import api, synth
    streams = [synth.generate(synth.config4_params(s, 1.5)) for s in range(80, 120)]
    datas = [s.data for s in streams]
    refs = [oracle.decode_all(d)[0] for d in datas]
    total = sum(r.size for r in refs)
    pin = api.PinnedBuffer(4 * (total + 8 * len(datas) + 64))
    out = pin.view(np.float32)
    pipe = af.BatchPipeline(device=0, lanes=3, wave_streams=7, prepass_threads=4)
    for _ in range(2):     # second pass runs through the recycled workspaces
        out[:] = 0
        info = pipe.decode_into(datas, out)
        for (off, frames, ch, hz), ref, st in zip(info, refs, streams):
            assert (frames, ch, hz) == (ref.shape[0], ref.shape[1], st.params.hz)
            got = out[off:off + frames * ch].reshape(frames, ch)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    pipe.close()
    pin.free()


def test_pipeline_over_several_devices_and_bad_inputs(ctx):
    """The library's multi-GPU batch entry point (l3b_pipeline_create with a device list): streams are assigned by file,
    longest first; here the "devices" are GPU 0 twice, which runs the same code path on one GPU.  Inputs that are not
    MPEG audio get their status and no PCM without disturbing the others; 16-bit delivery."""
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import api, synth
    streams = [synth.generate(synth.config4_params(s, 1.0 + 0.25 * (s % 5))) for s in range(200, 230)]
    datas = [s.data for s in streams]
    datas[7] = b"this is not an MPEG audio stream" * 40
    datas[19] = datas[19][:9]
    refs = [oracle.decode_all(d)[0] if i not in (7, 19) else None for i, d in enumerate(datas)]
    total = sum(r.size for r in refs if r is not None)
    pin = api.PinnedBuffer(2 * (total + 16 * len(datas) + 64))
    out = pin.view(np.int16)
    pipe = af.BatchPipeline(device=[0, 0], lanes=2, wave_streams=4, prepass_threads=3, s16=True)
    info = pipe.decode_into(datas, out)
    assert info[7] is None and info[19] is None and pipe.status[7] == api.E_USER and pipe.status[19] != 0
    loads = [0, 0]
    for i, (inf, ref) in enumerate(zip(info, refs)):
        if ref is None:
            continue
        off, frames, ch, hz = inf
        assert (frames, ch) == ref.shape
        assert np.array_equal(out[off:off + frames * ch].reshape(frames, ch).astype(np.int32), q16(ref)), i
        loads[pipe.device_of[i]] += len(datas[i])
    assert min(loads) > 0.8 * max(loads)          # longest-first assignment balances the two halves by bytes
    pipe.close()
    pin.free()


def test_large_batch_sampled_against_oracle_and_checksums(ctx):
    """A config-2 shaped batch too big to check stream by stream on the CPU in seconds: 192 x 60 s streams
    (1.76 M granule-channels).  A seeded sample is compared bit-exactly with the oracle; every stream is checked
    through a size-independent property: decoding it inside the batch == decoding it alone (checksum of checksums)."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    import audio_formats_b200 as af
    import oracle
    from audio_formats_b200 import synth
    n = 192
    with ThreadPoolExecutor(8) as ex:
        streams = list(ex.map(lambda s: synth.generate(synth.config2_params(s, 60.0)), range(5000, 5000 + n)))
        scans = list(ex.map(lambda st: af.Scan(st.data), streams))
    outs = ctx.decode_scans(scans)
    assert all(o.shape == (st.frames * 1152, 2) for o, st in zip(outs, streams))
    rng = np.random.default_rng(11)
    for i in rng.choice(n, 6, replace=False):
        ref = oracle.transcode_loop(streams[i].data, keep=True)[3]
        assert np.array_equal(outs[i].view(np.uint32), ref.view(np.uint32)), i
    batch_crc = [zlib.crc32(o.tobytes()) for o in outs]
    for i in rng.choice(n, 24, replace=False):      # alone == inside the batch (different tiles/CTAs/neighbours)
        (alone,) = ctx.decode_scans([scans[i]])
        assert zlib.crc32(alone.tobytes()) == batch_crc[i], i
    assert len(set(batch_crc)) == n                 # no two seeds collide: outputs are really per stream


@pytest.mark.parametrize("seed", range(48))
def test_random_generator_profiles(ctx, seed):
    """Randomised sweep over the generator's knobs (format, bitrate, block switching, stereo mode, reservoir
    pressure, scfsi, escapes, gain, level, CRC, tags): spectra and PCM stay bit-exact."""
    from audio_formats_b200 import synth
    rng = np.random.default_rng(1000 + seed)
    hz = int(rng.choice([8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000]))
    nch = int(rng.integers(1, 3))
    rates = [32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320] if hz >= 32000 else \
            [8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160]
    lo = 32 if hz >= 32000 else (16 if hz >= 16000 else 8)
    rate = int(rng.choice([r for r in rates if r >= lo * nch and not (nch == 1 and hz >= 32000 and r > 192)]))
    p = synth.SynthParams(seed=seed, hz=hz, nch=nch, bitrate_kbps=rate, nframes=int(rng.integers(12, 70)),
                          block_mode=int(rng.integers(0, 2)), stereo_mode=int(rng.integers(0, 3)) if nch == 2 else 0,
                          reservoir=int(rng.integers(0, 3)), scfsi=int(rng.integers(0, 2)), crc=int(rng.integers(0, 2)),
                          escapes=int(rng.integers(0, 2)), gain_base=int(rng.integers(150, 200)),
                          level=float(rng.choice([0.3, 1.0, 3.0, 8.0, 20.0])),
                          small_scalefactors=int(rng.integers(0, 2)), table_cycle=int(rng.integers(0, 2)),
                          table_cycle_pos=seed, no_padding=int(rng.integers(0, 2)),
                          id3v2_bytes=int(rng.choice([0, 0, 10, 777])), id3v1=int(rng.integers(0, 2)),
                          mixed_only_short=int(seed % 3 == 0), private_bits=int(seed % 4 == 1), vbr=int(seed % 5 == 2), mode_ext_any=int(seed % 7 == 3))
    check_stream(ctx, synth.generate(p, want_quantised=True), f"random[{seed}] {p}")
