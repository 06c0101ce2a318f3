"""The oracle, the generator and the golden fixtures agree (CPU only)."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = sorted((Path(__file__).resolve().parent / "golden").glob("*.npz"))


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_oracle_reproduces_golden_pcm_bit_exactly(built, path):
    import oracle
    g = np.load(path)
    data = g["mp3"].tobytes()
    pcm, taps = oracle.decode_all(data, taps=len(g["quantised"]) + 4)
    assert pcm.shape == g["pcm"].shape
    assert np.array_equal(pcm.view(np.uint32), g["pcm"].view(np.uint32))
    nch = g["quantised"].shape[1]
    assert np.array_equal(taps["is"][:, :nch], g["quantised"])
    assert np.array_equal(taps["iscf"][:, :nch], g["iscf"])


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_generator_is_deterministic(built, path):
    from audio_formats_b200 import synth
    g = np.load(path)
    p = synth.SynthParams(**json.loads(g["params"].tobytes().decode()))
    st = synth.generate(p, want_quantised=True)
    assert st.data == g["mp3"].tobytes()
    assert np.array_equal(st.quantised, g["quantised"])


def test_chunked_reads_equal_one_big_read(built):
    """minimp3_ex.d:800-813: leftovers of a frame are drained first, so chunking is invisible."""
    import oracle
    from audio_formats_b200 import synth
    st = synth.generate(synth.config3_params(5, 2.0))
    big, _ = oracle.decode_all(st.data, chunk_frames=1 << 20)
    for chunk in (1, 7, 1024, 1153, 5000):
        part, _ = oracle.decode_all(st.data, chunk_frames=chunk)
        assert np.array_equal(part.view(np.uint32), big.view(np.uint32)), chunk


def test_oracle_seek_assertions(built):
    """The debug(checkSeeking) list of examples/transcode/source/main.d:90-162, on the oracle."""
    import oracle
    from audio_formats_b200 import synth
    st = synth.generate(synth.config3_params(8, 4.0))
    s = oracle.OracleStream(st.data)
    n = s.length_frames
    assert s.tell() == 0
    assert s.seek(0) and s.tell() == 0
    assert not s.seek(n + 1) and s.tell() == 0
    assert not s.seek(-1) and s.tell() == 0
    assert s.seek(n // 2) and s.tell() == n // 2
    assert s.seek(n - 1) and s.tell() == n - 1
    assert len(s.read_float(2)) == 1
    assert s.seek(n) and len(s.read_float(2)) == 0
    assert s.seek(0)
    assert len(s.read_float(16)) == 16 and s.tell() == 16
    # seeking lands on the same samples as linear decode
    full, _ = oracle.decode_all(st.data)
    for pos in (1, 575, 576, 1151, 1152, 5000, n // 3, n - 2000):
        assert s.seek(pos)
        got = s.read_float(700)
        want = full[pos:pos + 700]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), pos
    s.close()
