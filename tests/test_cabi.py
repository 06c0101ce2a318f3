"""The C-ABI library loads and exports every symbol include/l3b200.h declares (no compute without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    h = (ROOT / "include" / "l3b200.h").read_text()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(l3b_[a-z0-9_]+)\s*\(", h)))


def test_every_declared_symbol_is_exported(built):
    import audio_formats_b200 as af
    lib = ctypes.CDLL(str(af.library_path()))
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header(built):
    import audio_formats_b200 as af
    L = af.load_library()
    assert sorted(L._l3b_signatures) == declared_symbols()


def test_no_cpu_fallback(built):
    """Without a CUDA device every compute entry point must fail loudly."""
    import audio_formats_b200 as af
    if af.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(af.L3BError) as e:
        af.Context(0)
    assert e.value.code == -16 and "no CPU fallback" in str(e.value)
    from audio_formats_b200 import synth
    st = synth.generate(synth.SynthParams(nframes=12))
    with pytest.raises(af.L3BError):
        af.AudioStream().openFromMemory(st.data)


def test_product_does_not_reference_the_oracle():
    pkg = ROOT / "audio_formats_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.c*")) + list(pkg.rglob("*.h*")):
        text = f.read_text(errors="ignore")
        assert "l3o_" not in text and "import oracle" not in text and "libl3oracle" not in text, f
