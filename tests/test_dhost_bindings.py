"""The (uncompiled) D bindings declare exactly the functions of the C header, so they cannot drift silently."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_d_bindings_match_header():
    h = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "l3b200.h").read_text(), flags=re.S)
    c_names = set(re.findall(r"\b(l3b_[a-z0-9_]+)\s*\(", h))
    d = (ROOT / "audio_formats_b200" / "dhost" / "l3b200.d").read_text()
    d = re.sub(r"/\*\*.*?\*/", "", d, flags=re.S)
    d_names = set(re.findall(r"\b(l3b_[a-z0-9_]+)\s*\(", d))
    assert c_names == d_names, (sorted(c_names - d_names), sorted(d_names - c_names))
    # error codes agree
    for name, val in re.findall(r"#define (L3B_[A-Z_]+) \(?(-?\d+)\)?", h):
        m = re.search(r"enum %s = (-?\d+);" % name, d)
        assert m and int(m.group(1)) == int(val), name
