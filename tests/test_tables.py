"""Structural checks of the derived constant tables (SURVEY.md 8c "what pins results instead")."""
import re
from fractions import Fraction
from pathlib import Path

import numpy as np

HDR = (Path(__file__).resolve().parent.parent / "audio_formats_b200" / "csrc" / "l3_tables_gen.h").read_text()


def arr(name, conv=int):
    m = re.search(name + r"\[[^\]]*\] = \{(.*?)\};", HDR, re.S)
    assert m, name
    return [conv(x.rstrip("f")) for x in re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?f?", m.group(1))]


def test_huffman_books_are_complete_prefix_codes():
    hlen, hcode, maxlen = arr("L3_HLEN"), arr("L3_HCODE"), arr("L3_BOOK_MAXLEN")
    counts = [4, 9, 9, 16, 16, 36, 36, 36, 64, 64, 64, 256, 256, 256, 256]  # ISO 11172-3 table 3-B.7 symbol counts
    for b in range(15):
        codes = [(hlen[b * 256 + s], hcode[b * 256 + s]) for s in range(256) if hlen[b * 256 + s]]
        assert len(codes) == counts[b]
        assert sum(Fraction(1, 1 << l) for l, _ in codes) == 1          # Kraft equality: complete
        assert max(l for l, _ in codes) == maxlen[b]
        for l, c in codes:                                                # prefix-free
            assert c < (1 << l)
            for l2, c2 in codes:
                if l2 > l:
                    assert (c2 >> (l2 - l)) != c
    assert maxlen[11] == 19 and maxlen[13] == 17 and maxlen[14] == 12 and maxlen[12] == 13


def test_count1_books():
    c1len, c1code = arr("L3_C1LEN"), arr("L3_C1CODE")
    for t in range(2):
        codes = [(c1len[t * 16 + f], c1code[t * 16 + f]) for f in range(16)]
        assert sum(Fraction(1, 1 << l) for l, _ in codes) == 1
    assert all(l == 4 for l in c1len[16:])   # book B is the fixed 4-bit code


def test_sfb_rows_sum_to_576():
    for name, w in (("L3_SFB_LONG", 23), ("L3_SFB_SHORT", 40), ("L3_SFB_MIXED", 40)):
        t = arr(name)
        assert len(t) == 8 * w
        for r in range(8):
            row = t[r * w:(r + 1) * w]
            assert sum(row) == 576 and row[-1] == 0
            assert all(v % 2 == 0 for v in row)   # pairs never straddle a band


def test_float_tables_sizes_and_symmetry():
    pow43 = arr("L3_POW43", float)
    assert len(pow43) == 129 and pow43[0] == 0 and pow43[1] == 1 and pow43[8] == 16 and pow43[27] == 81
    assert np.allclose(pow43, np.arange(129) ** (4 / 3), rtol=2e-7, atol=1e-6)
    win = arr("L3_WIN", float)
    assert len(win) == 240 and max(win) == 74992 and min(win) == -62684
    assert len(arr("L3_SEC", float)) == 24 and len(arr("L3_TWID9", float)) == 18 and len(arr("L3_AA", float)) == 16


def test_linbits_and_book_map():
    sel2book, linbits = arr("L3_SEL2BOOK"), arr("L3_LINBITS")
    assert [i for i, b in enumerate(sel2book) if b < 0] == [0, 4, 14]
    assert linbits[16:24] == [1, 2, 3, 4, 6, 8, 10, 13] and linbits[24:] == [4, 5, 6, 7, 8, 9, 11, 13]
    assert len(set(sel2book[16:24])) == 1 and len(set(sel2book[24:])) == 1
