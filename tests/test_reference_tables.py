"""Pins the constant tables the product AND the oracle share (audio_formats_b200/csrc/l3_tables_gen.h) to data the reference
itself holds: the literal tables of /root/reference/source/audioformats/minimp3.d, parsed at test time.

A wrong entry in l3_tables_gen.h would be common-mode (GPU and oracle would agree with each other and both be wrong), so:
  * the Huffman books are checked by DECODING every 19-bit prefix (the longest code is 19 bits) with the reference's own tree
    blob `tabs` through the reference's own walk (minimp3.d:797-804: 5-bit first peek, negative leaf = sub-table) and comparing
    value pair and code length with the canonical (length, code) books for every table_select;
  * the count1 books by decoding every 6-bit prefix through tab32 / tab33 (minimp3.d:857-864);
  * every other integer and float table element by element.
Skipped when the reference tree is absent (it does not travel to the GPU box); nothing here needs a GPU.
"""
import re
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference/source/audioformats/minimp3.d")
GEN = Path(__file__).resolve().parent.parent / "audio_formats_b200" / "csrc" / "l3_tables_gen.h"

pytestmark = pytest.mark.skipif(not REF.exists(), reason="the reference tree is not present on this machine")

NUM = r"-?\d+\.?\d*(?:[eE][-+]?\d+)?"


def ref_array(name, conv=float):
    """All numeric literals of `name = [ ... ];` in the D source."""
    src = REF.read_text()
    m = re.search(r"\b" + re.escape(name) + r"\s*=\s*\[(.*?)\];", src, re.S)
    assert m, name
    body = re.sub(r"//[^\n]*", "", m.group(1))
    return [conv(x) for x in re.findall(NUM + r"(?=f?\s*[,\]\s]|f?$)", body)]


def ref_rows(name, width):
    """A two-dimensional literal, each row zero-padded to the declared width (rows of g_scf_mixed are written short)."""
    src = REF.read_text()
    m = re.search(r"\b" + re.escape(name) + r"\s*=\s*\[(.*?)\];", src, re.S)
    assert m, name
    out = []
    for row in re.findall(r"\[([^\[\]]*)\]", m.group(1)):
        vals = [int(x) for x in re.findall(r"-?\d+", row)]
        assert len(vals) <= width, (name, len(vals))
        out += vals + [0] * (width - len(vals))
    return out


def gen_array(name, conv=float):
    m = re.search(re.escape(name) + r"\[[^\]]*\] = \{(.*?)\};", GEN.read_text(), re.S)
    assert m, name
    return [conv(x) for x in re.findall(NUM, re.sub(r"(?<=\d)f\b", "", m.group(1)))]


def test_huffman_books_decode_like_the_reference_tree_for_every_19_bit_prefix():
    tabs = np.array(ref_array("tabs", int), dtype=np.int32)
    tabindex = ref_array("tabindex", int)
    assert len(tabs) == 2164 and len(tabindex) == 32
    hlen = np.array(gen_array("L3_HLEN", int)).reshape(15, 256)
    hcode = np.array(gen_array("L3_HCODE", int), dtype=np.int64).reshape(15, 256)
    sel2book = gen_array("L3_SEL2BOOK", int)
    prefixes = np.arange(1 << 19, dtype=np.int64)
    cache0 = prefixes << 13                      # the prefix at the top of the reference's 32-bit bs_cache
    for sel in range(32):
        # ---- the reference's walk (minimp3.d:797-804) over its own blob, vectorised over all prefixes ----
        book_base = tabindex[sel]
        cache = cache0.copy()
        used = np.zeros(len(prefixes), np.int64)
        w = np.full(len(prefixes), 5, np.int64)
        leaf = tabs[book_base + (cache >> (32 - 5))].astype(np.int64)
        for _ in range(8):
            neg = leaf < 0
            if not neg.any():
                break
            cache = np.where(neg, (cache << w) & 0xFFFFFFFF, cache)     # FLUSH_BITS(w)
            used = np.where(neg, used + w, used)
            w = np.where(neg, leaf & 7, w)                                  # two's complement: w = leaf & 7
            idx = np.where(neg, book_base + (cache >> (32 - np.maximum(w, 1))) - (leaf >> 3), 0)   # arithmetic shift of the negative leaf
            leaf = np.where(neg, tabs[idx], leaf)
        assert (leaf >= 0).all(), sel
        ref_len = used + (leaf >> 8)
        ref_x, ref_y = leaf & 15, (leaf >> 4) & 15            # first value of the pair in the low nibble (minimp3.d:805-806)
        # ---- the canonical books the product and the oracle are built from ----
        b = sel2book[sel]
        if b < 0:                                              # table_select 0 / 4 / 14: the all-zero book, no bits
            assert (ref_len == 0).all() and (ref_x == 0).all() and (ref_y == 0).all(), sel
            continue
        want_len = np.zeros(1 << 19, np.int64)
        want_x = np.zeros(1 << 19, np.int64)
        want_y = np.zeros(1 << 19, np.int64)
        for s in range(256):
            ln = int(hlen[b, s])
            if not ln:
                continue
            lo = int(hcode[b, s]) << (19 - ln)
            hi = lo + (1 << (19 - ln))
            want_len[lo:hi] = ln
            want_x[lo:hi] = s >> 4                             # symbol s = x*16 + y, x is decoded first
            want_y[lo:hi] = s & 15
        assert (want_len > 0).all(), sel                        # complete code: every prefix decodes
        assert np.array_equal(ref_len, want_len), (sel, np.argwhere(ref_len != want_len)[:3])
        assert np.array_equal(ref_x, want_x) and np.array_equal(ref_y, want_y), sel
    assert gen_array("L3_LINBITS", int) == ref_array("g_linbits", int)


def test_count1_books_decode_like_tab32_tab33():
    c1len = np.array(gen_array("L3_C1LEN", int)).reshape(2, 16)
    c1code = np.array(gen_array("L3_C1CODE", int)).reshape(2, 16)
    for t, name in enumerate(("tab32", "tab33")):
        tab = ref_array(name, int)
        for v in range(64):                                     # six bits are enough: the longest count1 code is 6 bits
            cache = v << 26
            leaf = tab[cache >> 28]
            if not (leaf & 8):
                leaf = tab[(leaf >> 3) + (((cache << 4) & 0xFFFFFFFF) >> (32 - (leaf & 3)))]
            ln, flags = leaf & 7, leaf >> 4                     # flags: bit 3 = first value of the quad (128 >> 0 of the leaf)
            match = [f for f in range(16) if (v >> (6 - c1len[t, f])) == c1code[t, f]]
            assert len(match) == 1, (name, v)
            assert c1len[t, match[0]] == ln and match[0] == flags, (name, v, leaf)


@pytest.mark.parametrize("ref_name,gen_name,conv", [
    ("g_scf_partitions", "L3_SCF_PARTITIONS", int), ("g_scfc_decode", "L3_SCFC_DECODE", int), ("g_mod", "L3_LSF_MOD", int),
    ("g_preamp", "L3_PREAMP", int), ("g_expfrac", "L3_EXPFRAC", float), ("g_aa", "L3_AA", float), ("g_twid9", "L3_TWID9", float),
    ("g_twid3", "L3_TWID3", float), ("g_mdct_window", "L3_MDCT_WINDOW", float), ("g_sec", "L3_SEC", float), ("g_pan", "L3_PAN", float),
])
def test_table_equals_the_reference_literal(ref_name, gen_name, conv):
    ref, gen = ref_array(ref_name, conv), gen_array(gen_name, conv)
    if conv is float:
        ref, gen = np.array(ref, np.float32), np.array(gen, np.float32)       # both are parsed by a float32 compiler
        assert ref.shape == gen.shape and np.array_equal(ref.view(np.uint32), gen.view(np.uint32)), (ref_name, ref[:4], gen[:4])
    else:
        assert ref == gen, ref_name


def test_scalefactor_band_tables():
    for ref_name, gen_name, width in (("g_scf_long", "L3_SFB_LONG", 23), ("g_scf_short", "L3_SFB_SHORT", 40), ("g_scf_mixed", "L3_SFB_MIXED", 40)):
        ref, gen = ref_rows(ref_name, width), gen_array(gen_name, int)
        assert len(ref) == 8 * width and ref == gen, ref_name


def test_pow43_and_synthesis_window():
    pow43 = np.array(ref_array("g_pow43"), np.float32)
    assert len(pow43) == 145
    gen = np.array(gen_array("L3_POW43"), np.float32)
    assert np.array_equal(pow43[16:].view(np.uint32), gen.view(np.uint32))
    assert np.array_equal(pow43[1:16], -pow43[17:32]) and pow43[0] == 0 and not np.signbit(pow43[0])   # entry 0 is +0 (minimp3.d:723)
    win = np.array(ref_array("g_win"), np.float32)
    assert len(win) == 240
    ours = np.array(gen_array("L3_WIN"), np.float32)
    # reference order (minimp3.d:1336-1352, consumed at :1371-1395): for i = 14..0, for k = 0..7: w0, w1; ours: [(k*2 + c)*15 + i]
    p = 0
    for i in range(14, -1, -1):
        for k in range(8):
            for c in range(2):
                assert ours[(k * 2 + c) * 15 + i] == win[p], (i, k, c)
                p += 1


def test_layer12_tables():
    """audio_formats_b200/csrc/l12_tables.h against the literals of minimp3.d:284-398 (the oracle carries its own copy)."""
    l12 = (GEN.parent / "l12_tables.h").read_text()

    def mine(name, conv):
        m = re.search(re.escape(name) + r"\[[^=]*=\s*\{(.*?)\};", l12, re.S)
        assert m, name
        return [conv(x) for x in re.findall(NUM, m.group(1))]

    assert mine("L12_BITALLOC_CODE_TAB", int) == ref_array("g_bitalloc_code_tab", int)
    ref_deq = np.array(ref_array("g_deq_L12", float), np.float64).astype(np.float32)      # D: double literals converted to float
    my_deq = np.array(mine("L12_DEQ", float), np.float64).astype(np.float32)
    assert len(ref_deq) == 54 and np.array_equal(ref_deq.view(np.uint32), my_deq.view(np.uint32))
    src = REF.read_text()
    for ref_name, my_name in (("g_alloc_L1", "L12_ALLOC_L1"), ("g_alloc_L2M2", "L12_ALLOC_L2M2"), ("g_alloc_L2M1", "L12_ALLOC_L2M1"),
                              ("g_alloc_L2M1_lowrate", "L12_ALLOC_L2M1_LOWRATE")):
        m = re.search(re.escape(ref_name) + r"\s*=\s*\[(.*?)\];", src, re.S)
        assert m, ref_name
        want = [int(x) for x in re.findall(r"L12_subband_alloc_t\(([^)]*)\)", m.group(1)) for x in x.split(",")]
        assert mine(my_name, int)[-len(want):] == want and len(want) % 3 == 0, ref_name
