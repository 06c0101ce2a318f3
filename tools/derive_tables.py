#!/usr/bin/env python3
"""Derive the ISO/IEC 11172-3 / 13818-3 Layer III constant tables in OUR layout.

Run at development time only (needs /root/reference, which does not exist on the GPU box):

    python tools/derive_tables.py            # writes audio_formats_b200/csrc/l3_tables_gen.h

What it does
------------
The reference decoder (source/audioformats/minimp3.d) stores the Huffman code books as
pre-built multi-level *decode trees* (`tabs`, minimp3.d:750-765, walked at :796-803) and the
count1 books as two small lookup arrays (`tab32`/`tab33`, :766-767, walked at :858-864).
We do not carry that blob.  This script walks every tree exactly the way `L3_huffman` does,
recovers the underlying canonical (codeword, length) -> (v0, v1) mapping -- i.e. the content
of ISO 11172-3 Annex B table 3-B.7 -- verifies it is a complete prefix code (Kraft sum == 1)
with the ISO symbol counts, and emits it as plain `hlen/hcode` arrays.  Our decoders (oracle,
CUDA) and the synthetic *encoder* all build their own lookup structures from these arrays.

The numeric float constants (pow43, window, twiddles ...) are facts of the algorithm and must be
bit-identical to the reference's literals, so they are re-emitted from the literals found in
the source, re-arranged into the layouts our kernels want (documented per table below).
"""
import re
import sys
from fractions import Fraction
from pathlib import Path

REF = Path("/root/reference/source/audioformats/minimp3.d")
OUT = Path(__file__).resolve().parent.parent / "audio_formats_b200" / "csrc" / "l3_tables_gen.h"


def grab(src: str, name: str, start: int = 0):
    """Return (list of numeric literal strings, end offset) of the array initialiser `name = [ ... ];`."""
    m = re.compile(r"\b" + re.escape(name) + r"\s*=\s*\[").search(src, start)
    if not m:
        raise SystemExit(f"array {name} not found")
    depth, i = 1, m.end()
    while depth:
        c = src[i]
        depth += (c == "[") - (c == "]")
        i += 1
    body = src[m.end(): i - 1]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    nums = re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?f?", body)
    return nums, i


def ints(xs):
    return [int(x) for x in xs]


def walk_pair_book(tabs, base):
    """Enumerate (code, length, v0, v1) of one big-values book by walking the decode tree
    like minimp3.d:795-803 (5-bit first peek, negative entry = link with width leaf&7 and
    offset -(leaf>>3), non-negative entry = leaf with length leaf>>8)."""
    out = {}

    def node(off, width, prefix, plen):
        for v in range(1 << width):
            leaf = tabs[base + off + v]
            if leaf < 0:
                node(-(leaf >> 3), leaf & 7, (prefix << width) | v, plen + width)
            else:
                ln = leaf >> 8
                assert ln <= width, (ln, width)
                code = (prefix << ln) | (v >> (width - ln))
                key = (code, plen + ln)
                sym = (leaf & 15, (leaf >> 4) & 15)
                assert out.setdefault(key, sym) == sym
    node(0, 5, 0, 0)
    return out


def walk_count1_book(tab, two_level):
    """Enumerate (code, length, flags) of a count1 book, minimp3.d:858-864."""
    out = {}
    for v in range(16):
        leaf = tab[v]
        if leaf & 8:
            ln = leaf & 7
            out[(v >> (4 - ln), ln)] = leaf >> 4
        else:
            assert two_level
            w = leaf & 3
            for u in range(1 << w):
                l2 = tab[(leaf >> 3) + u]
                assert l2 & 8
                ln = l2 & 7
                assert 4 < ln <= 4 + w
                code = ((v << w) | u) >> (4 + w - ln)
                out[(code, ln)] = l2 >> 4
    return out


def kraft(book):
    return sum(Fraction(1, 1 << ln) for (_, ln) in book)


def main():
    src = REF.read_text()
    tabs, _ = grab(src, "tabs"); tabs = ints(tabs)
    tabindex, _ = grab(src, "tabindex"); tabindex = ints(tabindex)
    linbits, _ = grab(src, "g_linbits"); linbits = ints(linbits)
    tab32, _ = grab(src, "tab32"); tab32 = ints(tab32)
    tab33, _ = grab(src, "tab33"); tab33 = ints(tab33)
    assert len(tabs) == 2164 and len(tabindex) == 32 and len(linbits) == 32

    # distinct books: map table_select -> book id
    bases = []
    sel2book = []
    for t in range(32):
        b = tabindex[t]
        if b == 0:
            sel2book.append(-1)          # all-zero book (table_select 0, 4, 14)
            continue
        if b not in bases:
            bases.append(b)
        sel2book.append(bases.index(b))
    nbooks = len(bases)
    assert nbooks == 15
    iso_counts = {1: 4, 2: 9, 3: 9, 5: 16, 6: 16, 7: 36, 8: 36, 9: 36, 10: 64, 11: 64, 12: 64,
                  13: 256, 15: 256, 16: 256, 24: 256}
    books = []
    maxlens = []
    for bi, b in enumerate(bases):
        bk = walk_pair_book(tabs, b)
        assert kraft(bk) == 1, "incomplete prefix code"
        first_sel = sel2book.index(bi)
        assert len(bk) == iso_counts[first_sel], (first_sel, len(bk))
        syms = sorted(bk.values())
        assert len(set(syms)) == len(syms)
        books.append(bk)
        maxlens.append(max(ln for (_, ln) in bk))
    c1 = [walk_count1_book(tab32, True), walk_count1_book(tab33, False)]
    for bk in c1:
        assert kraft(bk) == 1 and len(bk) == 16 and sorted(bk.values()) == list(range(16))

    # ---- float constant tables (literal text preserved so the float32 value is identical) ----
    pow43, _ = grab(src, "g_pow43")
    assert len(pow43) == 145
    pow43_pos = pow43[16:]                       # x = 0..128 ; negatives are exact mirrors
    for k in range(16):
        assert float(pow43[k].rstrip("f")) == -float(pow43[16 + k].rstrip("f"))
    expfrac, _ = grab(src, "g_expfrac")
    aa, _ = grab(src, "g_aa")
    twid9, _ = grab(src, "g_twid9")
    twid3, _ = grab(src, "g_twid3")
    mdctw, _ = grab(src, "g_mdct_window")
    sec, _ = grab(src, "g_sec")
    win, _ = grab(src, "g_win")
    pan, _ = grab(src, "g_pan")
    assert (len(expfrac), len(aa), len(twid9), len(twid3), len(mdctw), len(sec), len(win), len(pan)) == \
        (4, 16, 18, 6, 36, 24, 240, 14)

    scf_long, _ = grab(src, "g_scf_long"); scf_long = ints(scf_long)
    scf_short, _ = grab(src, "g_scf_short"); scf_short = ints(scf_short)
    scf_mixed, _ = grab(src, "g_scf_mixed"); scf_mixed = ints(scf_mixed)
    assert len(scf_long) == 8 * 23 and len(scf_short) == 8 * 40
    # g_scf_mixed rows have 37..40 entries in the source (the D compiler zero-fills to 40): split on the 0 terminator
    rows, cur = [], []
    for v in scf_mixed:
        cur.append(v)
        if v == 0:
            rows.append(cur); cur = []
    assert len(rows) == 8 and not cur
    scf_mixed_rows = [r + [0] * (40 - len(r)) for r in rows]
    for r in range(8):
        assert sum(scf_long[r * 23:(r + 1) * 23]) == 576
        assert sum(scf_short[r * 40:(r + 1) * 40]) == 576
        assert sum(scf_mixed_rows[r]) == 576, (r, sum(scf_mixed_rows[r]))

    partitions, _ = grab(src, "g_scf_partitions"); partitions = ints(partitions)
    scfc_decode, _ = grab(src, "g_scfc_decode"); scfc_decode = ints(scfc_decode)
    gmod, _ = grab(src, "g_mod"); gmod = ints(gmod)
    preamp, _ = grab(src, "g_preamp"); preamp = ints(preamp)
    halfrate, _ = grab(src, "halfrate"); halfrate = ints(halfrate)
    assert len(partitions) == 84 and len(scfc_decode) == 16 and len(gmod) == 24 and len(preamp) == 10
    assert len(halfrate) == 90

    def carr(ctype, name, vals, per_line=16, dims=""):
        s = f"static const {ctype} {name}{dims} = {{\n"
        for i in range(0, len(vals), per_line):
            s += "    " + ",".join(str(v) for v in vals[i:i + per_line]) + ",\n"
        return s + "};\n"

    o = []
    o.append("/* GENERATED by tools/derive_tables.py -- do not edit.\n"
             " * Layer III constant data in this project's own layout (see the script for provenance:\n"
             " * Huffman books recovered as canonical (length, codeword) lists = ISO 11172-3 table 3-B.7;\n"
             " * float constants keep the literal text of minimp3.d so the float32 values are identical). */\n"
             "#ifndef L3_TABLES_GEN_H\n#define L3_TABLES_GEN_H\n#include <stdint.h>\n")
    o.append(f"#define L3_NBOOKS {nbooks}\n")
    o.append("/* table_select (0..31) -> book id (0..14) or -1 for the all-zero book (selects 0, 4, 14). */")
    o.append(carr("int8_t", "L3_SEL2BOOK", sel2book, 32, "[32]"))
    o.append("/* escape bits per table_select (minimp3.d:769). */")
    o.append(carr("uint8_t", "L3_LINBITS", linbits, 32, "[32]"))
    o.append("/* longest codeword of each book. */")
    o.append(carr("uint8_t", "L3_BOOK_MAXLEN", maxlens, 32, f"[{nbooks}]"))
    hlen = []
    hcode = []
    for bk in books:
        ln_row = [0] * 256
        cd_row = [0] * 256
        for (code, ln), (v0, v1) in bk.items():
            ln_row[v0 * 16 + v1] = ln
            cd_row[v0 * 16 + v1] = code
        hlen += ln_row
        hcode += cd_row
    o.append("/* codeword length of the pair (v0,v1) at [book][v0*16+v1]; 0 = pair not in this book.\n"
             " * v0 is the FIRST value of the pair in the output order (minimp3.d:805-807). */")
    o.append(carr("uint8_t", "L3_HLEN", hlen, 16, f"[{nbooks}*256]"))
    o.append("/* codeword (right-aligned, MSB first) of the pair (v0,v1). */")
    o.append(carr("uint32_t", "L3_HCODE", hcode, 16, f"[{nbooks}*256]"))
    c1len, c1code = [], []
    for bk in c1:
        ln_row = [0] * 16
        cd_row = [0] * 16
        for (code, ln), flags in bk.items():
            ln_row[flags] = ln
            cd_row[flags] = code
        c1len += ln_row
        c1code += cd_row
    o.append("/* count1 books A (count1_table=0) and B (=1): [book*16 + flags], flags = v0<<3|v1<<2|v2<<1|v3\n"
             " * (nonzero markers of the quad, minimp3.d:874-878). */")
    o.append(carr("uint8_t", "L3_C1LEN", c1len, 16, "[32]"))
    o.append(carr("uint8_t", "L3_C1CODE", c1code, 16, "[32]"))

    o.append("/* scalefactor-band widths: [sr_idx][..] with 0 terminator (minimp3.d:489-519); every row sums to 576. */")
    o.append(carr("uint8_t", "L3_SFB_LONG", scf_long, 23, "[8*23]"))
    o.append(carr("uint8_t", "L3_SFB_SHORT", scf_short, 40, "[8*40]"))
    o.append(carr("uint8_t", "L3_SFB_MIXED", sum(scf_mixed_rows, []), 40, "[8*40]"))
    o.append("/* scalefactor partition counts [3][28] (minimp3.d:661-665), slen decode (:674), LSF radix (:680), preemphasis (:707). */")
    o.append(carr("uint8_t", "L3_SCF_PARTITIONS", partitions, 28, "[3*28]"))
    o.append(carr("uint8_t", "L3_SCFC_DECODE", scfc_decode, 16, "[16]"))
    o.append(carr("uint8_t", "L3_LSF_MOD", gmod, 24, "[24]"))
    o.append(carr("uint8_t", "L3_PREAMP", preamp, 10, "[10]"))
    o.append("/* half bitrates [mpeg1?][layer-1][bitrate_idx] (minimp3.d:251-255). */")
    o.append(carr("uint8_t", "L3_HALFRATE", halfrate, 15, "[2*3*15]"))

    o.append("/* x^(4/3) for x = 0..128 (minimp3.d:722-725, positive half; the negative half is its exact mirror). */")
    o.append(carr("float", "L3_POW43", pow43_pos, 8, "[129]"))
    o.append("/* 2^-30 * 2^(-k/4), k=0..3 (minimp3.d:648-649). */")
    o.append(carr("float", "L3_EXPFRAC", expfrac, 4, "[4]"))
    o.append("/* alias-reduction butterflies: cs[8] then ca[8] (minimp3.d:1004-1007). */")
    o.append(carr("float", "L3_AA", aa, 8, "[16]"))
    o.append("/* IMDCT-36 twiddles (minimp3.d:1065-1067), IMDCT-12 twiddles (:1113). */")
    o.append(carr("float", "L3_TWID9", twid9, 9, "[18]"))
    o.append(carr("float", "L3_TWID3", twid3, 6, "[6]"))
    o.append("/* IMDCT window: [0]=normal/start, [1]=stop (minimp3.d:1154-1157). */")
    o.append(carr("float", "L3_MDCT_WINDOW", mdctw, 9, "[36]"))
    o.append("/* DCT-32 first-stage secants, 3 per butterfly index (minimp3.d:1234-1236). */")
    o.append(carr("float", "L3_SEC", sec, 6, "[24]"))
    o.append("/* intensity-stereo pan pairs (kl,kr) for is_pos 0..6 (minimp3.d:930). */")
    o.append(carr("float", "L3_PAN", pan, 14, "[14]"))
    # synthesis window, transposed for per-lane register residency:
    # reference order (minimp3.d:1336-1352, consumed at :1388-1395): for i = 14..0, for k = 0..7: w0, w1.
    # ours: L3_WIN[(k*2 + c)*15 + i]  with c=0 -> w0, c=1 -> w1   (i = the reference's loop variable)
    wt = [None] * 240
    p = 0
    for i in range(14, -1, -1):
        for k in range(8):
            for c in range(2):
                wt[(k * 2 + c) * 15 + i] = win[p] + ".0f" if "." not in win[p] and "f" not in win[p] else win[p]
                p += 1
    o.append("/* synthesis window, L3_WIN[(k*2+c)*15 + i]: tap k (0..7), c=0:w0 c=1:w1, i = inner index 0..14\n"
             " * of minimp3.d:1371-1395 (the reference stores it i-major starting from i=14). */")
    o.append(carr("float", "L3_WIN", wt, 15, "[240]"))
    o.append("#endif\n")
    text = "\n".join(o)
    # float literal hygiene: make sure every float literal ends in 'f'
    def fix_float_arrays(t):
        def fx(m):
            body = m.group(2)
            body = re.sub(r"(?<![\w.])(-?\d+(?:\.\d*)?(?:[eE][-+]?\d+)?)(f?)(?=[,\s])",
                          lambda q: (q.group(1) if "." in q.group(1) or "e" in q.group(1).lower()
                                     else q.group(1) + ".0") + "f", body)
            return m.group(1) + body + m.group(3)
        return re.sub(r"(static const float \w+\[\d+\] = \{\n)(.*?)(\};)", fx, t, flags=re.S)
    text = fix_float_arrays(text)
    OUT.write_text(text)
    print(f"wrote {OUT} ({len(text)} bytes); books={nbooks} maxlens={maxlens}")
    print("sel2book", sel2book)


if __name__ == "__main__":
    sys.exit(main())
