// l3_device_tables.hpp -- host-side construction of the lookup structures the kernels use.
// Everything is derived at context creation from the canonical data in l3_tables_gen.h.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "l3_tables_gen.h"

namespace l3b {

// ---- Huffman decode LUT (our own layout) ---------------------------------------------------------
// 16-bit entries, per book a root table of 2^root_bits entries followed by its sub-tables.
//   leaf : bit15 = 0 : [v3:1 @13][v2:1 @12][len:4 @8][v1:4 @4][v0:4 @0]   len = bits consumed at THIS level (0..8)
//   link : bit15 = 1 : [width-1:3 @12][offset:12 @0]    offset relative to the book's base
// Book 15 (index L3_NBOOKS) is the all-zero book used by table_select 0/4/14 (minimp3.d:768: tabindex 0).
// Books 16 and 17 are the count1 books A and B (minimp3.d:766-767) as 6-bit root tables in the same leaf format:
// v0..v3 are the non-zero flags of the quad, so the pair decoder treats a quad as two pairs of 0/1 magnitudes.
struct HuffLut {
    std::vector<uint16_t> entries;
    uint16_t base[L3_NBOOKS + 3];
    uint8_t root_bits[L3_NBOOKS + 3];
    uint8_t count1[2][64];  // 6-bit peek -> flags<<4 | len
};

namespace detail {
inline int find_code(int book, int len, uint32_t code) {
    for (int s = 0; s < 256; s++)
        if (L3_HLEN[book * 256 + s] == len && L3_HCODE[book * 256 + s] == code) return s;
    return -1;
}
inline int longest_under(int book, uint32_t prefix, int plen) {
    int m = 0;
    for (int s = 0; s < 256; s++) {
        int l = L3_HLEN[book * 256 + s];
        if (l > plen && (L3_HCODE[book * 256 + s] >> (l - plen)) == prefix && l > m) m = l;
    }
    return m;
}
inline void build_level(int book, std::vector<uint16_t>& e, size_t book_base, uint32_t prefix, int plen, int width) {
    size_t at = e.size();
    e.resize(at + ((size_t)1 << width), 0);
    for (uint32_t v = 0; v < (1u << width); v++) {
        uint32_t bits = (prefix << width) | v;
        bool done = false;
        for (int l = plen + 1; l <= plen + width && !done; l++) {
            int s = find_code(book, l, bits >> (plen + width - l));
            if (s >= 0) {
                e[at + v] = (uint16_t)(((l - plen) << 8) | ((s & 15) << 4) | (s >> 4));  // s = v0*16+v1
                done = true;
            }
        }
        if (!done) {
            int rest = longest_under(book, bits, plen + width) - (plen + width);
            int w = rest > 8 ? 8 : rest;
            size_t child = e.size() - book_base;
            e[at + v] = (uint16_t)(0x8000u | ((uint32_t)(w - 1) << 12) | (uint32_t)child);
            build_level(book, e, book_base, bits, plen + width, w);
        }
    }
}
}  // namespace detail

inline HuffLut build_huff_lut() {
    HuffLut L;
    for (int b = 0; b < L3_NBOOKS; b++) {
        int rb = L3_BOOK_MAXLEN[b] < 8 ? L3_BOOK_MAXLEN[b] : 8;
        L.base[b] = (uint16_t)L.entries.size();
        L.root_bits[b] = (uint8_t)rb;
        detail::build_level(b, L.entries, L.entries.size(), 0, 0, rb);
    }
    L.base[L3_NBOOKS] = (uint16_t)L.entries.size();
    L.root_bits[L3_NBOOKS] = 1;
    L.entries.push_back(0);
    L.entries.push_back(0);
    for (int t = 0; t < 2; t++) {
        L.base[L3_NBOOKS + 1 + t] = (uint16_t)L.entries.size();
        L.root_bits[L3_NBOOKS + 1 + t] = 6;
        for (int v = 0; v < 64; v++) {
            uint16_t e = 0;
            for (int f = 0; f < 16; f++) {
                int ln = L3_C1LEN[t * 16 + f];
                if ((v >> (6 - ln)) == L3_C1CODE[t * 16 + f])
                    e = (uint16_t)((ln << 8) | ((f >> 3) & 1) | (((f >> 2) & 1) << 4) | (((f >> 1) & 1) << 12) | ((f & 1) << 13));
            }
            L.entries.push_back(e);
        }
    }
    memset(L.count1, 0, sizeof L.count1);
    for (int t = 0; t < 2; t++)
        for (int v = 0; v < 64; v++)
            for (int f = 0; f < 16; f++) {
                int ln = L3_C1LEN[t * 16 + f];
                if ((v >> (6 - ln)) == L3_C1CODE[t * 16 + f]) L.count1[t][v] = (uint8_t)((f << 4) | ln);
            }
    return L;
}

// ---- Huffman decode LUT, 32-bit entries (the lane-decoupled entropy kernels) ---------------------------
// Per book a root table of 2^root_bits entries followed by its sub-tables, fields placed so that the common path
// needs no masking beyond what the shifter does for free:
//   leaf : [a0:4 @0][tot:4 @4][len:4 @8][n1 @15][a1:4 @16][n0 @24][esc @30]
//          len = code bits consumed at THIS level (0..8); n0/n1 = a0/a1 non-zero; tot = len + n0 + n1 (code + sign bits);
//          esc = the book has linbits and a0 or a1 is 15 (minimp3.d:805-813): taken out of line
//   link : bit31 = 1 : [offset:16 @0][32-width:5 @16][adv:5 @22]   offset relative to the book's base, width of the
//          sub-table, adv = bits to consume before descending (the width of the table holding the link)
// Book L3_NBOOKS is the all-zero book (table_select 0/4/14).  The count1 books are separate tables:
//   c1code[t][6-bit peek] = len | flags<<4 | (len + popcount(flags))<<8       (minimp3.d:857-866)
//   c1val[flags<<4 | s]   = the quad's two packed int16x2 words when the next four bits are s: sign bits are consumed
//                           by the non-zero values in order v0..v3 (minimp3.d:869-878)
#ifndef L3B_HUFF_ROOT_BITS
#define L3B_HUFF_ROOT_BITS 9   // root-table width of the big_values books: measured entropy stage 8: 8.19 ms (18 KB of LUT), 9: 8.06 (27 KB), 10: 8.45 (43 KB, one CTA fewer per SM)
#endif
struct HuffLut32 {
    std::vector<uint32_t> entries;
    uint32_t base[L3_NBOOKS + 1];
    uint8_t root_bits[L3_NBOOKS + 1];
    uint16_t c1code[2][64];
    uint32_t c1val[256][2];
};

namespace detail {
inline void build_level32(int book, bool has_linbits, std::vector<uint32_t>& e, size_t book_base, uint32_t prefix, int plen, int width) {
    size_t at = e.size();
    e.resize(at + ((size_t)1 << width), 0);
    for (uint32_t v = 0; v < (1u << width); v++) {
        uint32_t bits = (prefix << width) | v;
        bool done = false;
        for (int l = plen + 1; l <= plen + width && !done; l++) {
            int s = find_code(book, l, bits >> (plen + width - l));
            if (s >= 0) {
                const uint32_t a0 = (uint32_t)(s >> 4), a1 = (uint32_t)(s & 15), len = (uint32_t)(l - plen);
                const uint32_t n0 = a0 != 0, n1 = a1 != 0;
                const uint32_t esc = has_linbits && (a0 == 15 || a1 == 15);
                e[at + v] = a0 | ((len + n0 + n1) << 4) | (len << 8) | (n1 << 15) | (a1 << 16) | (n0 << 24) | (esc << 30);
                done = true;
            }
        }
        if (!done) {
            int rest = longest_under(book, bits, plen + width) - (plen + width);
            int w = rest > 8 ? 8 : rest;
            size_t child = e.size() - book_base;
            e[at + v] = 0x80000000u | (uint32_t)child | ((uint32_t)(32 - w) << 16) | ((uint32_t)width << 22);
            build_level32(book, has_linbits, e, book_base, bits, plen + width, w);
        }
    }
}
}  // namespace detail

inline HuffLut32 build_huff_lut32() {
    HuffLut32 L;
    for (int b = 0; b < L3_NBOOKS; b++) {
        int rb = L3_BOOK_MAXLEN[b] < L3B_HUFF_ROOT_BITS ? L3_BOOK_MAXLEN[b] : L3B_HUFF_ROOT_BITS;
        bool lin = false;
        for (int sel = 0; sel < 32; sel++) lin |= (L3_SEL2BOOK[sel] == b && L3_LINBITS[sel] != 0);
        L.base[b] = (uint32_t)L.entries.size();
        L.root_bits[b] = (uint8_t)rb;
        detail::build_level32(b, lin, L.entries, L.entries.size(), 0, 0, rb);
    }
    L.base[L3_NBOOKS] = (uint32_t)L.entries.size();
    L.root_bits[L3_NBOOKS] = 1;
    L.entries.push_back(0);
    L.entries.push_back(0);
    memset(L.c1code, 0, sizeof L.c1code);
    for (int t = 0; t < 2; t++)
        for (int v = 0; v < 64; v++)
            for (int f = 0; f < 16; f++) {
                int ln = L3_C1LEN[t * 16 + f];
                if ((v >> (6 - ln)) == L3_C1CODE[t * 16 + f])
                    L.c1code[t][v] = (uint16_t)(ln | (f << 4) | ((ln + __builtin_popcount((unsigned)f)) << 8));
            }
    for (int f = 0; f < 16; f++)
        for (int s = 0; s < 16; s++) {
            int val[4], k = 0;
            for (int i = 0; i < 4; i++) {
                val[i] = 0;
                if (f & (8 >> i)) { val[i] = (s & (8 >> k)) ? -1 : 1; k++; }
            }
            L.c1val[f * 16 + s][0] = ((uint32_t)val[0] & 0xFFFFu) | ((uint32_t)val[1] << 16);
            L.c1val[f * 16 + s][1] = ((uint32_t)val[2] & 0xFFFFu) | ((uint32_t)val[3] << 16);
        }
    return L;
}

// ---- scalefactor-band maps -----------------------------------------------------------------------
// kind: 0 long, 1 short, 2 mixed.   sfb_of_pair[row][kind][p] = sfb index of coefficients 2p, 2p+1.
struct SfbMaps {
    uint8_t sfb_of_pair[8][3][288];
    uint8_t width[8][3][40];
    uint16_t start[8][3][40];
    uint16_t perm[8][2][576];  // [row][0 short / 1 mixed][k] = source index of output k for L3_reorder (minimp3.d:984-1000)
};

inline const uint8_t* sfb_row(int row, int kind) {
    return kind == 0 ? L3_SFB_LONG + row * 23 : kind == 1 ? L3_SFB_SHORT + row * 40 : L3_SFB_MIXED + row * 40;
}

inline void build_sfb_maps(SfbMaps* M) {
    memset(M, 0, sizeof *M);
    for (int row = 0; row < 8; row++) {
        for (int kind = 0; kind < 3; kind++) {
            const uint8_t* t = sfb_row(row, kind);
            int pos = 0;
            for (int i = 0; i < 40 && (kind == 0 ? i < 23 : true) && t[i]; i++) {
                M->width[row][kind][i] = t[i];
                M->start[row][kind][i] = (uint16_t)pos;
                for (int k = 0; k < t[i]; k += 2) M->sfb_of_pair[row][kind][(pos + k) >> 1] = (uint8_t)i;
                pos += t[i];
            }
        }
        const bool mpeg1 = row >= 5;
        for (int mixed = 0; mixed < 2; mixed++) {
            for (int k = 0; k < 576; k++) M->perm[row][mixed][k] = (uint16_t)k;
            int n_long_bands = mixed ? (2 << (row == 1 ? 1 : 0)) : 0;  // minimp3.d:1218 (<<1 only at 8 kHz, sfb row 1)
            int n_long_sfb = mixed ? (mpeg1 ? 8 : 6) : 0;
            const uint8_t* sfb = sfb_row(row, mixed ? 2 : 1) + n_long_sfb;
            int base = n_long_bands * 18, src = 0, dst = 0;
            for (; *sfb; sfb += 3) {
                int len = *sfb;
                for (int i = 0; i < len; i++, src++) {
                    for (int w = 0; w < 3; w++, dst++) {
                        // the reference overruns its buffer for 8 kHz mixed blocks (UB); we stay inside 576
                        if (base + dst < 576 && base + src + w * len < 576)
                            M->perm[row][mixed][base + dst] = (uint16_t)(base + src + w * len);
                    }
                }
                src += 2 * len;
            }
        }
    }
}

}  // namespace l3b
