// Exhaustive host-side check of the 32-bit Huffman LUT (l3_device_tables.hpp) that the big_values and count1 kernels
// use: every code of every book is pushed through the same lookup sequence as the kernels (root table, sub-table walk
// inside one 32-bit window, leaf fields), with random bits behind it, and must give back its symbol, its length and
// its sign handling.  Exit code 0 = all good; prints the first mismatch otherwise.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "l3_device_tables.hpp"

using namespace l3b;

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state; }

int main() {
    HuffLut32 L = build_huff_lut32();
    long checked = 0;
    for (int b = 0; b < L3_NBOOKS; b++) {
        bool lin = false;
        for (int sel = 0; sel < 32; sel++) lin |= (L3_SEL2BOOK[sel] == b && L3_LINBITS[sel] != 0);
        const uint32_t base = L.base[b];
        const uint32_t sh = 32u - L.root_bits[b];
        for (int s = 0; s < 256; s++) {
            const int len = L3_HLEN[b * 256 + s];
            if (!len) continue;
            const uint32_t code = L3_HCODE[b * 256 + s];
            const int a0 = s >> 4, a1 = s & 15;
            for (int rep = 0; rep < 4; rep++) {
                const uint32_t tail = rnd();
                const uint32_t bits = (code << (32 - len)) | (len < 32 ? (tail >> len) : 0u);
                // ---- the kernel's lookup (l3_entropy.cu, l3_huff_big_kernel) ----
                uint32_t e = L.entries[base + (bits >> sh)];
                uint32_t used = 0;
                int hops = 0;
                while ((int32_t)e < 0) {
                    used += (e >> 22) & 31u;
                    const uint32_t s2 = (e >> 16) & 31u;
                    e = L.entries[base + (e & 0xFFFFu) + ((bits << used) >> s2)];
                    if (++hops > 3) { printf("book %d sym %d: endless link walk\n", b, s); return 1; }
                }
                const int g0 = (int)(e & 15u), g1 = (int)((e >> 16) & 15u), glen = (int)((e >> 8) & 15u), gtot = (int)((e >> 4) & 15u);
                const int n0 = (int)((e >> 24) & 1u), n1 = (int)((e >> 15) & 1u), esc = (int)((e >> 30) & 1u);
                if (g0 != a0 || g1 != a1 || (int)used + glen != len || n0 != (a0 != 0) || n1 != (a1 != 0) ||
                    gtot != glen + n0 + n1 || esc != (lin && (a0 == 15 || a1 == 15)) || (e & 0x1F003000u & ~0x01000000u)) {
                    printf("book %d sym %d (len %d): got a0=%d a1=%d used=%u len=%d tot=%d n0=%d n1=%d esc=%d entry=%08x\n", b, s, len, g0,
                           g1, used, glen, gtot, n0, n1, esc, e);
                    return 1;
                }
                if (!esc) {   // the sign fix-up of the fast path
                    const uint32_t sb = (bits << used) << ((e >> 8) & 31u);
                    const uint32_t s0 = (sb >> 31) & (e >> 24);
                    const uint32_t sb1 = sb << ((e >> 24) & 31u);
                    const uint32_t s1 = (sb1 >> 31) & (e >> 15) & 1u;
                    const uint32_t inc = s0 | (s1 << 16);
                    const uint32_t pk = ((e & 0x000F000Fu) ^ (inc * 0xFFFFu)) + inc;
                    // expected: sign bits follow the code, one per non-zero value, in order
                    int pos = len, v0 = a0, v1 = a1;
                    if (a0) { if ((bits >> (31 - pos)) & 1u) v0 = -a0; pos++; }
                    if (a1) { if ((bits >> (31 - pos)) & 1u) v1 = -a1; pos++; }
                    const uint32_t want = ((uint32_t)v0 & 0xFFFFu) | ((uint32_t)v1 << 16);
                    if (pk != want || (int)used + gtot != pos) {
                        printf("book %d sym %d: packed %08x, expected %08x (bits %08x)\n", b, s, pk, want, bits);
                        return 1;
                    }
                }
                checked++;
            }
        }
    }
    // the all-zero book: one bit of root, zero-length leaves
    if (L.entries[L.base[L3_NBOOKS]] != 0 || L.entries[L.base[L3_NBOOKS] + 1] != 0 || L.root_bits[L3_NBOOKS] != 1) { printf("zero book\n"); return 1; }
    // count1: code table + value table (l3_huff_c1_kernel)
    for (int t = 0; t < 2; t++)
        for (int f = 0; f < 16; f++) {
            const int len = L3_C1LEN[t * 16 + f];
            const uint32_t code = L3_C1CODE[t * 16 + f];
            for (int signs = 0; signs < 16; signs++) {
                const uint32_t bits = (code << (32 - len)) | ((uint32_t)signs << (28 - len)) | (rnd() >> (len + 4));
                const uint32_t e = L.c1code[t][bits >> 26];
                const uint32_t glen = e & 15u, flags = (e >> 4) & 15u, tot = e >> 8;
                if ((int)glen != len || (int)flags != f || tot != glen + (uint32_t)__builtin_popcount(flags)) {
                    printf("count1 table %d flags %d: len %u flags %u tot %u\n", t, f, glen, flags, tot);
                    return 1;
                }
                const uint32_t sb = bits << glen;
                const uint32_t* v = L.c1val[(e & 0xF0u) | (sb >> 28)];
                int want[4], k = 0;
                for (int i = 0; i < 4; i++) { want[i] = 0; if (f & (8 >> i)) { want[i] = (signs & (8 >> k)) ? -1 : 1; k++; } }
                const uint32_t w01 = ((uint32_t)want[0] & 0xFFFFu) | ((uint32_t)want[1] << 16);
                const uint32_t w23 = ((uint32_t)want[2] & 0xFFFFu) | ((uint32_t)want[3] << 16);
                if (v[0] != w01 || v[1] != w23) { printf("count1 values table %d flags %d signs %d\n", t, f, signs); return 1; }
                checked++;
            }
        }
    printf("lut32 ok: %ld lookups, %zu entries, root bits %d\n", checked, L.entries.size(), L3B_HUFF_ROOT_BITS);
    return 0;
}
