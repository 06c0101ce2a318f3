/* l12synth.c -- deterministic synthetic MPEG-1/2 Layer I and Layer II bitstream generator (bench/test infrastructure).
 *
 * Writes legal frames at the syntax level (ISO 11172-3 2.4.1.5/2.4.1.6, 13818-3 LSF): header, optional CRC word, bit
 * allocation, scfsi (Layer II), scalefactors, sample codes, from a seeded SplitMix64 stream.  Allocation choices follow the
 * table selection the reference decoder makes (minimp3.d:284-350), and every frame fits its size exactly (zero stuffing
 * at the end).  No psychoacoustics: the content is noise with a spectral tilt, at -20 dBFS or so. */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    uint64_t seed;
    int layer;         /* 1 or 2 */
    int hz;            /* 32000/44100/48000 (MPEG-1); 16000/22050/24000 (MPEG-2 LSF, Layer II only uses its own table) */
    int nch;           /* 1 or 2 */
    int bitrate_kbps;  /* a legal rate of that layer / version */
    int nframes;
    int joint;         /* 1: joint stereo with a random bound (mode_extension) per frame; 0: plain stereo */
    int crc;           /* 1: 16-bit CRC word present (never verified by the reference) */
    int padding;       /* 1: set the padding bit on alternating frames */
    int fill;          /* 0..100: how much of every frame's bit budget the allocation tries to use */
    int ref_syntax;    /* Layer II: 1 = write the stream the way the D reference READS it: minimp3.d:417-421 evaluates get_bits(2) for
                        * all 2 x sblimit band-channel entries, also where the allocation is zero (the upstream C reads scfsi only for
                        * allocated entries, like ISO 11172-3).  0 = ISO syntax: the reference then mis-reads every frame that has an
                        * unallocated entry -- deterministically, which is all a drop-in has to reproduce. */
    int all_alloc;     /* 1: every band of every channel gets a non-zero allocation (with plain stereo the two syntaxes then coincide) */
} l12s_params_t;

typedef struct { uint64_t s; } rng_t;
static uint64_t rng_next(rng_t* r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t rng_below(rng_t* r, uint32_t n) { return n ? (uint32_t)((rng_next(r) >> 32) * (uint64_t)n >> 32) : 0; }

typedef struct { uint8_t* p; size_t cap; size_t bits; } bitwr_t;
static void bw_put(bitwr_t* w, uint32_t v, int n)
{
    for (int i = n - 1; i >= 0; i--) {
        size_t byte = w->bits >> 3;
        if (byte < w->cap && ((v >> i) & 1)) w->p[byte] |= (uint8_t)(0x80 >> (w->bits & 7));
        w->bits++;
    }
}

/* bit-allocation code tables: index written in the stream -> "ba" (0 none, 2..16 plain bits, 17/18/19 grouped 3/5/9 levels) */
static const uint8_t kCodeTab[] = {
    0, 17, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,
    0, 17, 18, 3, 19, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16,
    0, 17, 18, 3, 19, 4, 5, 16,
    0, 17, 18, 16,
    0, 17, 18, 19, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    0, 17, 18, 3, 19, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14,
    0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16
};
typedef struct { uint8_t tab_offset, width, count; } alloc_t;
static const alloc_t kL1[] = {{76, 4, 32}};
static const alloc_t kL2M2[] = {{60, 4, 4}, {44, 3, 7}, {44, 2, 19}};
static const alloc_t kL2M1[] = {{0, 4, 3}, {16, 4, 8}, {32, 3, 12}, {40, 2, 7}};
static const alloc_t kL2M1low[] = {{44, 4, 2}, {44, 3, 10}};

static const int kRateL1[15] = {0, 32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 352, 384, 416, 448};
static const int kRateL2[15] = {0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384};
static const int kRateLsfL1[15] = {0, 32, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256};
static const int kRateLsfL2[15] = {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160};

size_t l12s_max_bytes(const l12s_params_t* p)
{
    if (!p || p->nframes <= 0 || p->hz <= 0) return 0;
    return (size_t)p->nframes * 2048 + 64;
}

/* Returns bytes written, < 0 on bad parameters. */
long long l12s_generate(const l12s_params_t* p, uint8_t* out, size_t cap)
{
    if (!p || (p->layer != 1 && p->layer != 2) || (p->nch != 1 && p->nch != 2)) return -1;
    const int mpeg1 = p->hz >= 32000;
    static const int hz1[3] = {44100, 48000, 32000};
    int sr = -1;
    for (int i = 0; i < 3; i++)
        if ((hz1[i] >> (mpeg1 ? 0 : 1)) == p->hz) sr = i;
    if (sr < 0) return -1;
    const int* rates = p->layer == 1 ? (mpeg1 ? kRateL1 : kRateLsfL1) : (mpeg1 ? kRateL2 : kRateLsfL2);
    int bri = -1;
    for (int i = 1; i < 15; i++)
        if (rates[i] == p->bitrate_kbps) bri = i;
    if (bri < 0) return -1;
    rng_t r = {p->seed * 0x2545F4914F6CDD1Dull + 12345u};
    size_t at = 0;
    for (int f = 0; f < p->nframes; f++) {
        const int pad = p->padding && (f & 1);
        const int samples = p->layer == 1 ? 384 : 1152;
        int fb = samples * p->bitrate_kbps * 125 / p->hz;
        if (p->layer == 1) fb &= ~3;
        fb += pad ? (p->layer == 1 ? 4 : 1) : 0;
        if (at + (size_t)fb > cap) return -3;
        uint8_t* h = out + at;
        memset(h, 0, (size_t)fb);
        const int mode = p->nch == 1 ? 3 : (p->joint ? 1 : 0);
        const int mode_ext = (int)rng_below(&r, 4);
        h[0] = 0xFF;
        h[1] = (uint8_t)(0xE0 | (mpeg1 ? 0x18 : 0x10) | ((4 - p->layer) << 1) | (p->crc ? 0 : 1));
        h[2] = (uint8_t)((bri << 4) | (sr << 2) | (pad ? 2 : 0));
        h[3] = (uint8_t)((mode << 6) | ((mode == 1 ? mode_ext : 0) << 4));
        bitwr_t w = {h + 4, (size_t)fb - 4, 0};
        if (p->crc) bw_put(&w, rng_below(&r, 65536), 16);
        /* the allocation table the reference will pick (minimp3.d:284-350) */
        const alloc_t* alloc;
        int nbands;
        if (p->layer == 1) { alloc = kL1; nbands = 32; }
        else if (!mpeg1) { alloc = kL2M2; nbands = 30; }
        else {
            int kbps = p->bitrate_kbps >> (mode != 3);
            alloc = kL2M1; nbands = 27;
            if (kbps < 56) { alloc = kL2M1low; nbands = sr == 2 ? 12 : 8; }
            else if (kbps >= 96 && sr != 1) nbands = 30;
        }
        int bound = mode == 3 ? 0 : (mode == 1 ? (mode_ext << 2) + 4 : 32);
        if (bound > nbands) bound = nbands;
        const int groups = p->layer == 1 ? 12 : 12;          /* sample groups per frame (1 or 3 samples each) */
        const int spg = p->layer == 1 ? 1 : 3;
        /* ---- choose allocations within the bit budget ---- */
        uint8_t idx[32][2], ba[32][2], scfsi[32][2];
        memset(idx, 0, sizeof idx); memset(ba, 0, sizeof ba); memset(scfsi, 0, sizeof scfsi);
        long budget = ((long)fb - 4) * 8 - (p->crc ? 16 : 0);
        budget = budget * (p->fill > 0 ? p->fill : 90) / 100;
        long used = (p->layer == 2 && p->ref_syntax) ? 2L * 2 * nbands : 0;   /* the reference reads a scfsi field for every entry */
        {   /* fixed cost: allocation fields (+ nothing else until a band is switched on) */
            int k = 0;
            const alloc_t* a = alloc;
            for (int sb = 0; sb < nbands; sb++) {
                if (sb == k) { k += a->count; a++; }
                used += (long)(a - 1)->width * (mode == 3 ? 1 : (sb < bound ? 2 : 1));
            }
        }
        for (int pass = p->all_alloc ? 0 : 1; pass < 2; pass++) {
            /* pass 0 (all_alloc): the cheapest allocation everywhere, so that every entry is non-zero; pass 1: random upgrades within the budget */
            int k = 0;
            const alloc_t* a = alloc;
            const alloc_t* cur = alloc;
            for (int sb = 0; sb < nbands; sb++) {
                if (sb == k) { cur = a; k += a->count; a++; }
                const int nchan_fields = mode == 3 ? 1 : (sb < bound ? 2 : 1);
                for (int c = 0; c < nchan_fields; c++) {
                    int span = 1 << cur->width;
                    if (p->layer == 1) span = 15;   /* index 15 is forbidden in Layer I */
                    int want = pass == 0 ? 1 : (int)rng_below(&r, (uint32_t)span);
                    if (pass == 1 && rng_below(&r, 32) < (uint32_t)sb / 2) want = 0;   /* quieter high bands */
                    const int chans = (mode != 3 && sb >= bound) ? 2 : 1;    /* a shared allocation carries scalefactors for both channels */
                    long have = 0;                                           /* what this entry costs already */
                    if (ba[sb][c]) {
                        int sel0 = p->layer == 1 ? 2 : scfsi[sb][c];
                        int nscf0 = sel0 == 0 ? 3 : (sel0 == 2 ? 1 : 2);
                        int b0 = ba[sb][c];
                        have = ((p->layer == 2 && !p->ref_syntax) ? 2 : 0) * chans + 6L * nscf0 * chans +
                               (b0 < 17 ? (long)b0 * spg : (b0 == 17 ? 5 : (b0 == 18 ? 7 : 10))) * groups;
                    }
                    for (; want > 0; want--) {
                        int b = kCodeTab[cur->tab_offset + want];
                        int sel = p->layer == 2 ? (pass == 0 ? 2 : (int)rng_below(&r, 4)) : 2;
                        int nscf = sel == 0 ? 3 : (sel == 2 ? 1 : 2);
                        long cost = ((p->layer == 2 && !p->ref_syntax) ? 2 : 0) * chans + 6L * nscf * chans;
                        cost += (b < 17 ? (long)b * spg : (b == 17 ? 5 : (b == 18 ? 7 : 10))) * groups;   /* samples are coded once for a shared band */
                        if (used - have + cost <= budget || pass == 0) {
                            idx[sb][c] = (uint8_t)want; ba[sb][c] = (uint8_t)b; scfsi[sb][c] = (uint8_t)sel;
                            used += cost - have;
                            break;
                        }
                    }
                    if (mode != 3 && sb >= bound) { idx[sb][1] = idx[sb][0]; ba[sb][1] = ba[sb][0]; scfsi[sb][1] = scfsi[sb][0]; }
                }
            }
        }
        /* ---- write: allocation ---- */
        {
            int k = 0;
            const alloc_t* a = alloc;
            const alloc_t* cur = alloc;
            for (int sb = 0; sb < nbands; sb++) {
                if (sb == k) { cur = a; k += a->count; a++; }
                bw_put(&w, idx[sb][0], cur->width);
                if (mode != 3 && sb < bound) bw_put(&w, idx[sb][1], cur->width);
            }
        }
        /* scfsi (Layer II): for every band / channel with an allocation */
        if (p->layer == 2)
            for (int sb = 0; sb < nbands; sb++)
                for (int c = 0; c < (p->ref_syntax ? 2 : p->nch); c++)
                    if (ba[sb][c] || p->ref_syntax) bw_put(&w, scfsi[sb][c], 2);
        /* scalefactors */
        for (int sb = 0; sb < nbands; sb++)
            for (int c = 0; c < p->nch; c++)
                if (ba[sb][c]) {
                    int sel = p->layer == 1 ? 2 : scfsi[sb][c];
                    int nscf = sel == 0 ? 3 : (sel == 2 ? 1 : 2);
                    for (int q = 0; q < nscf; q++) bw_put(&w, (p->layer == 2 ? 14u : 8u) + (uint32_t)sb / 2 + rng_below(&r, 12), 6);
                }
        /* samples */
        for (int g = 0; g < groups; g++)
            for (int sb = 0; sb < nbands; sb++)
                for (int c = 0; c < (mode == 3 ? 1 : (sb < bound ? 2 : 1)); c++) {
                    int b = ba[sb][c];
                    if (!b) continue;
                    if (b < 17) {
                        for (int q = 0; q < spg; q++) {
                            uint32_t v = rng_below(&r, (1u << b) - 1u);   /* the all-ones code is forbidden */
                            bw_put(&w, v, b);
                        }
                    } else {
                        int mod = (2 << (b - 17)) + 1, nb = b == 17 ? 5 : (b == 18 ? 7 : 10);
                        bw_put(&w, rng_below(&r, (uint32_t)(mod * mod * mod)), nb);
                    }
                }
        if ((long)w.bits > ((long)fb - 4) * 8) return -2;
        at += (size_t)fb;
    }
    return (long long)at;
}
