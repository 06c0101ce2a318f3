/* l3synth.c -- deterministic synthetic MPEG-1/2/2.5 Layer III bitstream generator.
 *
 * Writes LEGAL Layer III frames directly at the syntax level (no psychoacoustics, no encoder):
 * it chooses block types, Huffman tables, region splits, scalefactors and quantised values from a
 * seeded counter-free SplitMix64 stream, Huffman-ENCODES them with the canonical books of
 * l3_tables_gen.h, and packs the granules through a real bit reservoir (main_data_begin).
 * It also returns the signed integers it encoded so tests can check decoder spectra bit-exactly.
 *
 * It honours the decoder quirks listed in SURVEY.md 8c: private bits 0, mode_ext 0 unless joint
 * stereo, first frame main_data_begin 0, no stuffing inside part2_3_length, table_select never
 * 4/14, big_values <= 288, block_type != 0 when window switching, no trailing bytes.
 *
 * Bench/test infrastructure: the decoder product (csrc/) does not depend on this file.
 */
#include "l3synth.h"

#include <stdlib.h>
#include <string.h>

#include "../csrc/l3_tables_gen.h"

/* ---------------------------------------------------------------- rng */
typedef struct { uint64_t s; } rng_t;
static uint64_t rng_next(rng_t* r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t rng_below(rng_t* r, uint32_t n) { return n ? (uint32_t)((rng_next(r) >> 32) * (uint64_t)n >> 32) : 0; }
static int rng_chance(rng_t* r, uint32_t num, uint32_t den) { return rng_below(r, den) < num; }
static double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

/* ---------------------------------------------------------------- bit writer */
typedef struct {
    uint8_t* buf;
    size_t cap;   /* bytes */
    uint64_t pos; /* bits */
    int overflow;
} bitwr_t;

/* OR-only writer: every target buffer is zero-initialised and never rewound into written data */
static void bw_put(bitwr_t* w, uint32_t v, int n)
{
    if (n <= 0) return;
    uint64_t pos = w->pos;
    w->pos += (uint64_t)n;
    if (((pos + (uint64_t)n + 7) >> 3) > w->cap) { w->overflow = 1; return; }
    while (n > 0) {
        int bit = (int)(pos & 7), room = 8 - bit;
        int take = n < room ? n : room;
        uint32_t chunk = (v >> (n - take)) & ((1u << take) - 1u);
        w->buf[pos >> 3] |= (uint8_t)(chunk << (room - take));
        pos += (uint64_t)take;
        n -= take;
    }
}

/* ---------------------------------------------------------------- format helpers */
static const int k_rates_m1[15] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320};
static const int k_rates_m2[15] = {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160};

typedef struct {
    int mpeg1, mpeg25, sr_code /* header 2-bit */, sr_idx /* sfb table row */, br_idx;
    int side_bytes, ngr, max_mdb;
} fmt_t;

static int resolve_format(const l3s_params_t* p, fmt_t* f)
{
    static const int base[3] = {44100, 48000, 32000};
    int found = 0;
    for (int v = 0; v < 3 && !found; v++)
        for (int c = 0; c < 3; c++)
            if ((base[c] >> v) == p->hz) {
                f->mpeg1 = (v == 0);
                f->mpeg25 = (v == 2);
                f->sr_code = c;
                found = 1;
                break;
            }
    if (!found) return -1;
    /* sfb row = HDR_GET_MY_SAMPLE_RATE - (that != 0)   (minimp3.d:135-138, 523) */
    int my = f->sr_code + ((f->mpeg1 ? 1 : 0) + (f->mpeg25 ? 0 : 1)) * 3;
    f->sr_idx = my - (my != 0);
    const int* rates = f->mpeg1 ? k_rates_m1 : k_rates_m2;
    f->br_idx = 0;
    for (int i = 1; i < 15; i++)
        if (rates[i] == p->bitrate_kbps) f->br_idx = i;
    if (!f->br_idx && !p->free_format) return -1;
    if (p->free_format) f->br_idx = 0;   /* bitrate index 0: the decoder finds the frame size by searching for the next header */
    if (p->nch != 1 && p->nch != 2) return -1;
    f->side_bytes = f->mpeg1 ? (p->nch == 1 ? 17 : 32) : (p->nch == 1 ? 9 : 17);
    f->ngr = f->mpeg1 ? 2 : 1;
    f->max_mdb = f->mpeg1 ? 511 : 255;
    return 0;
}

/* ---------------------------------------------------------------- per granule-channel plan */
typedef struct {
    int part23, big_values, global_gain, scalefac_compress;
    int window_switching, block_type, mixed;
    int table_select[3], region0, region1, subblock_gain[3];
    int preflag, scalefac_scale, count1_table, scfsi;
    /* payload */
    uint8_t scf_bits_len[40];
    uint8_t scf_val[40];
    int n_scf;        /* transmitted scalefactors in order */
    int16_t is[576];
    int count1_end;   /* index just past the last coded count1 quad */
} grch_t;

/* highest value encodable by table_select t (with escape bits where the table has them) */
static int g_maxval[32];
static int g_maxval_ready = 0;
static int table_maxval(int t)
{
    if (!g_maxval_ready) {
        for (int q = 0; q < 32; q++) {
            int book = L3_SEL2BOOK[q], mx = 0;
            if (book >= 0) {
                for (int s = 0; s < 256; s++)
                    if (L3_HLEN[book * 256 + s] && (s >> 4) > mx) mx = s >> 4;
                if (L3_LINBITS[q]) mx = 15 + (1 << L3_LINBITS[q]) - 1;
            }
            g_maxval[q] = mx;
        }
        __atomic_store_n(&g_maxval_ready, 1, __ATOMIC_RELEASE);
    }
    return g_maxval[t];
}

static int pair_cost(int t, int a0, int a1)
{
    int book = L3_SEL2BOOK[t];
    if (book < 0) return (a0 | a1) ? 1 << 20 : 0;
    int lb = L3_LINBITS[t];
    int c0 = a0 > 15 ? 15 : a0, c1 = a1 > 15 ? 15 : a1;
    int len = L3_HLEN[book * 256 + c0 * 16 + c1];
    if (!len) return 1 << 20;
    if (lb) { if (c0 == 15) len += lb; if (c1 == 15) len += lb; }
    return len + (a0 != 0) + (a1 != 0);
}

static void pair_emit(bitwr_t* w, int t, int v0, int v1)
{
    int book = L3_SEL2BOOK[t];
    if (book < 0) return;
    int lb = L3_LINBITS[t];
    int a0 = abs(v0), a1 = abs(v1);
    int c0 = a0 > 15 ? 15 : a0, c1 = a1 > 15 ? 15 : a1;
    bw_put(w, L3_HCODE[book * 256 + c0 * 16 + c1], L3_HLEN[book * 256 + c0 * 16 + c1]);
    /* order in the stream (minimp3.d:805-820): [linbits0][sign0][linbits1][sign1] */
    if (lb && c0 == 15) bw_put(w, (uint32_t)(a0 - 15), lb);
    if (a0) bw_put(w, v0 < 0, 1);
    if (lb && c1 == 15) bw_put(w, (uint32_t)(a1 - 15), lb);
    if (a1) bw_put(w, v1 < 0, 1);
}

static const int k_pair_tables[] = {1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31};
#define N_PAIR_TABLES ((int)(sizeof k_pair_tables / sizeof k_pair_tables[0]))

static const uint8_t* sfb_table(const fmt_t* f, int block_type, int mixed)
{
    if (block_type == 2) return mixed ? L3_SFB_MIXED + f->sr_idx * 40 : L3_SFB_SHORT + f->sr_idx * 40;
    return L3_SFB_LONG + f->sr_idx * 23;
}

/* geometric-ish magnitude with mean `m`, capped at `cap` */
static int draw_mag(rng_t* r, double m, int cap)
{
    if (cap <= 0 || m <= 0) return 0;
    double u = rng_unit(r);
    /* P(X >= k) = q^k with mean q/(1-q) = m  */
    double q = m / (1.0 + m);
    int k = 0;
    double acc = q;
    while (u < acc && k < cap) { k++; acc *= q; }
    return k;
}

typedef struct {
    int bt_state[2]; /* per channel block-type state machine: 0 long, 1 start, 2 short, 3 stop */
    int short_left[2];
    int mixed_run[2];
} chan_state_t;

static int next_block_type(const l3s_params_t* p, chan_state_t* cs, rng_t* r, int ch, int* mixed)
{
    *mixed = 0;
    if (!p->block_mode) return 0;
    int st = cs->bt_state[ch], nt;
    switch (st) {
    case 0:
    default:
        nt = rng_chance(r, 1, st == 0 ? 6 : 8) ? 1 : 0;
        /* whether the coming short run is mixed is decided with its START block: an encoder that keeps the two
         * lowest subbands on long transforms signals mixed_block_flag on the start, short and stop blocks alike
         * (ISO 11172-3 2.4.2.7), and the reference honours the flag on all of them (minimp3.d:1212, 1158-1167) */
        if (nt == 1) cs->mixed_run[ch] = p->block_mode == 1 ? rng_chance(r, 1, 3) : 0;
        break;
    case 1:
        nt = 2;
        cs->short_left[ch] = 1 + (int)rng_below(r, 3);
        break;
    case 2:
        if (--cs->short_left[ch] > 0) nt = 2; else nt = 3;
        break;
    }
    cs->bt_state[ch] = nt;
    /* 8 kHz mixed blocks are not generated: the reference's L3_reorder starts 72 coefficients in but walks a band
     * table that assumes 48 (minimp3.d:1218, 1223 with the sfb row of 8 kHz) and runs past its 576-float buffer. */
    if (nt == 2 || (nt != 0 && !p->mixed_only_short)) *mixed = cs->mixed_run[ch] && p->hz != 8000;
    return nt;
}

/* Fill one granule-channel within `budget` bits. `force_zero_above` (>=0): coefficients at or above
 * that index must stay zero (used to give intensity stereo something to do on channel 1). */
static void plan_grch(const l3s_params_t* p, const fmt_t* f, rng_t* r, grch_t* g, int budget, int block_type, int mixed,
                      int gr_index, int ch, int istereo_ch1, int allow_scfsi, const grch_t* gr0, int force_zero_above,
                      int leaked_nibble)
{
    memset(g, 0, sizeof *g);
    if (budget > 4095) budget = 4095;
    g->block_type = block_type;
    g->mixed = mixed;
    g->window_switching = block_type != 0;
    g->scalefac_scale = (int)rng_below(r, 2);
    g->count1_table = (int)rng_below(r, 2);
    g->preflag = f->mpeg1 ? (int)rng_below(r, 2) : 0;
    int gain_jitter = (int)rng_below(r, 9) - 4;
    int esc_granule = p->escapes && rng_chance(r, 1, 3); /* escapes cluster in some granules, like loud passages */
    if (block_type == 2)
        for (int i = 0; i < 3; i++) g->subblock_gain[i] = rng_chance(r, 1, 2) ? (int)rng_below(r, 3) : 0;
    else if (g->window_switching)
        for (int i = 0; i < 3; i++) g->subblock_gain[i] = (int)rng_below(r, 8); /* ignored by decoders for non-short */

    /* ---- scalefactors ---- */
    const uint8_t* part = L3_SCF_PARTITIONS + 28 * ((block_type == 2) + (block_type == 2 && !mixed));
    uint8_t slen[4];
    int used = 0;
    if (f->mpeg1) {
        int sfc = (int)rng_below(r, 16);
        if (p->small_scalefactors) sfc = (int)rng_below(r, 4);
        int pp = L3_SCFC_DECODE[sfc];
        slen[0] = slen[1] = (uint8_t)(pp >> 2);
        slen[2] = slen[3] = (uint8_t)(pp & 3);
        g->scalefac_compress = sfc;
        g->scfsi = 0;
        /* granule 0: the reference reads the private bits together with scfsi and they end up as THIS granule-channel's
         * scfsi nibble (minimp3.d:530-540, 600-601); a flagged partition is then "copied" from the frame's zeroed scratch
         * and its bits are not read -- so they must not be written either.  Short blocks clear the nibble (:568). */
        if (gr_index == 0 && block_type != 2) g->scfsi = leaked_nibble;
        if (allow_scfsi && gr_index == 1 && block_type != 2 && gr0 && gr0->block_type != 2 && rng_chance(r, 1, 2))
            g->scfsi = (int)rng_below(r, 16);
    } else {
        int sfc;
        if (istereo_ch1) sfc = (int)rng_below(r, 512);
        else sfc = rng_chance(r, 1, 8) ? 500 + (int)rng_below(r, 12) : (int)rng_below(r, 500);
        if (p->small_scalefactors) sfc = istereo_ch1 ? (int)rng_below(r, 2 * 180) : (int)rng_below(r, 200);
        g->scalefac_compress = sfc;
        g->preflag = 0; /* derived by the decoder from scalefac_compress >= 500 */
        int ist = istereo_ch1 ? 1 : 0, k, modprod, s2 = sfc >> ist;
        for (k = ist * 12; s2 >= 0; s2 -= modprod, k += 4) {
            modprod = 1;
            for (int i = 3; i >= 0; i--) {
                slen[i] = (uint8_t)(s2 / modprod % L3_LSF_MOD[k + i]);
                modprod *= L3_LSF_MOD[k + i];
            }
        }
        part += k;
    }
    int scf_bits = 0, n = 0, sc = g->scfsi;
    for (int i = 0; i < 4 && part[i]; i++, sc *= 2) {
        for (int k = 0; k < part[i]; k++) {
            if ((sc & 8) || !slen[i]) continue;
            g->scf_bits_len[n] = slen[i];
            /* keep attenuation moderate: bias towards small values */
            uint32_t mx = 1u << slen[i];
            uint32_t v = rng_chance(r, 3, 4) ? rng_below(r, mx > 4 ? 4 : mx) : rng_below(r, mx);
            g->scf_val[n++] = (uint8_t)v;
            scf_bits += slen[i];
        }
    }
    g->n_scf = n;
    if (scf_bits > budget) { /* cannot afford: fall back to zero-length scalefactors */
        if (f->mpeg1) { g->scalefac_compress = 0; g->scfsi = 0; }
        else g->scalefac_compress = 0;
        g->n_scf = 0;
        scf_bits = 0;
    }
    used = scf_bits;

    /* ---- region split + tables ---- */
    const uint8_t* sfb = sfb_table(f, block_type, mixed);
    int r1_start, r2_start;
    if (g->window_switching) {
        g->region0 = (block_type == 2 && !mixed) ? 8 : 7;
        g->region1 = 255;
        int acc = 0;
        for (int i = 0; i <= g->region0; i++) acc += sfb[i];
        r1_start = acc;
        r2_start = 576;
    } else {
        g->region0 = (int)rng_below(r, 16);
        g->region1 = (int)rng_below(r, 8);
        int acc = 0, i = 0;
        for (; i <= g->region0 && sfb[i]; i++) acc += sfb[i];
        r1_start = acc;
        for (int j = 0; j <= g->region1 && sfb[i]; j++, i++) acc += sfb[i];
        r2_start = acc;
    }
    for (int i = 0; i < 3; i++) {
        if (p->table_cycle) g->table_select[i] = k_pair_tables[(p->table_cycle_pos + i * 7 + gr_index * 3 + ch * 11 + (int)rng_below(r, 3)) % N_PAIR_TABLES];
        else g->table_select[i] = k_pair_tables[rng_below(r, N_PAIR_TABLES)];
        if (!p->escapes && L3_LINBITS[g->table_select[i]]) g->table_select[i] = 13;
        if (rng_chance(r, 1, 40)) g->table_select[i] = 0; /* the all-zero book is legal */
    }
    if (g->window_switching) g->table_select[2] = 0; /* not transmitted */

    /* ---- big values ---- */
    int limit_idx = force_zero_above >= 0 ? force_zero_above : 576;
    int bv_budget_frac = 60 + (int)rng_below(r, 35); /* % of the remaining budget spent in the big_values region */
    int bv_budget = (budget - used) * bv_budget_frac / 100;
    int bv_spent = 0, pidx = 0;
    double tilt = 0.985 + 0.01 * rng_unit(r);
    double level = p->level * (0.5 + rng_unit(r));
    for (; pidx < 288 && 2 * pidx + 1 < limit_idx; pidx++) {
        int idx = 2 * pidx;
        int t = g->table_select[idx < r1_start ? 0 : idx < r2_start ? 1 : 2];
        int cap = table_maxval(t);
        double m = level;
        level *= tilt;
        int a0 = draw_mag(r, m, cap), a1 = draw_mag(r, m, cap);
        if (esc_granule && L3_LINBITS[t] && rng_chance(r, 1, 24)) a0 = 15 + (int)rng_below(r, 1u << L3_LINBITS[t]);
        if (esc_granule && L3_LINBITS[t] && rng_chance(r, 1, 200)) a1 = cap; /* hit the largest escape */
        int c = pair_cost(t, a0, a1);
        if (bv_spent + c > bv_budget) {
            /* try the cheapest pair (0,0) so a zero-book region does not end big_values prematurely */
            a0 = a1 = 0;
            c = pair_cost(t, 0, 0);
            if (bv_spent + c > bv_budget) break;
        }
        g->is[idx] = (int16_t)(rng_below(r, 2) ? -a0 : a0);
        g->is[idx + 1] = (int16_t)(rng_below(r, 2) ? -a1 : a1);
        bv_spent += c;
        if (rng_chance(r, 1, 400)) { pidx++; break; } /* occasional early end */
    }
    g->big_values = pidx;
    used += bv_spent;

    /* ---- count1 ---- */
    int idx = 2 * g->big_values;
    int c1b = g->count1_table ? 16 : 0;
    while (idx + 4 <= limit_idx + 0 && idx + 4 <= 576) {
        int flags = 0;
        for (int s = 0; s < 4; s++) flags = (flags << 1) | rng_chance(r, 2, 5);
        int c = L3_C1LEN[c1b + flags] + __builtin_popcount((unsigned)flags);
        if (used + c > budget) break;
        for (int s = 0; s < 4; s++)
            if (flags & (8 >> s)) g->is[idx + s] = (int16_t)(rng_below(r, 2) ? -1 : 1);
        used += c;
        idx += 4;
        if (rng_chance(r, 1, 60)) break;
    }
    g->count1_end = idx;
    g->part23 = used;

    /* global_gain follows the granule's largest magnitude M the way an encoder's step size follows the
     * signal level: xr_max = M^(4/3) * 2^((gg-210)/4) stays near 2^((gain_base-210)/4). */
    int maxmag = 1;
    for (int i = 0; i < 576; i++) { int a = abs(g->is[i]); if (a > maxmag) maxmag = a; }
    int drop = 0; /* floor(16/3*log2(M)) by integer search: smallest d with 2^(3(d+1)/16) > M  <=>  2^(3(d+1)) > M^16 */
    {
        /* compare in log domain with exact integer arithmetic on bit lengths is awkward; a small table of
         * thresholds T[d] = ceil(2^(3d/16)) generated by repeated multiplication in double is deterministic
         * enough (IEEE mul only). */
        double thr = 1.0;
        const double step = 1.1388381188899948; /* 2^(3/16) */
        while (drop < 80) { thr *= step; if (thr > (double)maxmag) break; drop++; }
    }
    g->global_gain = p->gain_base - drop + gain_jitter;
    if (g->global_gain < 0) g->global_gain = 0;
    if (g->global_gain > 255) g->global_gain = 255;
}

static void emit_grch(bitwr_t* w, const grch_t* g, const fmt_t* f)
{
    (void)f;
    for (int i = 0; i < g->n_scf; i++) bw_put(w, g->scf_val[i], g->scf_bits_len[i]);
    const uint8_t* sfb = sfb_table(f, g->block_type, g->mixed);
    int r1_start, r2_start, acc = 0, i = 0;
    if (g->window_switching) {
        for (; i <= g->region0; i++) acc += sfb[i];
        r1_start = acc; r2_start = 576;
    } else {
        for (; i <= g->region0 && sfb[i]; i++) acc += sfb[i];
        r1_start = acc;
        for (int j = 0; j <= g->region1 && sfb[i]; j++, i++) acc += sfb[i];
        r2_start = acc;
    }
    for (int pidx = 0; pidx < g->big_values; pidx++) {
        int idx = 2 * pidx;
        int t = g->table_select[idx < r1_start ? 0 : idx < r2_start ? 1 : 2];
        pair_emit(w, t, g->is[idx], g->is[idx + 1]);
    }
    int c1b = g->count1_table ? 16 : 0;
    for (int idx = 2 * g->big_values; idx < g->count1_end; idx += 4) {
        int flags = 0;
        for (int s = 0; s < 4; s++) flags = (flags << 1) | (g->is[idx + s] != 0);
        bw_put(w, L3_C1CODE[c1b + flags], L3_C1LEN[c1b + flags]);
        for (int s = 0; s < 4; s++)
            if (g->is[idx + s]) bw_put(w, g->is[idx + s] < 0, 1);
    }
}

/* ---------------------------------------------------------------- side info + header */
static void write_side_info(bitwr_t* w, const fmt_t* f, int nch, int mdb, const grch_t* g /* [ngr][nch] */, int private_bits)
{
    if (f->mpeg1) {
        bw_put(w, (uint32_t)mdb, 9);
        bw_put(w, (uint32_t)private_bits, nch == 1 ? 5 : 3); /* they leak into granule 0's scfsi in the reference (SURVEY 8c quirk i) */
        for (int ch = 0; ch < nch; ch++) bw_put(w, (uint32_t)g[1 * nch + ch].scfsi, 4);
    } else {
        bw_put(w, (uint32_t)mdb, 8);
        bw_put(w, (uint32_t)private_bits & (nch == 1 ? 1u : 3u), nch == 1 ? 1 : 2); /* read and dropped by the reference (minimp3.d:534) */
    }
    for (int gr = 0; gr < f->ngr; gr++)
        for (int ch = 0; ch < nch; ch++) {
            const grch_t* q = &g[gr * nch + ch];
            bw_put(w, (uint32_t)q->part23, 12);
            bw_put(w, (uint32_t)q->big_values, 9);
            bw_put(w, (uint32_t)q->global_gain, 8);
            bw_put(w, (uint32_t)q->scalefac_compress, f->mpeg1 ? 4 : 9);
            bw_put(w, (uint32_t)q->window_switching, 1);
            if (q->window_switching) {
                bw_put(w, (uint32_t)q->block_type, 2);
                bw_put(w, (uint32_t)q->mixed, 1);
                bw_put(w, (uint32_t)q->table_select[0], 5);
                bw_put(w, (uint32_t)q->table_select[1], 5);
                for (int i = 0; i < 3; i++) bw_put(w, (uint32_t)q->subblock_gain[i], 3);
            } else {
                for (int i = 0; i < 3; i++) bw_put(w, (uint32_t)q->table_select[i], 5);
                bw_put(w, (uint32_t)q->region0, 4);
                bw_put(w, (uint32_t)q->region1, 3);
            }
            if (f->mpeg1) bw_put(w, (uint32_t)q->preflag, 1);
            bw_put(w, (uint32_t)q->scalefac_scale, 1);
            bw_put(w, (uint32_t)q->count1_table, 1);
        }
}

size_t l3s_max_bytes(const l3s_params_t* p)
{
    fmt_t f;
    if (resolve_format(p, &f)) return 0;
    const int top = p->vbr ? (f.mpeg1 ? 320 : 160) : p->bitrate_kbps;
    size_t per = (size_t)((f.mpeg1 ? 144000 : 72000) * top / p->hz) + 2;
    return per * (size_t)p->nframes + 8192 + (size_t)p->id3v2_bytes + (p->id3v1 ? 128 : 0);
}

long long l3s_generate(const l3s_params_t* p, uint8_t* out, size_t cap, int16_t* is_out, l3s_info_t* info_out)
{
    fmt_t f;
    if (resolve_format(p, &f) || p->nframes < 1) return -1;
    const int nch = p->nch, ngr = f.ngr;
    rng_t rng = {p->seed * 0xD1342543DE82EF95ull + 0x1234567};
    chan_state_t cs;
    memset(&cs, 0, sizeof cs);

    /* frame geometry */
    const int spf = f.mpeg1 ? 1152 : 576;
    const long long num = (long long)spf / 8 * p->bitrate_kbps * 1000; /* bytes*hz per frame */
    size_t nfr = (size_t)p->nframes;
    int* fbytes = (int*)malloc(nfr * sizeof(int));
    int* slot = (int*)malloc(nfr * sizeof(int));
    uint8_t* bri = (uint8_t*)malloc(nfr);   /* bitrate index and padding bit of every frame */
    uint8_t* padb = (uint8_t*)malloc(nfr);
    rng_t vrng = {p->seed * 0x9E3779B97F4A7C15ull + 0x7654321};   /* its own stream: vbr = 0 streams are unchanged */
    long long rem = 0;
    size_t total_slots = 0;
    for (size_t i = 0; i < nfr; i++) {
        int base = (int)(num / p->hz);
        rem += num % p->hz;
        int pad = 0;
        if (rem >= p->hz) { rem -= p->hz; pad = 1; }
        if (p->no_padding) pad = 0;
        bri[i] = (uint8_t)f.br_idx;
        if (p->vbr && !p->free_format) {   /* the bitrate index (and padding bit) may change from frame to frame (minimp3.d:241-247) */
            const int* rates = f.mpeg1 ? k_rates_m1 : k_rates_m2;
            int lo = f.br_idx - 4 < 1 ? 1 : f.br_idx - 4, hi = f.br_idx + 3 > 14 ? 14 : f.br_idx + 3;
            bri[i] = (uint8_t)(lo + (int)rng_below(&vrng, (uint32_t)(hi - lo + 1)));
            base = (int)((long long)spf / 8 * rates[bri[i]] * 1000 / p->hz);
            pad = p->no_padding ? 0 : (int)rng_below(&vrng, 2);
        }
        padb[i] = (uint8_t)pad;
        fbytes[i] = base + pad;
        slot[i] = fbytes[i] - 4 - (p->crc ? 2 : 0) - f.side_bytes;
        if (slot[i] < 0) { free(fbytes); free(slot); free(bri); free(padb); return -1; }
        total_slots += (size_t)slot[i];
    }
    /* main-data stream (all slots concatenated) */
    uint8_t* md = (uint8_t*)calloc(total_slots + 64, 1);
    bitwr_t mw = {md, total_slots + 64, 0, 0};
    int* mdb_of = (int*)malloc(nfr * sizeof(int));
    grch_t* plans = (grch_t*)malloc(nfr * (size_t)(ngr * nch) * sizeof(grch_t));
    uint8_t* hdr3 = (uint8_t*)malloc(nfr);
    uint8_t* priv = (uint8_t*)calloc(nfr, 1);

    size_t slot_start = 0;
    long long granules_out = 0;
    for (size_t fi = 0; fi < nfr; fi++) {
        size_t prev_end = (size_t)((mw.pos + 7) >> 3);
        size_t begin = prev_end;
        size_t earliest = slot_start > (size_t)f.max_mdb ? slot_start - (size_t)f.max_mdb : 0;
        if (p->reservoir == 0) earliest = slot_start;
        if (begin < earliest) begin = earliest;        /* stuffing (ancillary) between frames */
        if (begin > slot_start) begin = slot_start;    /* cannot happen: data never ends past its slot */
        mw.pos = (uint64_t)begin * 8;
        int mdb = (int)(slot_start - begin);
        mdb_of[fi] = mdb;
        int avail = (mdb + slot[fi]) * 8;

        /* stereo mode for this frame */
        int mode = nch == 1 ? 3 : 0, mode_ext = 0;
        if (nch == 2 && p->stereo_mode) {
            if (rng_chance(&rng, 3, 4)) {
                mode = 1;
                int ms = rng_chance(&rng, 1, 2);
                int is = p->stereo_mode >= 2 ? rng_chance(&rng, 1, 3) : 0;
                mode_ext = (ms ? 2 : 0) | (is ? 1 : 0);
            }
        }
        /* mode_extension is "don't care" outside joint stereo, but the reference tests its bits in every mode
         * (HDR_TEST_I_STEREO / HDR_TEST_MS_STEREO, minimp3.d:100-103); for stereo only with tied block types (see below) */
        if (p->mode_ext_any && mode != 1 && (nch == 1 || (p->stereo_mode >= 2 && !p->istereo_untied))) mode_ext = (int)rng_below(&vrng, 4);
        hdr3[fi] = (uint8_t)((mode << 6) | (mode_ext << 4) | (p->emphasis_bits & 0xF));
        if (p->private_bits) priv[fi] = (uint8_t)rng_below(&rng, nch == 1 ? 32 : 8);

        /* how much of what is available this frame uses */
        double use;
        switch (p->reservoir) {
        case 0: use = 0.80 + 0.20 * rng_unit(&rng); break;
        case 1: use = 0.55 + 0.45 * rng_unit(&rng); break;
        default: use = (fi & 1) ? 1.0 : 0.25 + 0.2 * rng_unit(&rng); break; /* sparse / dense alternation */
        }
        int target = (int)(avail * use);
        if (target > avail) target = avail;
        int parts = ngr * nch;
        int left = target;
        grch_t* G = plans + fi * (size_t)parts;
        for (int gr = 0; gr < ngr; gr++) {
            int zero_above[2] = {-1, -1};
            int bt[2], mx[2];
            for (int ch = 0; ch < nch; ch++) bt[ch] = next_block_type(p, &cs, &rng, ch, &mx[ch]);
            if (nch == 2 && p->stereo_mode >= 2 && !p->istereo_untied) {
                /* With intensity stereo the reference walks channel 0's band layout over channel 1's ist_pos
                 * array (minimp3.d:963-981); if channel 0 has more bands than channel 1 transmitted it reads
                 * uninitialised scratch (UB upstream).  Encoders that use intensity stereo keep both channels
                 * on the same block type, and so do we. */
                bt[1] = bt[0];
                mx[1] = mx[0];
                cs.bt_state[1] = cs.bt_state[0];
                cs.short_left[1] = cs.short_left[0];
                cs.mixed_run[1] = cs.mixed_run[0];
            }
            if (nch == 2 && (mode_ext & 1)) {
                /* intensity stereo needs an all-zero top in channel 1 */
                zero_above[1] = 2 * (int)(40 + rng_below(&rng, 200));
            }
            for (int ch = 0; ch < nch; ch++) {
                int k = gr * nch + ch;
                int remaining_parts = parts - k;
                int share = left / remaining_parts;
                int budget = (int)(share * (0.6 + 0.8 * rng_unit(&rng)));
                if (budget > left) budget = left;
                if (remaining_parts == 1) budget = left;
                /* MPEG-1: stereo -> the three private bits become granule 0 / channel 1's scfsi nibble; mono -> the first of
                 * the five private bits becomes granule 0's lowest scfsi bit */
                const int leaked = !f.mpeg1 ? 0 : (nch == 2 ? (ch == 1 ? priv[fi] & 7 : 0) : (priv[fi] >> 4) & 1);
                plan_grch(p, &f, &rng, &G[k], budget, bt[ch], mx[ch], gr, ch, (mode_ext & 1) && ch == 1, p->scfsi,
                          gr == 1 ? &G[ch] : NULL, zero_above[ch], leaked);
                left -= G[k].part23;
                emit_grch(&mw, &G[k], &f);
                if (is_out) memcpy(is_out + (granules_out * nch + ch) * 576, G[k].is, 576 * sizeof(int16_t));
            }
            granules_out++;
        }
        slot_start += (size_t)slot[fi];
        if (((mw.pos + 7) >> 3) > slot_start) { /* internal error: overran the slot */
            free(fbytes); free(slot); free(md); free(mdb_of); free(plans); free(hdr3); free(priv); free(bri); free(padb);
            return -2;
        }
    }

    /* ---- assemble the file ---- */
    size_t o = 0;
    size_t need = (size_t)p->id3v2_bytes + (p->id3v1 ? 128 : 0);
    for (size_t i = 0; i < nfr; i++) need += (size_t)fbytes[i];
    if (need > cap) { free(fbytes); free(slot); free(md); free(mdb_of); free(plans); free(hdr3); free(priv); free(bri); free(padb); return -3; }
    memset(out, 0, need);
    if (p->id3v2_bytes >= 10) {
        int body = p->id3v2_bytes - 10;
        out[0] = 'I'; out[1] = 'D'; out[2] = '3'; out[3] = 4; out[4] = 0; out[5] = 0;
        out[6] = (uint8_t)((body >> 21) & 0x7f); out[7] = (uint8_t)((body >> 14) & 0x7f);
        out[8] = (uint8_t)((body >> 7) & 0x7f); out[9] = (uint8_t)(body & 0x7f);
        o = (size_t)p->id3v2_bytes;
    }
    size_t sp = 0;
    for (size_t fi = 0; fi < nfr; fi++) {
        uint8_t* h = out + o;
        h[0] = 0xFF;
        h[1] = (uint8_t)((f.mpeg25 ? 0xE0 : 0xF0) | (f.mpeg1 ? 0x08 : 0x00) | 0x02 | (p->crc ? 0 : 1));
        h[2] = (uint8_t)((bri[fi] << 4) | (f.sr_code << 2) | (padb[fi] << 1));
        h[3] = hdr3[fi];
        size_t q = 4;
        if (p->crc) { h[4] = 0xAB; h[5] = 0xCD; q = 6; } /* never verified by the reference (minimp3.d:1533-1536) */
        bitwr_t sw = {h + q, (size_t)f.side_bytes, 0, 0};
        write_side_info(&sw, &f, nch, mdb_of[fi], plans + fi * (size_t)(ngr * nch), priv[fi]);
        memcpy(h + q + f.side_bytes, md + sp, (size_t)slot[fi]);
        sp += (size_t)slot[fi];
        o += (size_t)fbytes[fi];
    }
    if (p->id3v1) {
        memcpy(out + o, "TAG", 3);
        o += 128;
    }
    if (info_out) {
        info_out->frames = (int)nfr;
        info_out->granules = (int)granules_out;
        info_out->bytes = (long long)o;
        info_out->samples_per_frame = spf;
        info_out->mpeg1 = f.mpeg1;
        info_out->sr_idx = f.sr_idx;
    }
    free(fbytes); free(slot); free(md); free(mdb_of); free(plans); free(hdr3); free(priv); free(bri); free(padb);
    return (long long)o;
}
