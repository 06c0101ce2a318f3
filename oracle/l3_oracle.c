/* l3_oracle.c -- CPU ORACLE (test infrastructure only; see l3_oracle.h).
 *
 * Restates the Layer III branch of /root/reference/source/audioformats/minimp3.d.  Float
 * arithmetic keeps the reference's operation order; build with -O2 -ffp-contract=off
 * -fno-fast-math so every a*b+c rounds twice like the un-fused D code on x86-64 SSE2.
 * PARITY UNPINNED against a run of the D reference (no D toolchain, no golden vectors upstream).
 */
#include "l3_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../audio_formats_b200/csrc/l3_tables_gen.h"

/* ------------------------------------------------------------------------------------------ */
/* constants (minimp3.d:53-63, 150-152)                                                        */
enum {
    MAX_FREE_FORMAT_FRAME_SIZE = 2304,
    MAX_FRAME_SYNC_MATCHES = 10,
    MAX_L3_FRAME_PAYLOAD_BYTES = 2304,
    MAX_BITRESERVOIR_BYTES = 511,
    SHORT_BLOCK_TYPE = 2,
    STOP_BLOCK_TYPE = 3,
    HDR_SIZE = 4,
    BITS_DEQUANTIZER_OUT = -1,
    MAX_SCF = 255 + BITS_DEQUANTIZER_OUT * 4 - 210,
    MAX_SCFI = (MAX_SCF + 3) & ~3
};

/* header predicates (minimp3.d:65-148) */
#define H_IS_MONO(h) (((h)[3] & 0xC0) == 0xC0)
#define H_IS_MS_STEREO(h) (((h)[3] & 0xE0) == 0x60)
#define H_IS_FREE_FORMAT(h) (((h)[2] & 0xF0) == 0)
#define H_IS_CRC(h) (!((h)[1] & 1))
#define H_TEST_PADDING(h) ((h)[2] & 0x2)
#define H_TEST_MPEG1(h) ((h)[1] & 0x8)
#define H_TEST_NOT_MPEG25(h) ((h)[1] & 0x10)
#define H_TEST_I_STEREO(h) ((h)[3] & 0x10)
#define H_TEST_MS_STEREO(h) ((h)[3] & 0x20)
#define H_GET_LAYER(h) (((h)[1] >> 1) & 3)
#define H_GET_BITRATE(h) ((h)[2] >> 4)
#define H_GET_SAMPLE_RATE(h) (((h)[2] >> 2) & 3)
#define H_GET_MY_SAMPLE_RATE(h) (H_GET_SAMPLE_RATE(h) + ((((h)[1] >> 3) & 1) + (((h)[1] >> 4) & 1)) * 3)
#define H_IS_FRAME_576(h) (((h)[1] & 14) == 2)
#define H_IS_LAYER_1(h) (((h)[1] & 6) == 6)

static int imin(int a, int b) { return a > b ? b : a; }
static int imax(int a, int b) { return a < b ? b : a; }

/* ------------------------------------------------------------------------------------------ */
/* taps + timers                                                                               */
static __thread l3o_tap_t* t_tap = NULL;
void l3o_set_tap(l3o_tap_t* tap) { t_tap = tap; }

static int g_timers_on = 0;
static __thread double t_timer[7];
void l3o_enable_timers(int on) { g_timers_on = on; memset(t_timer, 0, sizeof t_timer); }
void l3o_get_timers(double out[7]) { memcpy(out, t_timer, sizeof t_timer); }
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
#define TIMED(slot, stmt)                          \
    do {                                           \
        if (g_timers_on) {                         \
            double t0_ = now_s();                  \
            stmt;                                  \
            t_timer[slot] += now_s() - t0_;        \
        } else {                                   \
            stmt;                                  \
        }                                          \
    } while (0)

/* ------------------------------------------------------------------------------------------ */
/* MSB-first bit reader used for side info and scalefactors (minimp3.d:169-173, 209-230):
 * a read that would pass `limit` still advances `pos` but yields 0.                            */
typedef struct {
    const uint8_t* buf;
    int pos, limit;
} bitrd_t;

static void rd_init(bitrd_t* b, const uint8_t* data, int bytes)
{
    b->buf = data;
    b->pos = 0;
    b->limit = bytes * 8;
}

static uint32_t rd_bits(bitrd_t* b, int n)
{
    int p = b->pos;
    b->pos += n;
    if (b->pos > b->limit) return 0;
    const uint8_t* q = b->buf + (p >> 3);
    int skip = p & 7;
    int nbytes = (skip + n + 7) >> 3;
    uint64_t w = 0;
    for (int i = 0; i < nbytes; i++) w = (w << 8) | q[i];
    w >>= nbytes * 8 - skip - n;
    return (uint32_t)(w & ((1ull << n) - 1));
}

/* ------------------------------------------------------------------------------------------ */
/* header helpers (minimp3.d:232-283)                                                          */
int l3o_hdr_valid(const uint8_t* h)
{
    return h[0] == 0xff && ((h[1] & 0xF0) == 0xf0 || (h[1] & 0xFE) == 0xe2) && (H_GET_LAYER(h) != 0) &&
           (H_GET_BITRATE(h) != 15) && (H_GET_SAMPLE_RATE(h) != 3);
}

static int hdr_compare(const uint8_t* h1, const uint8_t* h2)
{
    return l3o_hdr_valid(h2) && ((h1[1] ^ h2[1]) & 0xFE) == 0 && ((h1[2] ^ h2[2]) & 0x0C) == 0 &&
           !(H_IS_FREE_FORMAT(h1) ^ H_IS_FREE_FORMAT(h2));
}

unsigned l3o_hdr_bitrate_kbps(const uint8_t* h)
{
    return 2u * L3_HALFRATE[((!!H_TEST_MPEG1(h)) * 3 + (H_GET_LAYER(h) - 1)) * 15 + H_GET_BITRATE(h)];
}

unsigned l3o_hdr_sample_rate_hz(const uint8_t* h)
{
    static const unsigned hz[3] = {44100, 48000, 32000};
    return hz[H_GET_SAMPLE_RATE(h)] >> (int)!H_TEST_MPEG1(h) >> (int)!H_TEST_NOT_MPEG25(h);
}

unsigned l3o_hdr_frame_samples(const uint8_t* h)
{
    return H_IS_LAYER_1(h) ? 384 : (1152 >> (int)H_IS_FRAME_576(h));
}

int l3o_hdr_frame_bytes(const uint8_t* h, int free_format_size)
{
    int frame_bytes = l3o_hdr_frame_samples(h) * l3o_hdr_bitrate_kbps(h) * 125 / l3o_hdr_sample_rate_hz(h);
    if (H_IS_LAYER_1(h)) frame_bytes &= ~3;
    return frame_bytes ? frame_bytes : free_format_size;
}

int l3o_hdr_padding(const uint8_t* h) { return H_TEST_PADDING(h) ? (H_IS_LAYER_1(h) ? 4 : 1) : 0; }

/* ------------------------------------------------------------------------------------------ */
/* side info (minimp3.d:189-196, 487-611)                                                      */
typedef struct {
    const uint8_t* sfbtab;
    uint16_t part_23_length, big_values, scalefac_compress;
    uint8_t global_gain, block_type, mixed_block_flag, n_long_sfb, n_short_sfb;
    uint8_t table_select[3], region_count[3], subblock_gain[3];
    uint8_t preflag, scalefac_scale, count1_table, scfsi;
} gr_info_t;

static int read_side_info(bitrd_t* bs, gr_info_t* gr, const uint8_t* hdr)
{
    unsigned tables, scfsi = 0;
    int main_data_begin, part_23_sum = 0;
    int sr_idx = H_GET_MY_SAMPLE_RATE(hdr);
    sr_idx -= (sr_idx != 0);
    int gr_count = H_IS_MONO(hdr) ? 1 : 2;

    if (H_TEST_MPEG1(hdr)) {
        gr_count *= 2;
        main_data_begin = rd_bits(bs, 9);
        scfsi = rd_bits(bs, 7 + gr_count);
    } else {
        main_data_begin = rd_bits(bs, 8 + gr_count) >> gr_count;
    }

    do {
        if (H_IS_MONO(hdr)) scfsi <<= 4;
        gr->part_23_length = (uint16_t)rd_bits(bs, 12);
        part_23_sum += gr->part_23_length;
        gr->big_values = (uint16_t)rd_bits(bs, 9);
        if (gr->big_values > 288) return -1;
        gr->global_gain = (uint8_t)rd_bits(bs, 8);
        gr->scalefac_compress = (uint16_t)rd_bits(bs, H_TEST_MPEG1(hdr) ? 4 : 9);
        gr->sfbtab = L3_SFB_LONG + sr_idx * 23;
        gr->n_long_sfb = 22;
        gr->n_short_sfb = 0;
        if (rd_bits(bs, 1)) {
            gr->block_type = (uint8_t)rd_bits(bs, 2);
            if (!gr->block_type) return -1;
            gr->mixed_block_flag = (uint8_t)rd_bits(bs, 1);
            gr->region_count[0] = 7;
            gr->region_count[1] = 255;
            if (gr->block_type == SHORT_BLOCK_TYPE) {
                scfsi &= 0x0F0F;
                if (!gr->mixed_block_flag) {
                    gr->region_count[0] = 8;
                    gr->sfbtab = L3_SFB_SHORT + sr_idx * 40;
                    gr->n_long_sfb = 0;
                    gr->n_short_sfb = 39;
                } else {
                    gr->sfbtab = L3_SFB_MIXED + sr_idx * 40;
                    gr->n_long_sfb = H_TEST_MPEG1(hdr) ? 8 : 6;
                    gr->n_short_sfb = 30;
                }
            }
            tables = rd_bits(bs, 10);
            tables <<= 5;
            gr->subblock_gain[0] = (uint8_t)rd_bits(bs, 3);
            gr->subblock_gain[1] = (uint8_t)rd_bits(bs, 3);
            gr->subblock_gain[2] = (uint8_t)rd_bits(bs, 3);
        } else {
            gr->block_type = 0;
            gr->mixed_block_flag = 0;
            tables = rd_bits(bs, 15);
            gr->region_count[0] = (uint8_t)rd_bits(bs, 4);
            gr->region_count[1] = (uint8_t)rd_bits(bs, 3);
            gr->region_count[2] = 255;
        }
        gr->table_select[0] = (uint8_t)(tables >> 10);
        gr->table_select[1] = (uint8_t)((tables >> 5) & 31);
        gr->table_select[2] = (uint8_t)(tables & 31);
        gr->preflag = H_TEST_MPEG1(hdr) ? (uint8_t)rd_bits(bs, 1) : (uint8_t)(gr->scalefac_compress >= 500);
        gr->scalefac_scale = (uint8_t)rd_bits(bs, 1);
        gr->count1_table = (uint8_t)rd_bits(bs, 1);
        gr->scfsi = (uint8_t)((scfsi >> 12) & 15);
        scfsi <<= 4;
        gr++;
    } while (--gr_count);

    if (part_23_sum + bs->pos > bs->limit + main_data_begin * 8) return -1;
    return main_data_begin;
}

/* ------------------------------------------------------------------------------------------ */
/* scalefactors (minimp3.d:613-720)                                                            */
static void read_scalefactors(uint8_t* scf, uint8_t* ist_pos, const uint8_t* scf_size, const uint8_t* scf_count,
                              bitrd_t* bitbuf, int scfsi)
{
    for (int i = 0; i < 4 && scf_count[i]; i++, scfsi *= 2) {
        int cnt = scf_count[i];
        if (scfsi & 8) {
            memcpy(scf, ist_pos, cnt);
        } else {
            int bits = scf_size[i];
            if (!bits) {
                memset(scf, 0, cnt);
                memset(ist_pos, 0, cnt);
            } else {
                int max_scf = (scfsi < 0) ? (1 << bits) - 1 : -1;
                for (int k = 0; k < cnt; k++) {
                    int s = (int)rd_bits(bitbuf, bits);
                    ist_pos[k] = (uint8_t)(s == max_scf ? -1 : s);
                    scf[k] = (uint8_t)s;
                }
            }
        }
        ist_pos += cnt;
        scf += cnt;
    }
    scf[0] = scf[1] = scf[2] = 0;
}

static float ldexp_q2(float y, int exp_q2)
{
    int e;
    do {
        e = imin(30 * 4, exp_q2);
        y *= L3_EXPFRAC[e & 3] * (float)(1 << 30 >> (e >> 2));
    } while ((exp_q2 -= e) > 0);
    return y;
}

static void decode_scalefactors(const uint8_t* hdr, uint8_t* ist_pos, bitrd_t* bs, const gr_info_t* gr, float* scf,
                                int ch, l3o_granule_tap_t* tap)
{
    const uint8_t* scf_partition = L3_SCF_PARTITIONS + 28 * (!!gr->n_short_sfb + !gr->n_long_sfb);
    uint8_t scf_size[4];
    uint8_t iscf[40];
    int i, scf_shift = gr->scalefac_scale + 1, gain_exp, scfsi = gr->scfsi;
    float gain;

    if (H_TEST_MPEG1(hdr)) {
        int part = L3_SCFC_DECODE[gr->scalefac_compress];
        scf_size[1] = scf_size[0] = (uint8_t)(part >> 2);
        scf_size[3] = scf_size[2] = (uint8_t)(part & 3);
    } else {
        int k, modprod, sfc, ist = H_TEST_I_STEREO(hdr) && ch;
        sfc = gr->scalefac_compress >> ist;
        for (k = ist * 3 * 4; sfc >= 0; sfc -= modprod, k += 4) {
            for (modprod = 1, i = 3; i >= 0; i--) {
                scf_size[i] = (uint8_t)(sfc / modprod % L3_LSF_MOD[k + i]);
                modprod *= L3_LSF_MOD[k + i];
            }
        }
        scf_partition += k;
        scfsi = -16;
    }
    read_scalefactors(iscf, ist_pos, scf_size, scf_partition, bs, scfsi);

    if (gr->n_short_sfb) {
        int sh = 3 - scf_shift;
        for (i = 0; i < gr->n_short_sfb; i += 3) {
            iscf[gr->n_long_sfb + i + 0] += gr->subblock_gain[0] << sh;
            iscf[gr->n_long_sfb + i + 1] += gr->subblock_gain[1] << sh;
            iscf[gr->n_long_sfb + i + 2] += gr->subblock_gain[2] << sh;
        }
    } else if (gr->preflag) {
        for (i = 0; i < 10; i++) iscf[11 + i] += L3_PREAMP[i];
    }

    gain_exp = gr->global_gain + BITS_DEQUANTIZER_OUT * 4 - 210 - (H_IS_MS_STEREO(hdr) ? 2 : 0);
    gain = ldexp_q2((float)(1 << (MAX_SCFI / 4)), MAX_SCFI - gain_exp);
    for (i = 0; i < (int)(gr->n_long_sfb + gr->n_short_sfb); i++) scf[i] = ldexp_q2(gain, iscf[i] << scf_shift);

    if (tap) {
        int n = gr->n_long_sfb + gr->n_short_sfb;
        memset(tap->iscf[ch], 0, 40);
        memcpy(tap->iscf[ch], iscf, n);
        memset(tap->scf[ch], 0, sizeof tap->scf[ch]);
        memcpy(tap->scf[ch], scf, n * sizeof(float));
        memcpy(tap->ist_pos[ch], ist_pos, 39);
        tap->ist_pos[ch][39] = 0;
        tap->gain_exp[ch] = gain_exp;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* x^(4/3) (minimp3.d:727-746)                                                                 */
static float pow_43(int x)
{
    float frac;
    int sign, mult = 256;
    if (x < 129) return L3_POW43[x];
    if (x < 1024) {
        mult = 16;
        x <<= 3;
    }
    sign = 2 * x & 64;
    frac = (float)((x & 63) - sign) / (float)((x & ~63) + sign);
    return L3_POW43[(x + sign) >> 6] * (1.0f + frac * ((4.0f / 3) + frac * (2.0f / 9))) * (float)mult;
}

/* ------------------------------------------------------------------------------------------ */
/* Huffman decode tables, built once from the canonical books.  Same scheme as the reference's
 * pre-built trees (5-bit first peek, then linked sub-tables; minimp3.d:795-803) but constructed
 * here; a complete prefix code decodes identically under any table layout.
 * entry >= 0: leaf   = len<<8 | v1<<4 | v0     entry < 0: link = -((offset<<3) | width)          */
static int16_t* g_pair_lut[L3_NBOOKS];
static int16_t g_zero_book[32];
static uint8_t g_c1_lut[2][64]; /* 6-bit peek -> flags<<4 | len */

static int book_find(int book, int len, uint32_t code)
{
    const uint8_t* hl = L3_HLEN + book * 256;
    const uint32_t* hc = L3_HCODE + book * 256;
    for (int s = 0; s < 256; s++)
        if (hl[s] == len && hc[s] == code) return s;
    return -1;
}

static int book_maxlen_under(int book, uint32_t prefix, int plen)
{
    const uint8_t* hl = L3_HLEN + book * 256;
    const uint32_t* hc = L3_HCODE + book * 256;
    int m = 0;
    for (int s = 0; s < 256; s++)
        if (hl[s] > plen && (hc[s] >> (hl[s] - plen)) == prefix && hl[s] > m) m = hl[s];
    return m;
}

static int build_node(int book, int16_t* lut, int* used, uint32_t prefix, int plen, int width)
{
    int base = *used;
    *used += 1 << width;
    for (int v = 0; v < (1 << width); v++) {
        uint32_t bits = (prefix << width) | (uint32_t)v;
        int found = 0;
        for (int l = plen + 1; l <= plen + width && !found; l++) {
            int s = book_find(book, l, bits >> (plen + width - l));
            if (s >= 0) {
                lut[base + v] = (int16_t)(((l - plen) << 8) | ((s & 15) << 4) | (s >> 4)); /* s = v0*16+v1 */
                found = 1;
            }
        }
        if (!found) {
            int rest = book_maxlen_under(book, bits, plen + width) - (plen + width);
            int w = rest > 6 ? 6 : rest;
            int child = build_node(book, lut, used, bits, plen + width, w);
            lut[base + v] = (int16_t)(-((child << 3) | w));
        }
    }
    return base;
}

__attribute__((constructor)) static void build_luts(void)
{
    for (int b = 0; b < L3_NBOOKS; b++) {
        int16_t* lut = (int16_t*)calloc(4096, sizeof(int16_t));
        int used = 0;
        /* books whose longest code is < 5 bits still use the 5-bit first peek */
        build_node(b, lut, &used, 0, 0, 5);
        if (used > 4096) abort();
        g_pair_lut[b] = lut;
    }
    memset(g_zero_book, 0, sizeof g_zero_book);
    for (int t = 0; t < 2; t++)
        for (int v = 0; v < 64; v++)
            for (int f = 0; f < 16; f++) {
                int ln = L3_C1LEN[t * 16 + f];
                if ((v >> (6 - ln)) == L3_C1CODE[t * 16 + f]) g_c1_lut[t][v] = (uint8_t)((f << 4) | ln);
            }
}

/* Huffman + requantisation of one granule-channel (minimp3.d:748-883).  `bs` is the shared
 * main-data reader; bytes past its limit are read like the reference does (callers pad). */
typedef struct {
    const uint8_t* next;
    uint32_t cache;
    int sh;
} hbits_t;

#define HB_FLUSH(hb, n) do { (hb).cache <<= (n); (hb).sh += (n); } while (0)
#define HB_REFILL(hb) while ((hb).sh >= 0) { (hb).cache |= (uint32_t)*(hb).next++ << (hb).sh; (hb).sh -= 8; }

static void huffman(float* dst, bitrd_t* bs, const gr_info_t* gr_info, const float* scf, int layer3gr_limit,
                    int16_t* tap_is)
{
    float one = 0.0f;
    int ireg = 0, big_val_cnt = gr_info->big_values;
    const uint8_t* sfb = gr_info->sfbtab;
    const uint8_t* p = bs->buf + bs->pos / 8;
    hbits_t hb;
    int pairs_to_decode, np;
    float* const dst0 = dst;
    hb.cache = ((((uint32_t)p[0] * 256u + p[1]) * 256u + p[2]) * 256u + p[3]) << (bs->pos & 7);
    hb.sh = (bs->pos & 7) - 8;
    hb.next = p + 4;

    while (big_val_cnt > 0) {
        int tab_num = gr_info->table_select[ireg];
        int sfb_cnt = gr_info->region_count[ireg++];
        int book = L3_SEL2BOOK[tab_num];
        const int16_t* codebook = book < 0 ? g_zero_book : g_pair_lut[book];
        int linbits = L3_LINBITS[tab_num];
        do {
            np = *sfb++ / 2;
            pairs_to_decode = imin(big_val_cnt, np);
            one = *scf++;
            do {
                int j, w = 5;
                int leaf = codebook[hb.cache >> (32 - w)];
                while (leaf < 0) {
                    HB_FLUSH(hb, w);
                    w = (-leaf) & 7;
                    leaf = codebook[((-leaf) >> 3) + (hb.cache >> (32 - w))];
                }
                HB_FLUSH(hb, leaf >> 8);

                for (j = 0; j < 2; j++, dst++, leaf >>= 4) {
                    int lsb = leaf & 0x0F;
                    if (linbits && lsb == 15) {
                        lsb += hb.cache >> (32 - linbits);
                        HB_FLUSH(hb, linbits);
                        HB_REFILL(hb);
                        *dst = one * pow_43(lsb) * ((int32_t)hb.cache < 0 ? -1 : 1);
                    } else {
                        /* reference (minimp3.d:816): g_pow43[16 + lsb - 16*sign]*one.  The table's lower half is the
                         * negation of the upper half EXCEPT entry 0, which is +0 in both (minimp3.d:722-724): a zero
                         * coefficient followed by a 1 bit stays +0.0, it does not become -0.0 */
                        float p43 = L3_POW43[lsb];
                        *dst = ((hb.cache >> 31) && lsb ? -p43 : p43) * one;
                    }
                    if (tap_is) tap_is[dst - dst0] = (int16_t)((lsb && (hb.cache >> 31)) ? -lsb : lsb);
                    HB_FLUSH(hb, lsb ? 1 : 0);
                }
                HB_REFILL(hb);
            } while (--pairs_to_decode);
        } while ((big_val_cnt -= np) > 0 && --sfb_cnt >= 0);
    }

    for (np = 1 - big_val_cnt;; dst += 4) {
        int leaf = g_c1_lut[gr_info->count1_table ? 1 : 0][hb.cache >> (32 - 6)];
        int flags = leaf >> 4;
        HB_FLUSH(hb, leaf & 15);
        if (((hb.next - bs->buf) * 8 - 24 + hb.sh) > layer3gr_limit) break;
#define RELOAD_SCALEFACTOR if (!--np) { np = *sfb++ / 2; if (!np) break; one = *scf++; }
#define DEQ_COUNT1(s)                                                       \
    if (flags & (8 >> (s))) {                                               \
        int neg_ = (int32_t)hb.cache < 0;                                   \
        dst[s] = neg_ ? -one : one;                                         \
        if (tap_is) tap_is[(dst - dst0) + (s)] = (int16_t)(neg_ ? -1 : 1);  \
        HB_FLUSH(hb, 1);                                                    \
    }
        RELOAD_SCALEFACTOR;
        DEQ_COUNT1(0);
        DEQ_COUNT1(1);
        RELOAD_SCALEFACTOR;
        DEQ_COUNT1(2);
        DEQ_COUNT1(3);
        HB_REFILL(hb);
    }
    bs->pos = layer3gr_limit;
}

/* ------------------------------------------------------------------------------------------ */
/* stereo (minimp3.d:885-982)                                                                  */
static void midside_stereo(float* left, int n)
{
    float* right = left + 576;
    for (int i = 0; i < n; i++) {
        float a = left[i];
        float b = right[i];
        left[i] = a + b;
        right[i] = a - b;
    }
}

static void intensity_stereo_band(float* left, int n, float kl, float kr)
{
    for (int i = 0; i < n; i++) {
        left[i + 576] = left[i] * kr;
        left[i] = left[i] * kl;
    }
}

static void stereo_top_band(const float* right, const uint8_t* sfb, int nbands, int max_band[3])
{
    max_band[0] = max_band[1] = max_band[2] = -1;
    for (int i = 0; i < nbands; i++) {
        for (int k = 0; k < sfb[i]; k += 2) {
            if (right[k] != 0 || right[k + 1] != 0) {
                max_band[i % 3] = i;
                break;
            }
        }
        right += sfb[i];
    }
}

static void stereo_process(float* left, const uint8_t* ist_pos, const uint8_t* sfb, const uint8_t* hdr,
                           int max_band[3], int mpeg2_sh)
{
    unsigned max_pos = H_TEST_MPEG1(hdr) ? 7 : 64;
    for (unsigned i = 0; sfb[i]; i++) {
        unsigned ipos = ist_pos[i];
        if ((int)i > max_band[i % 3] && ipos < max_pos) {
            float kl, kr, s = H_TEST_MS_STEREO(hdr) ? 1.41421356f : 1;
            if (H_TEST_MPEG1(hdr)) {
                kl = L3_PAN[2 * ipos];
                kr = L3_PAN[2 * ipos + 1];
            } else {
                kl = 1;
                kr = ldexp_q2(1, (ipos + 1) >> 1 << mpeg2_sh);
                if (ipos & 1) {
                    kl = kr;
                    kr = 1;
                }
            }
            intensity_stereo_band(left, sfb[i], kl * s, kr * s);
        } else if (H_TEST_MS_STEREO(hdr)) {
            midside_stereo(left, sfb[i]);
        }
        left += sfb[i];
    }
}

static void intensity_stereo(float* left, uint8_t* ist_pos, const gr_info_t* gr, const uint8_t* hdr)
{
    int max_band[3];
    int n_sfb = gr->n_long_sfb + gr->n_short_sfb;
    int i, max_blocks = gr->n_short_sfb ? 3 : 1;

    stereo_top_band(left + 576, gr->sfbtab, n_sfb, max_band);
    if (gr->n_long_sfb) max_band[0] = max_band[1] = max_band[2] = imax(imax(max_band[0], max_band[1]), max_band[2]);
    for (i = 0; i < max_blocks; i++) {
        int default_pos = H_TEST_MPEG1(hdr) ? 3 : 0;
        int itop = n_sfb - max_blocks + i;
        int prev = itop - max_blocks;
        ist_pos[itop] = (uint8_t)(max_band[i] >= prev ? default_pos : ist_pos[prev]);
    }
    stereo_process(left, ist_pos, gr->sfbtab, hdr, max_band, gr[1].scalefac_compress & 1);
}

/* ------------------------------------------------------------------------------------------ */
/* reorder + alias reduction (minimp3.d:984-1020)                                              */
static void reorder(float* grbuf, float* scratch, const uint8_t* sfb)
{
    int i, len;
    float *src = grbuf, *dst = scratch;
    for (; 0 != (len = *sfb); sfb += 3, src += 2 * len) {
        for (i = 0; i < len; i++, src++) {
            *dst++ = src[0 * len];
            *dst++ = src[1 * len];
            *dst++ = src[2 * len];
        }
    }
    memcpy(grbuf, scratch, (dst - scratch) * sizeof(float));
}

static void antialias(float* grbuf, int nbands)
{
    for (; nbands > 0; nbands--, grbuf += 18) {
        for (int i = 0; i < 8; i++) {
            float u = grbuf[18 + i];
            float d = grbuf[17 - i];
            grbuf[18 + i] = u * L3_AA[i] - d * L3_AA[8 + i];
            grbuf[17 - i] = u * L3_AA[8 + i] + d * L3_AA[i];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* IMDCT (minimp3.d:1022-1168)                                                                 */
static void dct3_9(float* y)
{
    float s0, s1, s2, s3, s4, s5, s6, s7, s8, t0, t2, t4;

    s0 = y[0]; s2 = y[2]; s4 = y[4]; s6 = y[6]; s8 = y[8];
    t0 = s0 + s6 * 0.5f;
    s0 -= s6;
    t4 = (s4 + s2) * 0.93969262f;
    t2 = (s8 + s2) * 0.76604444f;
    s6 = (s4 - s8) * 0.17364818f;
    s4 += s8 - s2;

    s2 = s0 - s4 * 0.5f;
    y[4] = s4 + s0;
    s8 = t0 - t2 + s6;
    s0 = t0 - t4 + t2;
    s4 = t0 + t4 - s6;

    s1 = y[1]; s3 = y[3]; s5 = y[5]; s7 = y[7];

    s3 *= 0.86602540f;
    t0 = (s5 + s1) * 0.98480775f;
    t4 = (s5 - s7) * 0.34202014f;
    t2 = (s1 + s7) * 0.64278761f;
    s1 = (s1 - s5 - s7) * 0.86602540f;

    s5 = t0 - s3 - t2;
    s7 = t4 - s3 - t0;
    s3 = t4 + s3 - t2;

    y[0] = s4 - s7;
    y[1] = s2 + s1;
    y[2] = s0 - s3;
    y[3] = s8 + s5;
    y[5] = s8 - s5;
    y[6] = s0 + s3;
    y[7] = s2 - s1;
    y[8] = s4 + s7;
}

static void imdct36(float* grbuf, float* overlap, const float* window, int nbands)
{
    for (int j = 0; j < nbands; j++, grbuf += 18, overlap += 9) {
        float co[9], si[9];
        int i;
        co[0] = -grbuf[0];
        si[0] = grbuf[17];
        for (i = 0; i < 4; i++) {
            si[8 - 2 * i] = grbuf[4 * i + 1] - grbuf[4 * i + 2];
            co[1 + 2 * i] = grbuf[4 * i + 1] + grbuf[4 * i + 2];
            si[7 - 2 * i] = grbuf[4 * i + 4] - grbuf[4 * i + 3];
            co[2 + 2 * i] = -(grbuf[4 * i + 3] + grbuf[4 * i + 4]);
        }
        dct3_9(co);
        dct3_9(si);

        si[1] = -si[1];
        si[3] = -si[3];
        si[5] = -si[5];
        si[7] = -si[7];

        for (i = 0; i < 9; i++) {
            float ovl = overlap[i];
            float sum = co[i] * L3_TWID9[9 + i] + si[i] * L3_TWID9[0 + i];
            overlap[i] = co[i] * L3_TWID9[0 + i] - si[i] * L3_TWID9[9 + i];
            grbuf[i] = ovl * window[0 + i] - sum * window[9 + i];
            grbuf[17 - i] = ovl * window[9 + i] + sum * window[0 + i];
        }
    }
}

static void idct3(float x0, float x1, float x2, float* dst)
{
    float m1 = x1 * 0.86602540f;
    float a1 = x0 - x2 * 0.5f;
    dst[1] = x0 + x2;
    dst[0] = a1 + m1;
    dst[2] = a1 - m1;
}

static void imdct12(float* x, float* dst, float* overlap)
{
    float co[3], si[3];
    idct3(-x[0], x[6] + x[3], x[12] + x[9], co);
    idct3(x[15], x[12] - x[9], x[6] - x[3], si);
    si[1] = -si[1];
    for (int i = 0; i < 3; i++) {
        float ovl = overlap[i];
        float sum = co[i] * L3_TWID3[3 + i] + si[i] * L3_TWID3[0 + i];
        overlap[i] = co[i] * L3_TWID3[0 + i] - si[i] * L3_TWID3[3 + i];
        dst[i] = ovl * L3_TWID3[2 - i] - sum * L3_TWID3[5 - i];
        dst[5 - i] = ovl * L3_TWID3[5 - i] + sum * L3_TWID3[2 - i];
    }
}

static void imdct_short(float* grbuf, float* overlap, int nbands)
{
    for (; nbands > 0; nbands--, overlap += 9, grbuf += 18) {
        float tmp[18];
        memcpy(tmp, grbuf, sizeof tmp);
        memcpy(grbuf, overlap, 6 * sizeof(float));
        imdct12(tmp, grbuf + 6, overlap + 6);
        imdct12(tmp + 1, grbuf + 12, overlap + 6);
        imdct12(tmp + 2, overlap, overlap + 6);
    }
}

static void change_sign(float* grbuf)
{
    int b, i;
    for (b = 0, grbuf += 18; b < 32; b += 2, grbuf += 36)
        for (i = 1; i < 18; i += 2) grbuf[i] = -grbuf[i];
}

static void imdct_gr(float* grbuf, float* overlap, unsigned block_type, unsigned n_long_bands)
{
    if (n_long_bands) {
        imdct36(grbuf, overlap, L3_MDCT_WINDOW, n_long_bands);
        grbuf += 18 * n_long_bands;
        overlap += 9 * n_long_bands;
    }
    if (block_type == SHORT_BLOCK_TYPE)
        imdct_short(grbuf, overlap, 32 - n_long_bands);
    else
        imdct36(grbuf, overlap, L3_MDCT_WINDOW + 18 * (block_type == STOP_BLOCK_TYPE), 32 - n_long_bands);
}

/* ------------------------------------------------------------------------------------------ */
/* per-frame scratch (minimp3.d:198-207) and reservoir (minimp3.d:1170-1194)                    */
typedef struct {
    bitrd_t bs;
    uint8_t maindata[MAX_BITRESERVOIR_BYTES + MAX_L3_FRAME_PAYLOAD_BYTES + 16];
    gr_info_t gr_info[4];
    float grbuf[2][576];
    float scf[40];
    float syn[18 + 15][2 * 32];
    uint8_t ist_pos[2][39];
} scratch_t;

static void save_reservoir(l3o_dec_t* h, scratch_t* s)
{
    int pos = (s->bs.pos + 7) / 8u;
    int remains = s->bs.limit / 8u - pos;
    if (remains > MAX_BITRESERVOIR_BYTES) {
        pos += remains - MAX_BITRESERVOIR_BYTES;
        remains = MAX_BITRESERVOIR_BYTES;
    }
    if (remains > 0) memmove(h->reserv_buf, s->maindata + pos, remains);
    h->reserv = remains;
}

static int restore_reservoir(l3o_dec_t* h, bitrd_t* bs, scratch_t* s, int main_data_begin)
{
    int frame_bytes = (bs->limit - bs->pos) / 8;
    int bytes_have = imin(h->reserv, main_data_begin);
    memcpy(s->maindata, h->reserv_buf + imax(0, h->reserv - main_data_begin), imin(h->reserv, main_data_begin));
    memcpy(s->maindata + bytes_have, bs->buf + bs->pos / 8, frame_bytes);
    /* the reference leaves the bytes past the main data uninitialised (the Huffman reader prefetches
     * up to 4+ bytes there, minimp3.d:775-778); zero them so the oracle is deterministic */
    memset(s->maindata + bytes_have + frame_bytes, 0, 16);
    rd_init(&s->bs, s->maindata, bytes_have + frame_bytes);
    return h->reserv >= main_data_begin;
}

/* one granule, all channels (minimp3.d:1196-1230) */
static void decode_granule(l3o_dec_t* h, scratch_t* s, gr_info_t* gr_info, int nch, l3o_granule_tap_t* tap)
{
    int ch;
    double t0 = g_timers_on ? now_s() : 0;
    for (ch = 0; ch < nch; ch++) {
        int layer3gr_limit = s->bs.pos + gr_info[ch].part_23_length;
        decode_scalefactors(h->header, s->ist_pos[ch], &s->bs, gr_info + ch, s->scf, ch, tap);
        if (tap) memset(tap->is[ch], 0, sizeof tap->is[ch]);
        huffman(s->grbuf[ch], &s->bs, gr_info + ch, s->scf, layer3gr_limit, tap ? tap->is[ch] : NULL);
    }
    if (g_timers_on) { double t1 = now_s(); t_timer[1] += t1 - t0; t0 = t1; }
    if (tap) memcpy(tap->xr, s->grbuf, sizeof tap->xr);

    if (H_TEST_I_STEREO(h->header)) {
        intensity_stereo(s->grbuf[0], s->ist_pos[1], gr_info, h->header);
    } else if (H_IS_MS_STEREO(h->header)) {
        midside_stereo(s->grbuf[0], 576);
    }
    if (g_timers_on) { double t1 = now_s(); t_timer[2] += t1 - t0; t0 = t1; }
    if (tap) memcpy(tap->st, s->grbuf, sizeof tap->st);

    for (ch = 0; ch < nch; ch++, gr_info++) {
        int aa_bands = 31;
        int n_long_bands = (gr_info->mixed_block_flag ? 2 : 0) << (int)(H_GET_MY_SAMPLE_RATE(h->header) == 2);

        if (gr_info->n_short_sfb) {
            aa_bands = n_long_bands - 1;
            reorder(s->grbuf[ch] + n_long_bands * 18, s->syn[0], gr_info->sfbtab + gr_info->n_long_sfb);
        }
        antialias(s->grbuf[ch], aa_bands);
        if (g_timers_on) { double t1 = now_s(); t_timer[3] += t1 - t0; t0 = t1; }
        imdct_gr(s->grbuf[ch], h->mdct_overlap[ch], gr_info->block_type, n_long_bands);
        change_sign(s->grbuf[ch]);
        if (g_timers_on) { double t1 = now_s(); t_timer[4] += t1 - t0; t0 = t1; }
    }
    if (tap) memcpy(tap->im, s->grbuf, sizeof tap->im);
}

/* ------------------------------------------------------------------------------------------ */
/* polyphase synthesis (minimp3.d:1232-1434)                                                   */
static void dct_ii(float* grbuf, int n)
{
    for (int k = 0; k < n; k++) {
        float t[4][8], *x, *y = grbuf + k;
        int i;

        for (x = t[0], i = 0; i < 8; i++, x++) {
            float x0 = y[i * 18];
            float x1 = y[(15 - i) * 18];
            float x2 = y[(16 + i) * 18];
            float x3 = y[(31 - i) * 18];
            float t0 = x0 + x3;
            float t1 = x1 + x2;
            float t2 = (x1 - x2) * L3_SEC[3 * i + 0];
            float t3 = (x0 - x3) * L3_SEC[3 * i + 1];
            x[0] = t0 + t1;
            x[8] = (t0 - t1) * L3_SEC[3 * i + 2];
            x[16] = t3 + t2;
            x[24] = (t3 - t2) * L3_SEC[3 * i + 2];
        }
        for (x = t[0], i = 0; i < 4; i++, x += 8) {
            float x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = x[4], x5 = x[5], x6 = x[6], x7 = x[7], xt;
            xt = x0 - x7; x0 += x7;
            x7 = x1 - x6; x1 += x6;
            x6 = x2 - x5; x2 += x5;
            x5 = x3 - x4; x3 += x4;
            x4 = x0 - x3; x0 += x3;
            x3 = x1 - x2; x1 += x2;
            x[0] = x0 + x1;
            x[4] = (x0 - x1) * 0.70710677f;
            x5 = x5 + x6;
            x6 = (x6 + x7) * 0.70710677f;
            x7 = x7 + xt;
            x3 = (x3 + x4) * 0.70710677f;
            x5 -= x7 * 0.198912367f; /* rotate by PI/8 */
            x7 += x5 * 0.382683432f;
            x5 -= x7 * 0.198912367f;
            x0 = xt - x6; xt += x6;
            x[1] = (xt + x7) * 0.50979561f;
            x[2] = (x4 + x3) * 0.54119611f;
            x[3] = (x0 - x5) * 0.60134488f;
            x[5] = (x0 + x5) * 0.89997619f;
            x[6] = (x4 - x3) * 1.30656302f;
            x[7] = (xt - x7) * 2.56291556f;
        }
        for (i = 0; i < 7; i++, y += 4 * 18) {
            y[0 * 18] = t[0][i];
            y[1 * 18] = t[2][i] + t[3][i] + t[3][i + 1];
            y[2 * 18] = t[1][i] + t[1][i + 1];
            y[3 * 18] = t[2][i + 1] + t[3][i] + t[3][i + 1];
        }
        y[0 * 18] = t[0][7];
        y[1 * 18] = t[2][7] + t[3][7];
        y[2 * 18] = t[1][7];
        y[3 * 18] = t[3][7];
    }
}

static float scale_pcm(float sample) { return sample * (1.0f / 32768.0f); }

static void synth_pair(float* pcm, int nch, const float* z)
{
    float a;
    a = (z[14 * 64] - z[0]) * 29;
    a += (z[1 * 64] + z[13 * 64]) * 213;
    a += (z[12 * 64] - z[2 * 64]) * 459;
    a += (z[3 * 64] + z[11 * 64]) * 2037;
    a += (z[10 * 64] - z[4 * 64]) * 5153;
    a += (z[5 * 64] + z[9 * 64]) * 6574;
    a += (z[8 * 64] - z[6 * 64]) * 37489;
    a += z[7 * 64] * 75038;
    pcm[0] = scale_pcm(a);

    z += 2;
    a = z[14 * 64] * 104;
    a += z[12 * 64] * 1567;
    a += z[10 * 64] * 9727;
    a += z[8 * 64] * 64019;
    a += z[6 * 64] * -9975;
    a += z[4 * 64] * -45;
    a += z[2 * 64] * 146;
    a += z[0 * 64] * -5;
    pcm[16 * nch] = scale_pcm(a);
}

/* The reference's window array is i-major from i=14 down; ours is L3_WIN[(k*2+c)*15+i]. */
#define WIN(i, k, c) L3_WIN[((k) * 2 + (c)) * 15 + (i)]

static void synth(float* xl, float* dstl, int nch, float* lins)
{
    int i;
    float* xr = xl + 576 * (nch - 1);
    float* dstr = dstl + (nch - 1);
    float* zlin = lins + 15 * 64;

    zlin[4 * 15] = xl[18 * 16];
    zlin[4 * 15 + 1] = xr[18 * 16];
    zlin[4 * 15 + 2] = xl[0];
    zlin[4 * 15 + 3] = xr[0];

    zlin[4 * 31] = xl[1 + 18 * 16];
    zlin[4 * 31 + 1] = xr[1 + 18 * 16];
    zlin[4 * 31 + 2] = xl[1];
    zlin[4 * 31 + 3] = xr[1];

    synth_pair(dstr, nch, lins + 4 * 15 + 1);
    synth_pair(dstr + 32 * nch, nch, lins + 4 * 15 + 64 + 1);
    synth_pair(dstl, nch, lins + 4 * 15);
    synth_pair(dstl + 32 * nch, nch, lins + 4 * 15 + 64);

    for (i = 14; i >= 0; i--) {
        float a[4], b[4];
        int j, k;

        zlin[4 * i] = xl[18 * (31 - i)];
        zlin[4 * i + 1] = xr[18 * (31 - i)];
        zlin[4 * i + 2] = xl[1 + 18 * (31 - i)];
        zlin[4 * i + 3] = xr[1 + 18 * (31 - i)];
        zlin[4 * (i + 16)] = xl[1 + 18 * (1 + i)];
        zlin[4 * (i + 16) + 1] = xr[1 + 18 * (1 + i)];
        zlin[4 * (i - 16) + 2] = xl[18 * (1 + i)];
        zlin[4 * (i - 16) + 3] = xr[18 * (1 + i)];

        /* S0(0) S2(1) S1(2) S2(3) S1(4) S2(5) S1(6) S2(7)  (minimp3.d:1373-1395) */
        for (k = 0; k < 8; k++) {
            float w0 = WIN(i, k, 0), w1 = WIN(i, k, 1);
            float* vz = &zlin[4 * i - k * 64];
            float* vy = &zlin[4 * i - (15 - k) * 64];
            if (k == 0) {
                for (j = 0; j < 4; j++) b[j] = vz[j] * w1 + vy[j] * w0, a[j] = vz[j] * w0 - vy[j] * w1;
            } else if (k & 1) {
                for (j = 0; j < 4; j++) b[j] += vz[j] * w1 + vy[j] * w0, a[j] += vy[j] * w1 - vz[j] * w0;
            } else {
                for (j = 0; j < 4; j++) b[j] += vz[j] * w1 + vy[j] * w0, a[j] += vz[j] * w0 - vy[j] * w1;
            }
        }

        dstr[(15 - i) * nch] = scale_pcm(a[1]);
        dstr[(17 + i) * nch] = scale_pcm(b[1]);
        dstl[(15 - i) * nch] = scale_pcm(a[0]);
        dstl[(17 + i) * nch] = scale_pcm(b[0]);
        dstr[(47 - i) * nch] = scale_pcm(a[3]);
        dstr[(49 + i) * nch] = scale_pcm(b[3]);
        dstl[(47 - i) * nch] = scale_pcm(a[2]);
        dstl[(49 + i) * nch] = scale_pcm(b[2]);
    }
}

static void synth_granule(float* qmf_state, float* grbuf, int nbands, int nch, float* pcm, float* lins,
                          l3o_granule_tap_t* tap)
{
    int i;
    double t0 = g_timers_on ? now_s() : 0;
    for (i = 0; i < nch; i++) dct_ii(grbuf + 576 * i, nbands);
    if (g_timers_on) { double t1 = now_s(); t_timer[5] += t1 - t0; t0 = t1; }
    if (tap) memcpy(tap->dct, grbuf, sizeof tap->dct);

    memcpy(lins, qmf_state, sizeof(float) * 15 * 64);
    for (i = 0; i < nbands; i += 2) synth(grbuf + i, pcm + 32 * nch * i, nch, lins + i * 64);

    if (nch == 1) {
        for (i = 0; i < 15 * 64; i += 2) qmf_state[i] = lins[nbands * 64 + i];
    } else {
        memcpy(qmf_state, lins + nbands * 64, sizeof(float) * 15 * 64);
    }
    if (g_timers_on) t_timer[6] += now_s() - t0;
}

/* ------------------------------------------------------------------------------------------ */
/* frame sync (minimp3.d:1436-1485)                                                            */
static int match_frame(const uint8_t* hdr, int mp3_bytes, int frame_bytes)
{
    int i, nmatch;
    for (i = 0, nmatch = 0; nmatch < MAX_FRAME_SYNC_MATCHES; nmatch++) {
        i += l3o_hdr_frame_bytes(hdr + i, frame_bytes) + l3o_hdr_padding(hdr + i);
        if (i + HDR_SIZE > mp3_bytes) return nmatch > 0;
        if (!hdr_compare(hdr, hdr + i)) return 0;
    }
    return 1;
}

static int find_frame(const uint8_t* mp3, int mp3_bytes, int* free_format_bytes, int* ptr_frame_bytes)
{
    int i, k;
    for (i = 0; i < mp3_bytes - HDR_SIZE; i++, mp3++) {
        if (l3o_hdr_valid(mp3)) {
            int frame_bytes = l3o_hdr_frame_bytes(mp3, *free_format_bytes);
            int frame_and_padding = frame_bytes + l3o_hdr_padding(mp3);

            for (k = HDR_SIZE; !frame_bytes && k < MAX_FREE_FORMAT_FRAME_SIZE && i + 2 * k < mp3_bytes - HDR_SIZE; k++) {
                if (hdr_compare(mp3, mp3 + k)) {
                    int fb = k - l3o_hdr_padding(mp3);
                    int nextfb = fb + l3o_hdr_padding(mp3 + k);
                    if (i + k + nextfb + HDR_SIZE > mp3_bytes || !hdr_compare(mp3, mp3 + k + nextfb)) continue;
                    frame_and_padding = k;
                    frame_bytes = fb;
                    *free_format_bytes = fb;
                }
            }
            if ((frame_bytes && i + frame_and_padding <= mp3_bytes && match_frame(mp3, mp3_bytes - i, frame_bytes)) ||
                (!i && frame_and_padding == mp3_bytes)) {
                *ptr_frame_bytes = frame_and_padding;
                return i;
            }
            *free_format_bytes = 0;
        }
    }
    *ptr_frame_bytes = 0;
    return mp3_bytes;
}

/* ------------------------------------------------------------------------------------------ */
/* Layer I / II (minimp3.d:286-484)                                                             */
#define H_GET_STEREO_MODE(h) (((h)[3] >> 6) & 3)
#define H_GET_STEREO_MODE_EXT(h) (((h)[3] >> 4) & 3)
#define MODE_MONO 3
#define MODE_JOINT_STEREO 1

typedef struct {                 /* L12_scale_info, minimp3.d:177-183 */
    float scf[3 * 64];
    uint8_t total_bands, stereo_bands, bitalloc[64], scfcod[64];
} l12_scale_info_t;

typedef struct { uint8_t tab_offset, code_tab_width, band_count; } l12_subband_alloc_t;   /* minimp3.d:172-175 */

/* minimp3.d:284-350 */
static const l12_subband_alloc_t* l12_subband_alloc_table(const uint8_t* hdr, l12_scale_info_t* sci)
{
    static const l12_subband_alloc_t g_alloc_L1[] = { { 76, 4, 32 } };
    static const l12_subband_alloc_t g_alloc_L2M2[] = { { 60, 4, 4 }, { 44, 3, 7 }, { 44, 2, 19 } };
    static const l12_subband_alloc_t g_alloc_L2M1[] = { { 0, 4, 3 }, { 16, 4, 8 }, { 32, 3, 12 }, { 40, 2, 7 } };
    static const l12_subband_alloc_t g_alloc_L2M1_lowrate[] = { { 44, 4, 2 }, { 44, 3, 10 } };
    const l12_subband_alloc_t* alloc;
    int mode = H_GET_STEREO_MODE(hdr);
    int nbands, stereo_bands = (mode == MODE_MONO) ? 0 : (mode == MODE_JOINT_STEREO) ? (H_GET_STEREO_MODE_EXT(hdr) << 2) + 4 : 32;

    if (H_IS_LAYER_1(hdr)) {
        alloc = g_alloc_L1;
        nbands = 32;
    } else if (!H_TEST_MPEG1(hdr)) {
        alloc = g_alloc_L2M2;
        nbands = 30;
    } else {
        int sample_rate_idx = H_GET_SAMPLE_RATE(hdr);
        unsigned kbps = l3o_hdr_bitrate_kbps(hdr) >> (int)(mode != MODE_MONO);
        if (!kbps) kbps = 192; /* free-format */
        alloc = g_alloc_L2M1;
        nbands = 27;
        if (kbps < 56) {
            alloc = g_alloc_L2M1_lowrate;
            nbands = sample_rate_idx == 2 ? 12 : 8;
        } else if (kbps >= 96 && sample_rate_idx != 1) {
            nbands = 30;
        }
    }
    sci->total_bands = (uint8_t)nbands;
    sci->stereo_bands = (uint8_t)imin(stereo_bands, nbands);
    return alloc;
}

/* minimp3.d:352-385.  The D literals are doubles converted to float by the static initialiser; so are these. */
static void l12_read_scalefactors(bitrd_t* bs, uint8_t* pba, uint8_t* scfcod, int bands, float* scf)
{
    static const float g_deq_L12[18 * 3] = {
        3.17891e-07, 2.52311e-07, 2.00259e-07, 1.36239e-07, 1.08133e-07, 8.58253e-08,
        6.35783e-08, 5.04621e-08, 4.00518e-08, 3.07637e-08, 2.44172e-08, 1.93799e-08,
        1.51377e-08, 1.20148e-08, 9.53615e-09, 7.50925e-09, 5.96009e-09, 4.73053e-09,
        3.7399e-09, 2.96836e-09, 2.35599e-09, 1.86629e-09, 1.48128e-09, 1.17569e-09,
        9.32233e-10, 7.39914e-10, 5.8727e-10, 4.65889e-10, 3.69776e-10, 2.93492e-10,
        2.32888e-10, 1.84843e-10, 1.4671e-10, 1.1643e-10, 9.24102e-11, 7.3346e-11,
        5.82112e-11, 4.62023e-11, 3.66708e-11, 2.91047e-11, 2.31004e-11, 1.83348e-11,
        1.45521e-11, 1.155e-11, 9.16727e-12, 3.17891e-07, 2.52311e-07, 2.00259e-07,
        1.90735e-07, 1.51386e-07, 1.20155e-07, 1.05964e-07, 8.41035e-08, 6.6753e-08
    };
    int i, m;
    for (i = 0; i < bands; i++) {
        float s = 0;
        int ba = *pba++;
        int mask = ba ? 4 + ((19 >> scfcod[i]) & 3) : 0;
        for (m = 4; m; m >>= 1) {
            if (mask & m) {
                int b = (int)rd_bits(bs, 6);
                s = g_deq_L12[ba * 3 - 6 + b % 3] * (float)(1 << 21 >> b / 3);
            }
            *scf++ = s;
        }
    }
}

/* minimp3.d:387-435 */
static void l12_read_scale_info(const uint8_t* hdr, bitrd_t* bs, l12_scale_info_t* sci)
{
    static const uint8_t g_bitalloc_code_tab[] = {
        0, 17, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,
        0, 17, 18, 3, 19, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16,
        0, 17, 18, 3, 19, 4, 5, 16,
        0, 17, 18, 16,
        0, 17, 18, 19, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
        0, 17, 18, 3, 19, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14,
        0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16
    };
    const l12_subband_alloc_t* subband_alloc = l12_subband_alloc_table(hdr, sci);
    int i, k = 0, ba_bits = 0;
    const uint8_t* ba_code_tab = g_bitalloc_code_tab;

    for (i = 0; i < sci->total_bands; i++) {
        uint8_t ba;
        if (i == k) {
            k += subband_alloc->band_count;
            ba_bits = subband_alloc->code_tab_width;
            ba_code_tab = g_bitalloc_code_tab + subband_alloc->tab_offset;
            subband_alloc++;
        }
        ba = ba_code_tab[rd_bits(bs, ba_bits)];
        sci->bitalloc[2 * i] = ba;
        if (i < sci->stereo_bands) ba = ba_code_tab[rd_bits(bs, ba_bits)];
        sci->bitalloc[2 * i + 1] = sci->stereo_bands ? ba : 0;
    }
    for (i = 0; i < 2 * sci->total_bands; i++) {
        uint8_t temp = H_IS_LAYER_1(hdr) ? 2 : (uint8_t)rd_bits(bs, 2);
        sci->scfcod[i] = sci->bitalloc[i] ? temp : 6;
    }
    l12_read_scalefactors(bs, sci->bitalloc, sci->scfcod, sci->total_bands * 2, sci->scf);
    for (i = sci->stereo_bands; i < sci->total_bands; i++) sci->bitalloc[2 * i + 1] = 0;
}

/* minimp3.d:437-470 */
static int l12_dequantize_granule(float* grbuf, bitrd_t* bs, l12_scale_info_t* sci, int group_size)
{
    int i, j, k, choff = 576;
    for (j = 0; j < 4; j++) {
        float* dst = grbuf + group_size * j;
        for (i = 0; i < 2 * sci->total_bands; i++) {
            int ba = sci->bitalloc[i];
            if (ba != 0) {
                if (ba < 17) {
                    int half = (1 << (ba - 1)) - 1;
                    for (k = 0; k < group_size; k++) dst[k] = (float)((int)rd_bits(bs, ba) - half);
                } else {
                    unsigned mod = (2 << (ba - 17)) + 1;                       /* 3, 5, 9 */
                    unsigned code = rd_bits(bs, mod + 2 - (mod >> 3));         /* 5, 7, 10 */
                    for (k = 0; k < group_size; k++, code /= mod) dst[k] = (float)((int)(code % mod - mod / 2));
                }
            }
            dst += choff;
            choff = 18 - choff;
        }
    }
    return group_size * 4;
}

/* minimp3.d:472-484 */
static void l12_apply_scf_384(l12_scale_info_t* sci, const float* scf, float* dst)
{
    int i, k;
    memcpy(dst + 576 + sci->stereo_bands * 18, dst + sci->stereo_bands * 18, (sci->total_bands - sci->stereo_bands) * 18 * sizeof(float));
    for (i = 0; i < sci->total_bands; i++, dst += 18, scf += 6) {
        for (k = 0; k < 12; k++) {
            dst[k + 0] *= scf[0];
            dst[k + 576] *= scf[3];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* frame entry (minimp3.d:1487-1581), Layer III branch only                                    */
void l3o_init(l3o_dec_t* dec) { dec->header[0] = 0; }

int l3o_decode_frame(l3o_dec_t* dec, const uint8_t* mp3, int mp3_bytes, float* pcm, l3o_frame_info_t* info)
{
    int i = 0, igr, frame_size = 0, success = 1;
    const uint8_t* hdr;
    bitrd_t bs_frame[1];
    scratch_t scratch;
    memset(&scratch, 0, sizeof scratch); /* D default-initialises locals (minimp3.d:1497): bytes and ints start at 0 */

    if (mp3_bytes > 4 && dec->header[0] == 0xff && hdr_compare(dec->header, mp3)) {
        frame_size = l3o_hdr_frame_bytes(mp3, dec->free_format_bytes) + l3o_hdr_padding(mp3);
        if (frame_size != mp3_bytes && (frame_size + HDR_SIZE > mp3_bytes || !hdr_compare(mp3, mp3 + frame_size)))
            frame_size = 0;
    }
    if (!frame_size) {
        memset(dec, 0, sizeof(l3o_dec_t));
        i = find_frame(mp3, mp3_bytes, &dec->free_format_bytes, &frame_size);
        if (!frame_size || i + frame_size > mp3_bytes) {
            info->frame_bytes = i;
            return 0;
        }
    }

    hdr = mp3 + i;
    memcpy(dec->header, hdr, HDR_SIZE);
    info->frame_bytes = i + frame_size;
    info->frame_offset = i;
    info->channels = H_IS_MONO(hdr) ? 1 : 2;
    info->hz = l3o_hdr_sample_rate_hz(hdr);
    info->layer = 4 - H_GET_LAYER(hdr);
    info->bitrate_kbps = l3o_hdr_bitrate_kbps(hdr);

    if (!pcm) return l3o_hdr_frame_samples(hdr);

    rd_init(bs_frame, hdr + HDR_SIZE, frame_size - HDR_SIZE);
    if (H_IS_CRC(hdr)) rd_bits(bs_frame, 16);

    if (info->layer == 3) {
        double t0 = g_timers_on ? now_s() : 0;
        int main_data_begin = read_side_info(bs_frame, scratch.gr_info, hdr);
        if (main_data_begin < 0 || bs_frame->pos > bs_frame->limit) {
            l3o_init(dec);
            return 0;
        }
        /* The reference's scratch is an uninitialised stack object; ist_pos entries a granule does not
         * transmit are read by L3_intensity_stereo when channel 0 has more bands than channel 1 (UB upstream).
         * The oracle zeroes them per frame so that it is deterministic. */
        memset(scratch.ist_pos, 0, sizeof scratch.ist_pos);
        success = restore_reservoir(dec, bs_frame, &scratch, main_data_begin);
        if (g_timers_on) t_timer[0] += now_s() - t0;
        if (success) {
            for (igr = 0; igr < (H_TEST_MPEG1(hdr) ? 2 : 1); igr++, pcm += 576 * info->channels) {
                l3o_granule_tap_t* tap = NULL;
                if (t_tap) {
                    if (t_tap->count < t_tap->capacity) {
                        tap = &t_tap->rec[t_tap->count];
                        memset(tap, 0, sizeof *tap);
                    }
                    t_tap->count++;
                }
                memset(scratch.grbuf[0], 0, 576 * 2 * sizeof(float));
                decode_granule(dec, &scratch, scratch.gr_info + igr * info->channels, info->channels, tap);
                synth_granule(dec->qmf_state, scratch.grbuf[0], 18, info->channels, pcm, scratch.syn[0], tap);
            }
        }
        if (g_timers_on) t0 = now_s();
        save_reservoir(dec, &scratch);
        if (g_timers_on) t_timer[0] += now_s() - t0;
    } else {
        /* Layer I / II (minimp3.d:1557-1579) */
        l12_scale_info_t sci[1];
        l12_read_scale_info(hdr, bs_frame, sci);
        memset(scratch.grbuf[0], 0, 576 * 2 * sizeof(float));
        for (i = 0, igr = 0; igr < 3; igr++) {
            if (12 == (i += l12_dequantize_granule(scratch.grbuf[0] + i, bs_frame, sci, info->layer | 1))) {
                i = 0;
                l12_apply_scf_384(sci, sci->scf + igr, scratch.grbuf[0]);
                synth_granule(dec->qmf_state, scratch.grbuf[0], 12, info->channels, pcm, scratch.syn[0], NULL);
                memset(scratch.grbuf[0], 0, 576 * 2 * sizeof(float));
                pcm += 384 * info->channels;
            }
            if (bs_frame->pos > bs_frame->limit) {
                l3o_init(dec);
                return 0;
            }
        }
    }
    return success * l3o_hdr_frame_samples(dec->header);
}

/* ------------------------------------------------------------------------------------------ */
/* side-info parse helper for the stream layer (uses the types above)                          */
int l3o__side_info_main_bytes(const uint8_t* hdr, int frame_size, int* main_data_begin_out)
{
    /* returns the frame's own main-data byte count, or -1 when the side info is rejected */
    bitrd_t bs[1];
    gr_info_t gr[4];
    rd_init(bs, hdr + HDR_SIZE, frame_size - HDR_SIZE);
    if (H_IS_CRC(hdr)) rd_bits(bs, 16);
    int mdb = read_side_info(bs, gr, hdr);
    if (mdb < 0) return -1;
    if (main_data_begin_out) *main_data_begin_out = mdb;
    return (bs->limit - bs->pos) / 8;
}

int l3o__side_info_end_byte(const uint8_t* hdr, int frame_size)
{
    /* byte offset (from hdr) just past the side info; -1 when rejected (for the VBR-tag probe) */
    bitrd_t bs[1];
    gr_info_t gr[4];
    rd_init(bs, hdr + HDR_SIZE, frame_size - HDR_SIZE);
    if (H_IS_CRC(hdr)) rd_bits(bs, 16);
    if (read_side_info(bs, gr, hdr) < 0) return -1;
    return HDR_SIZE + bs->pos / 8;
}

int l3o__find_frame(const uint8_t* mp3, int mp3_bytes, int* free_format_bytes, int* ptr_frame_bytes)
{
    return find_frame(mp3, mp3_bytes, free_format_bytes, ptr_frame_bytes);
}
