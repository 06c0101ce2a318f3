// l3_kernels.cu -- see l3_kernels.cuh.  Compile with -fmad=false (bit-exactness contract).
#include "l3_kernels.cuh"

#include "l3_tables_gen.h"

namespace l3b {

// ---- constant-bank tables (only ever indexed uniformly across a warp, or tiny) ---------------------
__constant__ float c_expfrac[4];
__constant__ float c_aa[16];
__constant__ float c_twid9[18];
__constant__ float c_twid3[6];
__constant__ float c_mdctw[36];
__constant__ float c_sec[24];
__constant__ float c_pan[14];
__constant__ uint8_t c_partitions[84];
__constant__ uint8_t c_scfc_decode[16];
__constant__ uint8_t c_lsf_mod[24];
__constant__ uint8_t c_preamp[10];
__constant__ uint8_t c_linbits[32];
__constant__ int8_t c_sel2book[32];

void upload_constants() {
    cudaMemcpyToSymbol(c_expfrac, L3_EXPFRAC, sizeof c_expfrac);
    cudaMemcpyToSymbol(c_aa, L3_AA, sizeof c_aa);
    cudaMemcpyToSymbol(c_twid9, L3_TWID9, sizeof c_twid9);
    cudaMemcpyToSymbol(c_twid3, L3_TWID3, sizeof c_twid3);
    cudaMemcpyToSymbol(c_mdctw, L3_MDCT_WINDOW, sizeof c_mdctw);
    cudaMemcpyToSymbol(c_sec, L3_SEC, sizeof c_sec);
    cudaMemcpyToSymbol(c_pan, L3_PAN, sizeof c_pan);
    cudaMemcpyToSymbol(c_partitions, L3_SCF_PARTITIONS, sizeof c_partitions);
    cudaMemcpyToSymbol(c_scfc_decode, L3_SCFC_DECODE, sizeof c_scfc_decode);
    cudaMemcpyToSymbol(c_lsf_mod, L3_LSF_MOD, sizeof c_lsf_mod);
    cudaMemcpyToSymbol(c_preamp, L3_PREAMP, sizeof c_preamp);
    cudaMemcpyToSymbol(c_linbits, L3_LINBITS, sizeof c_linbits);
    cudaMemcpyToSymbol(c_sel2book, L3_SEL2BOOK, sizeof c_sel2book);
}

// ---- descriptor field access ------------------------------------------------------------------------
struct Desc {
    uint32_t bit_start, w1, w2, w3;
    __device__ __forceinline__ int part23() const { return w1 & 0xFFF; }
    __device__ __forceinline__ int big_values() const { return (w1 >> 12) & 0x1FF; }
    __device__ __forceinline__ int global_gain() const { return (w1 >> 21) & 0xFF; }
    __device__ __forceinline__ int block_type() const { return (w1 >> 29) & 3; }
    __device__ __forceinline__ int mixed() const { return w1 >> 31; }
    __device__ __forceinline__ int scalefac_compress() const { return w2 & 0x1FF; }
    __device__ __forceinline__ int table_select(int r) const { return (w2 >> (9 + 5 * r)) & 31; }
    __device__ __forceinline__ int preflag() const { return (w2 >> 24) & 1; }
    __device__ __forceinline__ int scalefac_scale() const { return (w2 >> 25) & 1; }
    __device__ __forceinline__ int count1_table() const { return (w2 >> 26) & 1; }
    __device__ __forceinline__ int scfsi() const { return (w2 >> 27) & 15; }
    __device__ __forceinline__ int second_granule() const { return w2 >> 31; }
    __device__ __forceinline__ int region1_start() const { return (w3 & 0x1FF) * 2; }
    __device__ __forceinline__ int region2_start() const { return ((w3 >> 9) & 0x1FF) * 2; }
    __device__ __forceinline__ int subblock_gain(int i) const { return (w3 >> (18 + 3 * i)) & 7; }
    __device__ __forceinline__ int hdr_bits() const { return (w3 >> 27) & 15; }  // header byte 3 >> 4
    __device__ __forceinline__ int reset_before() const { return w3 >> 31; }
    // 0 long, 1 short, 2 mixed
    __device__ __forceinline__ int kind() const { return block_type() == 2 ? (mixed() ? 2 : 1) : 0; }
};

__device__ __forceinline__ Desc load_desc(const l3b_grch_desc_t* p) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    Desc d;
    d.bit_start = v.x; d.w1 = v.y; d.w2 = v.z; d.w3 = v.w;
    return d;
}

// =====================================================================================================
// Entropy kernel
// =====================================================================================================

// MSB-first reader over 32-bit words of the stream's main-data blob.  cache holds `nbits` valid bits,
// left aligned; the next word to load is `next`.  Reads past the blob return zero bits.
struct BitCursor {
    const uint32_t* words;
    uint32_t nwords, next, pos;
    uint64_t cache;
    int nbits;
    __device__ __forceinline__ uint32_t ldw(uint32_t i) const {
        return i < nwords ? __byte_perm(__ldg(words + i), 0, 0x0123) : 0u;
    }
    __device__ __forceinline__ void init(const uint32_t* w, uint32_t nw, uint32_t bitpos) {
        words = w; nwords = nw; pos = bitpos;
        uint32_t wi = bitpos >> 5, off = bitpos & 31;
        cache = (((uint64_t)ldw(wi) << 32) | ldw(wi + 1)) << off;
        nbits = 64 - (int)off;
        next = wi + 2;
    }
    __device__ __forceinline__ void refill() {
        if (nbits <= 32) {
            cache |= (uint64_t)ldw(next++) << (32 - nbits);
            nbits += 32;
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)(cache >> (64 - n)); }  // 1..32
    __device__ __forceinline__ void skip(int n) { cache <<= n; nbits -= n; pos += n; }
    __device__ __forceinline__ uint32_t get(int n) {  // n >= 1
        refill();
        uint32_t v = peek(n);
        skip(n);
        return v;
    }
};

__device__ __forceinline__ uint32_t peek_bits_at(const uint32_t* words, uint32_t nwords, uint32_t bitpos, int n) {
    uint32_t wi = bitpos >> 5, off = bitpos & 31;
    uint32_t a = wi < nwords ? __byte_perm(__ldg(words + wi), 0, 0x0123) : 0u;
    uint32_t b = wi + 1 < nwords ? __byte_perm(__ldg(words + wi + 1), 0, 0x0123) : 0u;
    uint64_t v = (((uint64_t)a << 32) | b) << off;
    return (uint32_t)(v >> (64 - n));
}

__device__ __forceinline__ uint32_t find_stream(const l3b_stream_desc_t* streams, uint32_t n, uint64_t gi) {
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi) {  // last stream with first_grch <= gi
        uint32_t mid = (lo + hi + 1) >> 1;
        if (streams[mid].first_grch <= gi) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(128) l3_entropy_kernel(BatchParams p) {
    extern __shared__ uint16_t s_lut[];  // huff entries, then 128 bytes of count1
    uint8_t* s_c1 = reinterpret_cast<uint8_t*>(s_lut + ((p.t.huff_entries + 7) & ~7u));
    for (uint32_t i = threadIdx.x; i < p.t.huff_entries; i += blockDim.x) s_lut[i] = p.t.huff[i];
    for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) s_c1[i] = p.t.count1[i];
    __syncthreads();

    const uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= p.n_grch) return;
    const uint32_t si = find_stream(p.streams, p.n_streams, gi);
    const l3b_stream_desc_t* S = p.streams + si;
    const int nch = S->nch;
    const int ch = (int)((gi - S->first_grch) % (uint64_t)nch);
    const bool mpeg1 = S->mpeg1 != 0;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(p.blob + S->maindata_off);
    const uint32_t nwords = (S->maindata_bytes >> 2) + 4;  // the batch blob keeps >= 16 zero bytes after each stream

    const Desc d = load_desc(p.grch + gi);
    BitCursor br;
    br.init(words, nwords, d.bit_start);
    const uint32_t limit = d.bit_start + (uint32_t)d.part23();

    // ---------------- scalefactors (minimp3.d:613-644, 659-712) ----------------
    uint8_t* rec = p.sf + gi * kSfRecBytes;
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4* r4 = reinterpret_cast<uint4*>(rec);
#pragma unroll
        for (int i = 0; i < kSfRecBytes / 16; i++) r4[i] = z;
    }
    const int kind = d.kind();
    const int n_long = kind == 0 ? 22 : (kind == 1 ? 0 : (mpeg1 ? 8 : 6));
    const int n_short = kind == 0 ? 0 : (kind == 1 ? 39 : 30);
    const uint8_t* part = c_partitions + 28 * (kind == 0 ? 0 : (kind == 2 ? 1 : 2));
    const int scf_shift = d.scalefac_scale() + 1;
    uint32_t slen = 0;  // four byte-sized lengths
    int scfsi = d.scfsi();
    const int istereo = d.hdr_bits() & 1;
    if (mpeg1) {
        int pp = c_scfc_decode[d.scalefac_compress() & 15];
        uint32_t a = (uint32_t)(pp >> 2), b = (uint32_t)(pp & 3);
        slen = a | (a << 8) | (b << 16) | (b << 24);
    } else {
        int ist = (istereo && ch) ? 1 : 0;
        int sfc = d.scalefac_compress() >> ist;
        int k = ist * 12;
        for (;; k += 4) {
            int modprod = 1;
            slen = 0;
#pragma unroll
            for (int i = 3; i >= 0; i--) {
                int m = c_lsf_mod[k + i];
                slen |= (uint32_t)(sfc / modprod % m) << (8 * i);
                modprod *= m;
            }
            sfc -= modprod;
            if (sfc < 0) break;
        }
        part += k + 4;  // the reference's for-loop increments k once more before its exit test (minimp3.d:683-691)
        scfsi = -16;
    }
    // granule-0 scalefactors for scfsi copies (MPEG-1 granule 1 only; both granules are long blocks then)
    uint32_t g0_slen = 0, g0_bits = 0;
    if (scfsi > 0 && d.second_granule() && gi >= S->first_grch + (uint64_t)nch) {
        const Desc d0 = load_desc(p.grch + gi - nch);
        int pp = c_scfc_decode[d0.scalefac_compress() & 15];
        uint32_t a = (uint32_t)(pp >> 2), b = (uint32_t)(pp & 3);
        g0_slen = a | (a << 8) | (b << 16) | (b << 24);
        g0_bits = d0.bit_start;
    } else if (scfsi > 0) {
        scfsi = 0;  // no granule 0 to copy from
    }
    {
        const int sbg_sh = 3 - scf_shift;
        int n = 0;
        uint32_t g0_off = g0_bits;
        for (int i = 0; i < 4; i++) {
            const int cnt = part[i];
            if (!cnt) break;
            const int bits = (slen >> (8 * i)) & 0xFF;
            const int bits0 = (g0_slen >> (8 * i)) & 0xFF;
            const bool copy = (scfsi & 8) != 0;
            for (int k = 0; k < cnt; k++, n++) {
                int s, ip;
                if (copy) {
                    s = bits0 ? (int)peek_bits_at(words, nwords, g0_off + (uint32_t)(k * bits0), bits0) : 0;
                    ip = s;
                } else if (!bits) {
                    s = 0; ip = 0;
                } else {
                    s = (int)br.get(bits);
                    ip = (scfsi < 0 && s == (1 << bits) - 1) ? 255 : s;
                }
                int adj = 0;
                if (n_short) { if (n >= n_long) adj = d.subblock_gain((n - n_long) % 3) << sbg_sh; }
                else if (d.preflag() && n >= 11 && n < 21) adj = c_preamp[n - 11];
                rec[n] = (uint8_t)(s + adj);
                rec[40 + n] = (uint8_t)ip;
            }
            g0_off += (uint32_t)(cnt * bits0);
            scfsi *= 2;
        }
        for (int j = 0; j < 3 && n < 40; j++, n++) {  // scf[0] = scf[1] = scf[2] = 0 after the last partition
            int adj = 0;
            if (n_short) { if (n >= n_long && n < n_long + n_short) adj = d.subblock_gain((n - n_long) % 3) << sbg_sh; }
            rec[n] = (uint8_t)adj;
        }
    }

    // ---------------- Huffman (minimp3.d:748-883), values only ----------------
    uint4* outp = p.is + gi * kIsChunks;
    uint4 q = make_uint4(0, 0, 0, 0);
    int idx = 0;
#define L3_EMIT_PAIR(v0, v1)                                                                   \
    do {                                                                                       \
        uint32_t pk_ = ((uint32_t)(v0) & 0xFFFFu) | ((uint32_t)(v1) << 16);                    \
        int pp_ = (idx >> 1) & 3;                                                              \
        if (pp_ == 0) q.x = pk_; else if (pp_ == 1) q.y = pk_; else if (pp_ == 2) q.z = pk_;   \
        else { q.w = pk_; outp[idx >> 3] = q; q = make_uint4(0, 0, 0, 0); }                    \
        idx += 2;                                                                              \
    } while (0)

    const int bv_end = 2 * d.big_values();
    for (int r = 0; r < 3 && idx < bv_end; r++) {
        int rend = r == 0 ? d.region1_start() : (r == 1 ? d.region2_start() : 576);
        if (rend > bv_end) rend = bv_end;
        const int sel = d.table_select(r);
        const int book = c_sel2book[sel] < 0 ? L3_NBOOKS : c_sel2book[sel];
        const int linbits = c_linbits[sel];
        const uint32_t base = p.t.huff_base[book];
        const int rootw = p.t.huff_root[book];
        while (idx < rend) {
            br.refill();
            int w = rootw;
            uint32_t e = s_lut[base + br.peek(w)];
            while (e & 0x8000u) {
                br.skip(w);
                w = (int)((e >> 12) & 7) + 1;
                e = s_lut[base + (e & 0xFFFu) + br.peek(w)];
            }
            br.skip((int)((e >> 8) & 15));
            int a0 = (int)(e & 15), a1 = (int)((e >> 4) & 15);
            if (linbits && a0 == 15) { br.refill(); a0 += (int)br.peek(linbits); br.skip(linbits); }
            if (a0) { if (br.peek(1)) a0 = -a0; br.skip(1); }
            if (linbits && a1 == 15) { br.refill(); a1 += (int)br.peek(linbits); br.skip(linbits); }
            if (a1) { if (br.peek(1)) a1 = -a1; br.skip(1); }
            L3_EMIT_PAIR(a0, a1);
        }
    }
    {
        const uint8_t* c1 = s_c1 + 64 * d.count1_table();
        for (;;) {
            br.refill();
            uint32_t e = c1[br.peek(6)];
            br.skip((int)(e & 15));
            if (br.pos > limit) break;  // tested after the code, before the signs (minimp3.d:866)
            if (idx >= 576) break;      // sfb terminator (minimp3.d:873)
            int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
            if (e & 0x80) { v0 = br.peek(1) ? -1 : 1; br.skip(1); }
            if (e & 0x40) { v1 = br.peek(1) ? -1 : 1; br.skip(1); }
            L3_EMIT_PAIR(v0, v1);
            if (idx >= 576) break;      // (minimp3.d:876)
            if (e & 0x20) { v2 = br.peek(1) ? -1 : 1; br.skip(1); }
            if (e & 0x10) { v3 = br.peek(1) ? -1 : 1; br.skip(1); }
            L3_EMIT_PAIR(v2, v3);
        }
    }
#undef L3_EMIT_PAIR
    int chunks = (idx + 7) >> 3;
    if ((idx >> 1) & 3) outp[idx >> 3] = q;
    if (p.zero_fill)
        for (int c = chunks; c < kIsChunks; c++) outp[c] = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint16_t*>(rec + 80) = (uint16_t)chunks;
}

// =====================================================================================================
// Granule kernel
// =====================================================================================================

// L3_ldexp_q2 (minimp3.d:646-657): y * 2^(-exp_q2/4) by repeated multiplication, same rounding steps.
__device__ __forceinline__ float ldexp_q2(float y, int exp_q2) {
    int e;
    do {
        e = exp_q2 < 120 ? exp_q2 : 120;
        y *= c_expfrac[e & 3] * (float)((1 << 30) >> (e >> 2));
    } while ((exp_q2 -= e) > 0);
    return y;
}

// L3_pow_43 for x >= 129 (minimp3.d:737-745)
__device__ __noinline__ float pow43_big(const float* pow43, int x) {
    int mult = 256;
    if (x < 1024) { mult = 16; x <<= 3; }
    int sign = 2 * x & 64;
    float frac = __fdiv_rn((float)((x & 63) - sign), (float)((x & ~63) + sign));
    return pow43[(x + sign) >> 6] * (1.0f + frac * ((4.0f / 3) + frac * (2.0f / 9))) * (float)mult;
}

__device__ __forceinline__ float requant(const float* pow43, int v, float s) {
    int a = v < 0 ? -v : v;
    float pw = a < 129 ? pow43[a] : pow43_big(pow43, a);
    float r = pw * s;
    return v < 0 ? -r : r;
}

// L3_dct3_9 (minimp3.d:1022-1060), in registers
__device__ __forceinline__ void dct3_9(float* y) {
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

// L3_imdct36 for one band (minimp3.d:1062-1100): x -> out, overlap updated in place
__device__ __forceinline__ void imdct36_band(const float* x, float* ovl, int wsel, float* out) {
    float co[9], si[9];
    co[0] = -x[0];
    si[0] = x[17];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        si[8 - 2 * i] = x[4 * i + 1] - x[4 * i + 2];
        co[1 + 2 * i] = x[4 * i + 1] + x[4 * i + 2];
        si[7 - 2 * i] = x[4 * i + 4] - x[4 * i + 3];
        co[2 + 2 * i] = -(x[4 * i + 3] + x[4 * i + 4]);
    }
    dct3_9(co);
    dct3_9(si);
    si[1] = -si[1];
    si[3] = -si[3];
    si[5] = -si[5];
    si[7] = -si[7];
    const float* window = c_mdctw + 18 * wsel;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        float o = ovl[i];
        float sum = co[i] * c_twid9[9 + i] + si[i] * c_twid9[0 + i];
        ovl[i] = co[i] * c_twid9[0 + i] - si[i] * c_twid9[9 + i];
        out[i] = o * window[0 + i] - sum * window[9 + i];
        out[17 - i] = o * window[9 + i] + sum * window[0 + i];
    }
}

// L3_idct3 / L3_imdct12 (minimp3.d:1102-1129); X(k) = x[OFF + 3k]
template <int OFF>
__device__ __forceinline__ void imdct12(const float* x, float* dst, float* overlap) {
    float co[3], si[3];
    {
        float x0 = -x[OFF + 0], x1 = x[OFF + 6] + x[OFF + 3], x2 = x[OFF + 12] + x[OFF + 9];
        float m1 = x1 * 0.86602540f, a1 = x0 - x2 * 0.5f;
        co[1] = x0 + x2; co[0] = a1 + m1; co[2] = a1 - m1;
    }
    {
        float x0 = x[OFF + 15], x1 = x[OFF + 12] - x[OFF + 9], x2 = x[OFF + 6] - x[OFF + 3];
        float m1 = x1 * 0.86602540f, a1 = x0 - x2 * 0.5f;
        si[1] = x0 + x2; si[0] = a1 + m1; si[2] = a1 - m1;
    }
    si[1] = -si[1];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float o = overlap[i];
        float sum = co[i] * c_twid3[3 + i] + si[i] * c_twid3[0 + i];
        overlap[i] = co[i] * c_twid3[0 + i] - si[i] * c_twid3[3 + i];
        dst[i] = o * c_twid3[2 - i] - sum * c_twid3[5 - i];
        dst[5 - i] = o * c_twid3[5 - i] + sum * c_twid3[2 - i];
    }
}

// L3_imdct_short for one band (minimp3.d:1131-1142)
__device__ __forceinline__ void imdct_short_band(const float* x, float* ovl, float* out) {
#pragma unroll
    for (int i = 0; i < 6; i++) out[i] = ovl[i];
    imdct12<0>(x, out + 6, ovl + 6);
    imdct12<1>(x, out + 12, ovl + 6);
    float nd[6];
    imdct12<2>(x, nd, ovl + 6);
#pragma unroll
    for (int i = 0; i < 6; i++) ovl[i] = nd[i];
}

// address of history row for absolute slot tt (>= 0) of one channel: parity array tt&1, row (tt>>1) mod 18
__device__ __forceinline__ const float* drow(const float* Dch, int tt) {
    int r = tt >> 1;
    r = r >= 36 ? r - 36 : (r >= 18 ? r - 18 : r);
    return Dch + (tt & 1) * kDParity + r * 33;
}

template <int NCH>
__global__ void __launch_bounds__(32 * NCH) l3_granule_kernel(BatchParams p, const Tile* tiles, uint32_t n_tiles) {
    __shared__ __align__(16) float s_xr[NCH][kXrStride];
    __shared__ __align__(16) float s_D[NCH][2 * kDParity];
    __shared__ float s_scf[NCH][40];
    __shared__ float s_pow43[132];
    __shared__ uint8_t s_sfbpair[3][288];
    __shared__ uint8_t s_sfbw[3][40];
    __shared__ uint16_t s_sfbo[3][40];
    __shared__ uint8_t s_ist[40];
    __shared__ uint8_t s_smode[40];
    __shared__ float s_kl[40], s_kr[40];
    __shared__ int s_maxband[3];

    if (blockIdx.x >= n_tiles) return;
    const Tile T = tiles[blockIdx.x];
    const l3b_stream_desc_t S = p.streams[T.stream];
    const int tid = threadIdx.x, nthr = 32 * NCH;
    const int ch = tid >> 5, lane = tid & 31;
    const bool mpeg1 = S.mpeg1 != 0;
    const int row = S.sr_idx;

    for (int i = tid; i < 129; i += nthr) s_pow43[i] = p.t.pow43[i];
    for (int i = tid; i < 3 * 288; i += nthr) (&s_sfbpair[0][0])[i] = p.t.sfb_of_pair[row * 3 * 288 + i];
    for (int i = tid; i < 3 * 40; i += nthr) {
        (&s_sfbw[0][0])[i] = p.t.sfb_width[row * 120 + i];
        (&s_sfbo[0][0])[i] = p.t.sfb_start[row * 120 + i];
    }
    for (int i = lane; i < 2 * kDParity; i += 32) s_D[ch][i] = 0.0f;

    // synthesis window weights of this lane: inner index ii = lane & 15 (0..14 active)
    const int ii = lane & 15, par = lane >> 4;
    float w0[8], w1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        w0[k] = ii < 15 ? __ldg(p.t.win + (k * 2 + 0) * 15 + ii) : 0.0f;
        w1[k] = ii < 15 ? __ldg(p.t.win + (k * 2 + 1) * 15 + ii) : 0.0f;
    }

    // recompute halo: up to two granules before the tile, unless decoder state was zeroed in between
    int start = (int)T.g0;
    for (int depth = 0; depth < 2 && start > 0; depth++) {
        const l3b_grch_desc_t* dp = p.grch + S.first_grch + (uint64_t)start * NCH;
        if (__ldg(&dp->w3) >> 31) break;
        start--;
    }
    float ovl[9];
#pragma unroll
    for (int i = 0; i < 9; i++) ovl[i] = 0.0f;
    __syncthreads();

    const float* Dch = s_D[ch];
    float* xr = s_xr[ch];
    const int n_long_bands_mixed = 2 << (row == 1 ? 1 : 0);  // minimp3.d:1218
    int dgc = 0;                                             // granules that went through the DCT stage
    const int g_end = (int)(T.g0 + T.ng);

    for (int g = start; g < g_end; g++) {
        const uint64_t di = S.first_grch + (uint64_t)g * NCH + ch;
        const Desc d = load_desc(p.grch + di);
        const int mode = g >= (int)T.g0 ? 2 : (g == (int)T.g0 - 1 ? 1 : 0);
        if (d.reset_before() && g != start) {
#pragma unroll
            for (int i = 0; i < 9; i++) ovl[i] = 0.0f;
            for (int i = lane; i < 2 * kDParity; i += 32) s_D[ch][i] = 0.0f;
            __syncwarp();
        }
        const int kind = d.kind();
        const int hb = d.hdr_bits();
        const bool ms_frame = (hb & 0xE) == 0x6;                 // HDR_IS_MS_STEREO
        const bool istereo = NCH == 2 && (hb & 1);               // HDR_TEST_I_STEREO
        const int n_long_sfb = kind == 0 ? 22 : (kind == 1 ? 0 : (mpeg1 ? 8 : 6));
        const int n_sfb = n_long_sfb + (kind == 0 ? 0 : (kind == 1 ? 39 : 30));
        const uint8_t* rec = p.sf + di * kSfRecBytes;

        // ---------------- band gains (minimp3.d:714-719) ----------------
        {
            const int gain_exp = d.global_gain() - 4 - 210 - (ms_frame ? 2 : 0);
            const float gain = ldexp_q2(2048.0f, 44 - gain_exp);
            const int scf_shift = d.scalefac_scale() + 1;
            for (int i = lane; i < 40; i += 32) {
                float v = 0.0f;
                if (i < n_sfb) v = ldexp_q2(gain, (int)__ldg(rec + i) << scf_shift);
                s_scf[ch][i] = v;
            }
            if (NCH == 2 && ch == 1 && istereo)
                for (int i = lane; i < 40; i += 32) s_ist[i] = __ldg(rec + 40 + i);
        }
        __syncwarp();

        // ---------------- requantisation (minimp3.d:813-816, 846, 874-878) ----------------
        {
            const int nchunks = *reinterpret_cast<const uint16_t*>(rec + 80);
            const uint32_t* isw = reinterpret_cast<const uint32_t*>(p.is + di * kIsChunks);
#pragma unroll
            for (int m = 0; m < 9; m++) {
                const int pi = lane + 32 * m;
                uint32_t v = (pi >> 2) < nchunks ? __ldg(isw + pi) : 0u;
                const float s = s_scf[ch][s_sfbpair[kind][pi]];
                const int v0 = (int)(int16_t)(v & 0xFFFFu), v1 = (int)(int16_t)(v >> 16);
                float2 o;
                o.x = requant(s_pow43, v0, s);
                o.y = requant(s_pow43, v1, s);
                *reinterpret_cast<float2*>(xr + 2 * pi) = o;
            }
        }

        // ---------------- stereo (minimp3.d:885-982, 1207-1213) ----------------
        if (NCH == 2) {
            __syncthreads();
            float* L = s_xr[0];
            float* R = s_xr[1];
            if (istereo) {
                // L3_intensity_stereo works on channel 0's band layout (gr_info of ch 0, minimp3.d:1209) and on
                // channel 1's scalefac_compress (gr[1], minimp3.d:981)
                const Desc d0 = load_desc(p.grch + S.first_grch + (uint64_t)g * NCH + 0);
                const Desc d1 = load_desc(p.grch + S.first_grch + (uint64_t)g * NCH + 1);
                const int kind0 = d0.kind();
                const int n_long_sfb0 = kind0 == 0 ? 22 : (kind0 == 1 ? 0 : (mpeg1 ? 8 : 6));
                const int n_sfb0 = n_long_sfb0 + (kind0 == 0 ? 0 : (kind0 == 1 ? 39 : 30));
                if (ch == 0) {
                    // L3_stereo_top_band: last sfb (per window) of the right channel holding a non-zero value
                    int mb0 = -1, mb1 = -1, mb2 = -1;
                    for (int i = lane; i < n_sfb0; i += 32) {
                        const int off = s_sfbo[kind0][i], wdt = s_sfbw[kind0][i];
                        bool nz = false;
                        for (int k = 0; k < wdt; k++) nz |= (R[off + k] != 0.0f);
                        if (nz) { int c = i % 3; if (c == 0) mb0 = max(mb0, i); else if (c == 1) mb1 = max(mb1, i); else mb2 = max(mb2, i); }
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) {
                        mb0 = max(mb0, __shfl_xor_sync(0xffffffffu, mb0, sft));
                        mb1 = max(mb1, __shfl_xor_sync(0xffffffffu, mb1, sft));
                        mb2 = max(mb2, __shfl_xor_sync(0xffffffffu, mb2, sft));
                    }
                    if (n_long_sfb0) mb0 = mb1 = mb2 = max(max(mb0, mb1), mb2);
                    if (lane == 0) {
                        const int max_blocks = kind0 == 0 ? 1 : 3;
                        const int default_pos = mpeg1 ? 3 : 0;
                        const int mb[3] = {mb0, mb1, mb2};
                        for (int i = 0; i < max_blocks; i++) {
                            int itop = n_sfb0 - max_blocks + i, prev = itop - max_blocks;
                            s_ist[itop] = (uint8_t)(mb[i] >= prev ? default_pos : s_ist[prev]);
                        }
                        s_maxband[0] = mb0; s_maxband[1] = mb1; s_maxband[2] = mb2;
                    }
                    __syncwarp();
                    // L3_stereo_process: per-sfb decision and gains
                    const unsigned max_pos = mpeg1 ? 7u : 64u;
                    const int mpeg2_sh = d1.scalefac_compress() & 1;
                    for (int i = lane; i < n_sfb0; i += 32) {
                        const unsigned ipos = s_ist[i];
                        uint8_t md = 0;
                        if (i > s_maxband[i % 3] && ipos < max_pos) {
                            float kl, kr, s = (hb & 2) ? 1.41421356f : 1.0f;
                            if (mpeg1) {
                                kl = c_pan[2 * ipos];
                                kr = c_pan[2 * ipos + 1];
                            } else {
                                kl = 1.0f;
                                kr = ldexp_q2(1.0f, (int)((ipos + 1) >> 1 << mpeg2_sh));
                                if (ipos & 1) { kl = kr; kr = 1.0f; }
                            }
                            s_kl[i] = kl * s;
                            s_kr[i] = kr * s;
                            md = 1;
                        } else if (hb & 2) {
                            md = 2;
                        }
                        s_smode[i] = md;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int m = 0; m < 9; m++) {
                    const int k = ch * 288 + lane + 32 * m;
                    const int sfb = s_sfbpair[kind0][k >> 1];
                    const int md = s_smode[sfb];
                    const float a = L[k], b = R[k];
                    if (md == 1) { R[k] = a * s_kr[sfb]; L[k] = a * s_kl[sfb]; }
                    else if (md == 2) { L[k] = a + b; R[k] = a - b; }
                }
            } else if (ms_frame) {
#pragma unroll
                for (int m = 0; m < 9; m++) {
                    const int k = ch * 288 + lane + 32 * m;
                    const float a = L[k], b = R[k];
                    L[k] = a + b;
                    R[k] = a - b;
                }
            }
            __syncthreads();
        } else {
            __syncwarp();
        }

        // ---------------- reorder + antialias + IMDCT + frequency inversion (minimp3.d:1215-1229) ----------
        {
            float x[18], y[18];
            const int nlb = kind == 2 ? n_long_bands_mixed : 0;
            if (kind == 0) {
#pragma unroll
                for (int i = 0; i < 18; i++) x[i] = xr[lane * 18 + i];
            } else {
                const uint16_t* pm = p.t.perm + (row * 2 + (kind == 2 ? 1 : 0)) * 576 + lane * 18;
#pragma unroll
                for (int i = 0; i < 18; i++) x[i] = xr[__ldg(pm + i)];
            }
            const int aa_bands = kind == 0 ? 31 : nlb - 1;
            if (aa_bands > 0) {
                const bool lower = lane >= 1 && lane - 1 < aa_bands;
                const bool upper = lane < aa_bands;
                float nlo[8], nhi[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float dn = __shfl_up_sync(0xffffffffu, x[17 - i], 1);   // band-1, element 17-i
                    const float up = __shfl_down_sync(0xffffffffu, x[i], 1);      // band+1, element i
                    nlo[i] = x[i] * c_aa[i] - dn * c_aa[8 + i];
                    nhi[i] = up * c_aa[8 + i] + x[17 - i] * c_aa[i];
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (lower) x[i] = nlo[i];
                    if (upper) x[17 - i] = nhi[i];
                }
            }
            if (d.block_type() == 2 && lane >= nlb) imdct_short_band(x, ovl, y);
            else imdct36_band(x, ovl, (d.block_type() == 3) ? 1 : 0, y);
            if (mode >= 1) {
                if (lane & 1) {
#pragma unroll
                    for (int i = 1; i < 18; i += 2) y[i] = -y[i];
                }
                __syncwarp();  // every lane has consumed its inputs; the buffer is reused in the padded layout
#pragma unroll
                for (int i = 0; i < 18; i++) xr[lane * 19 + i] = y[i];
            }
        }
        if (mode == 0) { __syncwarp(); continue; }
        __syncwarp();

        // ---------------- DCT-32 matrixing across bands, one time slot per lane (minimp3.d:1232-1298) -------
        const int hbase = 18 * (dgc & 1) + 36;  // absolute slot of this granule's slot 0 in the 36-slot ring (+36)
        if (lane < 18) {
            float t[4][8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float x0 = xr[i * 19 + lane];
                float x1 = xr[(15 - i) * 19 + lane];
                float x2 = xr[(16 + i) * 19 + lane];
                float x3 = xr[(31 - i) * 19 + lane];
                float t0 = x0 + x3;
                float t1 = x1 + x2;
                float t2 = (x1 - x2) * c_sec[3 * i + 0];
                float t3 = (x0 - x3) * c_sec[3 * i + 1];
                t[0][i] = t0 + t1;
                t[1][i] = (t0 - t1) * c_sec[3 * i + 2];
                t[2][i] = t3 + t2;
                t[3][i] = (t3 - t2) * c_sec[3 * i + 2];
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                float x0 = t[r][0], x1 = t[r][1], x2 = t[r][2], x3 = t[r][3], x4 = t[r][4], x5 = t[r][5], x6 = t[r][6], x7 = t[r][7], xt;
                xt = x0 - x7; x0 += x7;
                x7 = x1 - x6; x1 += x6;
                x6 = x2 - x5; x2 += x5;
                x5 = x3 - x4; x3 += x4;
                x4 = x0 - x3; x0 += x3;
                x3 = x1 - x2; x1 += x2;
                t[r][0] = x0 + x1;
                t[r][4] = (x0 - x1) * 0.70710677f;
                x5 = x5 + x6;
                x6 = (x6 + x7) * 0.70710677f;
                x7 = x7 + xt;
                x3 = (x3 + x4) * 0.70710677f;
                x5 -= x7 * 0.198912367f;
                x7 += x5 * 0.382683432f;
                x5 -= x7 * 0.198912367f;
                x0 = xt - x6; xt += x6;
                t[r][1] = (xt + x7) * 0.50979561f;
                t[r][2] = (x4 + x3) * 0.54119611f;
                t[r][3] = (x0 - x5) * 0.60134488f;
                t[r][5] = (x0 + x5) * 0.89997619f;
                t[r][6] = (x4 - x3) * 1.30656302f;
                t[r][7] = (xt - x7) * 2.56291556f;
            }
            float* out = const_cast<float*>(drow(Dch, hbase + lane));
#pragma unroll
            for (int i = 0; i < 7; i++) {
                out[4 * i + 0] = t[0][i];
                out[4 * i + 1] = t[2][i] + t[3][i] + t[3][i + 1];
                out[4 * i + 2] = t[1][i] + t[1][i + 1];
                out[4 * i + 3] = t[2][i + 1] + t[3][i] + t[3][i + 1];
            }
            out[28] = t[0][7];
            out[29] = t[2][7] + t[3][7];
            out[30] = t[1][7];
            out[31] = t[3][7];
        }
        dgc++;
        __syncwarp();
        if (mode == 1) continue;

        // ---------------- 512-tap window (minimp3.d:1305-1406) ----------------
        {
            const uint64_t gbase = (uint64_t)g * 576u * NCH;  // interleaved sample index of this granule's first sample
            float* pcm = p.pcm + S.pcm_off;
            const uint64_t skip = S.pcm_skip, count = S.pcm_count;
            const float scale = 1.0f / 32768.0f;
            // main part: lane (par, ii) produces samples 15-ii and 17+ii of slots s = 2q + par
            if (ii < 15) {
                float V[32];
                // V[j] = D[slot (j - 15 + par)][ j odd ? 31-ii : 1+ii ]   (rows of the reference's zlin)
#pragma unroll
                for (int j = 0; j < 16; j++)
                    V[j] = drow(Dch, hbase + par + j - 15)[(j & 1) ? 31 - ii : 1 + ii];
#pragma unroll
                for (int q = 0; q < 9; q++) {
                    if (q > 0) {
                        V[2 * q + 14] = drow(Dch, hbase + par + 2 * q - 1)[1 + ii];
                        V[2 * q + 15] = drow(Dch, hbase + par + 2 * q)[31 - ii];
                    }
                    float a, b;
                    {
                        const float vz = V[2 * q + 15], vy = V[2 * q + 0];
                        b = vz * w1[0] + vy * w0[0];
                        a = vz * w0[0] - vy * w1[0];
                    }
#pragma unroll
                    for (int k = 1; k < 8; k++) {
                        const float vz = V[2 * q + 15 - k], vy = V[2 * q + k];
                        b += vz * w1[k] + vy * w0[k];
                        if (k & 1) a += vy * w1[k] - vz * w0[k];
                        else a += vz * w0[k] - vy * w1[k];
                    }
                    const int s = 2 * q + par;
                    const uint64_t ea = gbase + (uint64_t)((32 * s + 15 - ii) * NCH + ch);
                    const uint64_t eb = gbase + (uint64_t)((32 * s + 17 + ii) * NCH + ch);
                    if (ea >= skip && ea - skip < count) pcm[ea - skip] = a * scale;
                    if (eb >= skip && eb - skip < count) pcm[eb - skip] = b * scale;
                }
            }
            // samples 0 and 16 of every slot (mp3d_synth_pair), one slot per lane
            if (lane < 18) {
                float z[15];
#pragma unroll
                for (int k = 0; k < 15; k++) z[k] = drow(Dch, hbase + lane - 15 + k)[16];
                float a;
                a = (z[14] - z[0]) * 29.0f;
                a += (z[1] + z[13]) * 213.0f;
                a += (z[12] - z[2]) * 459.0f;
                a += (z[3] + z[11]) * 2037.0f;
                a += (z[10] - z[4]) * 5153.0f;
                a += (z[5] + z[9]) * 6574.0f;
                a += (z[8] - z[6]) * 37489.0f;
                a += z[7] * 75038.0f;
                const uint64_t e0 = gbase + (uint64_t)((32 * lane) * NCH + ch);
                if (e0 >= skip && e0 - skip < count) pcm[e0 - skip] = a * scale;
#pragma unroll
                for (int k = 0; k < 15; k += 2) z[k] = drow(Dch, hbase + lane - 15 + k)[0];
                a = z[14] * 104.0f;
                a += z[12] * 1567.0f;
                a += z[10] * 9727.0f;
                a += z[8] * 64019.0f;
                a += z[6] * -9975.0f;
                a += z[4] * -45.0f;
                a += z[2] * 146.0f;
                a += z[0] * -5.0f;
                const uint64_t e16 = gbase + (uint64_t)((32 * lane + 16) * NCH + ch);
                if (e16 >= skip && e16 - skip < count) pcm[e16 - skip] = a * scale;
            }
        }
        __syncwarp();
    }
}

template __global__ void l3_granule_kernel<1>(BatchParams, const Tile*, uint32_t);
template __global__ void l3_granule_kernel<2>(BatchParams, const Tile*, uint32_t);

void launch_entropy(const BatchParams& p, cudaStream_t s) {
    if (!p.n_grch) return;
    size_t smem = (size_t)((p.t.huff_entries + 7) & ~7u) * 2 + 128;
    unsigned blocks = (unsigned)((p.n_grch + 127) / 128);
    l3_entropy_kernel<<<blocks, 128, smem, s>>>(p);
}

void launch_granule(const BatchParams& p, const Tile* tiles_stereo, uint32_t n_stereo, const Tile* tiles_mono,
                    uint32_t n_mono, cudaStream_t s, cudaEvent_t ev_mid) {
    if (n_stereo) l3_granule_kernel<2><<<n_stereo, 64, 0, s>>>(p, tiles_stereo, n_stereo);
    if (ev_mid) cudaEventRecord(ev_mid, s);
    if (n_mono) l3_granule_kernel<1><<<n_mono, 32, 0, s>>>(p, tiles_mono, n_mono);
}

}  // namespace l3b
