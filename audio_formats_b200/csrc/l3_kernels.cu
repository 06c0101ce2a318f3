// l3_kernels.cu -- see l3_kernels.cuh.  Compile with -fmad=false (bit-exactness contract).
#include "l3_kernels.cuh"

#include <cstdlib>

#include "l3_desc.cuh"
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
// (1.0f, 1.0f), written at context creation.  ptxas cannot know its value, so FFMA2(a, c_one2, b) stays what it is:
// a packed add that rounds exactly like add.rn (a * 1 is exact).  See VT<2, false>.
__constant__ float2 c_one2;

void upload_constants() {
    cudaMemcpyToSymbol(c_expfrac, L3_EXPFRAC, sizeof c_expfrac);
    cudaMemcpyToSymbol(c_aa, L3_AA, sizeof c_aa);
    cudaMemcpyToSymbol(c_twid9, L3_TWID9, sizeof c_twid9);
    cudaMemcpyToSymbol(c_twid3, L3_TWID3, sizeof c_twid3);
    cudaMemcpyToSymbol(c_mdctw, L3_MDCT_WINDOW, sizeof c_mdctw);
    cudaMemcpyToSymbol(c_sec, L3_SEC, sizeof c_sec);
    cudaMemcpyToSymbol(c_pan, L3_PAN, sizeof c_pan);
    const float2 one = make_float2(1.0f, 1.0f);
    cudaMemcpyToSymbol(c_one2, &one, sizeof one);
}

// =====================================================================================================
// Granule kernel
// =====================================================================================================
//
// One warp decodes a run of consecutive granules of one stream.  For stereo streams the two channels
// travel together in every register as a float2 (T = float2) and ALL arithmetic is packed (sm_100 FMUL2 /
// FFMA2: one issue slot for both channels).  Stage mapping inside the warp:
//   requant + MS stereo     lane = coefficient pair (9 per lane)
//   antialias + IMDCT       lane = subband (overlap of the previous granule lives in registers)
//   DCT-32                  lane = time slot (18 lanes)
//   window                  lane = (slot parity, inner index i of minimp3.d:1371), 16-tap sliding register window
// Shared memory per warp: spectrum T[608], DCT history T[33 rows x 33], small tables.
//
// Two arithmetic modes, one source (template parameter FUSED):
//   FUSED = false  bit-exact: every product and every sum is rounded separately, in the reference's order, so PCM is
//                  bit-identical to the un-fused scalar reference.  ptxas 12.9 contracts a packed multiply feeding a
//                  packed add into FFMA2 even under -fmad=false and explicit .rn, so the packed ADD is written as
//                  FFMA2(a, ONE, b) with ONE = (1, 1) read from constant memory: a * 1 is exact, the sum is rounded
//                  once, denormals and signed zeros behave like add.rn, and nothing is left for ptxas to contract.
//                  Subtraction is the addition of the negated operand (a - b == a + (-b) in IEEE arithmetic, signed zeros
//                  included), and a product to be subtracted is formed with the negated weight (a * (-w) == -(a * w)).
//   FUSED = true   tolerance mode: multiply-adds of the reference are contracted into FFMA2 by hand (one rounding
//                  instead of two).  Not bit-identical; measured against the north star's 1e-5 FS / 99.99 % bar.

#ifndef L3B_RQ_UNROLL
#define L3B_RQ_UNROLL 1
#endif
// Unroll factor of the requantisation loop (9 trips).  Rolled: the loop of a granule must stay inside the 32 KB instruction
// cache that the 16 warps of an SM share (measured: unrolled 19.3 ms, rolled 18.9 ms, and only then do the phase barriers
// that round 1 needed stop paying).
constexpr int kRqUnroll = L3B_RQ_UNROLL;

template <int NCH, bool FUSED> struct VT;
template <bool FUSED> struct VT<1, FUSED> {
    typedef float T;
    static __device__ __forceinline__ T zero() { return 0.0f; }
    static __device__ __forceinline__ T add(T a, T b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ T sub(T a, T b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ T muls(T a, float s) { return __fmul_rn(a, s); }
    static __device__ __forceinline__ T mulw(T a, float w0, float) { return __fmul_rn(a, w0); }
    static __device__ __forceinline__ T neg(T a) { return -a; }
    // c + a*w, c - a*w
    static __device__ __forceinline__ T mac(T c, T a, float w) { return FUSED ? __fmaf_rn(a, w, c) : __fadd_rn(c, __fmul_rn(a, w)); }
    static __device__ __forceinline__ T msc(T c, T a, float w) { return FUSED ? __fmaf_rn(-a, w, c) : __fsub_rn(c, __fmul_rn(a, w)); }
    // a*wa + b*wb, a*wa - b*wb
    static __device__ __forceinline__ T mm_add(T a, float wa, T b, float wb) {
        return FUSED ? __fmaf_rn(b, wb, __fmul_rn(a, wa)) : __fadd_rn(__fmul_rn(a, wa), __fmul_rn(b, wb));
    }
    static __device__ __forceinline__ T mm_sub(T a, float wa, T b, float wb) {
        return FUSED ? __fmaf_rn(-b, wb, __fmul_rn(a, wa)) : __fsub_rn(__fmul_rn(a, wa), __fmul_rn(b, wb));
    }
    static __device__ __forceinline__ T shfl_up(T a) { return __shfl_up_sync(0xffffffffu, a, 1); }
    static __device__ __forceinline__ T shfl_down(T a) { return __shfl_down_sync(0xffffffffu, a, 1); }
    static __device__ __forceinline__ T sel(bool c0, bool, T a, T b) { return c0 ? a : b; }
    static __device__ __forceinline__ float ch(T a, int) { return a; }
    static __device__ __forceinline__ T pack(float a, float) { return a; }
    static __device__ __forceinline__ T flip(T a, uint32_t signmask) { return __uint_as_float(__float_as_uint(a) ^ signmask); }
};
template <bool FUSED> struct VT<2, FUSED> {
    typedef float2 T;
    static __device__ __forceinline__ T zero() { return make_float2(0.0f, 0.0f); }
    static __device__ __forceinline__ T neg(T a) { return make_float2(-a.x, -a.y); }
    static __device__ __forceinline__ T bc(float s) { return make_float2(s, s); }
    // a * m + c on both halves, m a 64-bit constant-bank value: kept as ONE 64-bit operand so that it lives in an
    // aligned (uniform) register pair instead of being re-assembled from two scalars in front of every use
    static __device__ __forceinline__ T fma_const(T a, const float2& m, T c) {
        unsigned long long r;
        asm("fma.rn.f32x2 %0, %1, %2, %3;"
            : "=l"(r)
            : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&m)),
              "l"(*reinterpret_cast<const unsigned long long*>(&c)));
        return *reinterpret_cast<T*>(&r);
    }
#ifdef L3B_EXP_SCALAR_MUL   // A/B only: round-1 arithmetic (scalar multiplies feeding packed adds)
    static __device__ __forceinline__ T add(T a, T b) { return __fadd2_rn(a, b); }
    static __device__ __forceinline__ T sub(T a, T b) { return __fadd2_rn(a, neg(b)); }
    static __device__ __forceinline__ T muls(T a, float s) { return FUSED ? __fmul2_rn(a, bc(s)) : make_float2(__fmul_rn(a.x, s), __fmul_rn(a.y, s)); }
    static __device__ __forceinline__ T mulw(T a, float w0, float w1) { return FUSED ? __fmul2_rn(a, make_float2(w0, w1)) : make_float2(__fmul_rn(a.x, w0), __fmul_rn(a.y, w1)); }
#else
    static __device__ __forceinline__ T add(T a, T b) { return FUSED ? __fadd2_rn(a, b) : fma_const(a, c_one2, b); }
    static __device__ __forceinline__ T sub(T a, T b) { return FUSED ? __fadd2_rn(a, neg(b)) : fma_const(a, c_one2, neg(b)); }   // a - b == a + (-b), exactly
    static __device__ __forceinline__ T muls(T a, float s) { return __fmul2_rn(a, bc(s)); }
    static __device__ __forceinline__ T mulw(T a, float w0, float w1) { return __fmul2_rn(a, make_float2(w0, w1)); }
#endif
    static __device__ __forceinline__ T mac(T c, T a, float w) { return FUSED ? __ffma2_rn(a, bc(w), c) : add(c, muls(a, w)); }
    static __device__ __forceinline__ T msc(T c, T a, float w) { return FUSED ? __ffma2_rn(a, bc(-w), c) : add(c, muls(a, -w)); }   // a*(-w) == -(a*w), exactly
    static __device__ __forceinline__ T mm_add(T a, float wa, T b, float wb) {
        return FUSED ? __ffma2_rn(b, bc(wb), muls(a, wa)) : add(muls(a, wa), muls(b, wb));
    }
    static __device__ __forceinline__ T mm_sub(T a, float wa, T b, float wb) {
        return FUSED ? __ffma2_rn(b, bc(-wb), muls(a, wa)) : add(muls(a, wa), muls(b, -wb));
    }
    static __device__ __forceinline__ T shfl_up(T a) {
        return make_float2(__shfl_up_sync(0xffffffffu, a.x, 1), __shfl_up_sync(0xffffffffu, a.y, 1));
    }
    static __device__ __forceinline__ T shfl_down(T a) {
        return make_float2(__shfl_down_sync(0xffffffffu, a.x, 1), __shfl_down_sync(0xffffffffu, a.y, 1));
    }
    static __device__ __forceinline__ T sel(bool c0, bool c1, T a, T b) { return make_float2(c0 ? a.x : b.x, c1 ? a.y : b.y); }
    static __device__ __forceinline__ float ch(T a, int c) { return c ? a.y : a.x; }
    static __device__ __forceinline__ T pack(float a, float b) { return make_float2(a, b); }
    static __device__ __forceinline__ T flip(T a, uint32_t signmask) {
        return make_float2(__uint_as_float(__float_as_uint(a.x) ^ signmask), __uint_as_float(__float_as_uint(a.y) ^ signmask));
    }
};

// L3_ldexp_q2 (minimp3.d:646-657): y * 2^(-exp_q2/4) by repeated multiplication, same rounding steps.
__device__ __forceinline__ float ldexp_q2(float y, int exp_q2) {
    int e;
    do {
        e = exp_q2 < 120 ? exp_q2 : 120;
        y *= c_expfrac[e & 3] * (float)((1 << 30) >> (e >> 2));
    } while ((exp_q2 -= e) > 0);
    return y;
}

// one step of L3_ldexp_q2 for 0 <= e <= 120: g_expfrac[e & 3] * (float)((1 << 30) >> (e >> 2)); the second factor is a power of two
__device__ __forceinline__ float ldexp_step(int e) { return __fmul_rn(c_expfrac[e & 3], __int_as_float((127 + 30 - (e >> 2)) << 23)); }

// L3_pow_43 for x >= 129 (minimp3.d:737-745)
__device__ __noinline__ float pow43_big(const float* pow43, int x) {
    int mult = 256;
    if (x < 1024) { mult = 16; x <<= 3; }
    int sign = 2 * x & 64;
    float frac = __fdiv_rn((float)((x & 63) - sign), (float)((x & ~63) + sign));
    return __ldg(pow43 + ((x + sign) >> 6)) * (1.0f + frac * ((4.0f / 3) + frac * (2.0f / 9))) * (float)mult;
}

// General path (a value outside the 9-bit range of the shared table somewhere in the lane's trip): the values inside the
// range from the shared table all the same, the others through L3_pow_43's interpolation (positive table of 129 entries in
// global memory, minimp3.d:722-725, 737-745), then the sign.  (-p)*s == -(p*s) exactly.
__device__ __forceinline__ float requant(const float* pow43s, const float* pow43g, int v, float s) {
    if ((unsigned)(v + 256) < 512u) return __fmul_rn(pow43s[v & 511], s);
    const float r = __fmul_rn(pow43_big(pow43g, v < 0 ? -v : v), s);
    return v < 0 ? -r : r;
}

// L3_dct3_9 (minimp3.d:1022-1060), in registers
template <class V>
__device__ __forceinline__ void dct3_9(typename V::T* y) {
    typedef typename V::T T;
    T s0, s1, s2, s3, s4, s5, s6, s7, s8, t0, t2, t4;
    s0 = y[0]; s2 = y[2]; s4 = y[4]; s6 = y[6]; s8 = y[8];
    t0 = V::mac(s0, s6, 0.5f);
    s0 = V::sub(s0, s6);
    t4 = V::muls(V::add(s4, s2), 0.93969262f);
    t2 = V::muls(V::add(s8, s2), 0.76604444f);
    s6 = V::muls(V::sub(s4, s8), 0.17364818f);
    s4 = V::add(s4, V::sub(s8, s2));

    s2 = V::msc(s0, s4, 0.5f);
    y[4] = V::add(s4, s0);
    s8 = V::add(V::sub(t0, t2), s6);
    s0 = V::add(V::sub(t0, t4), t2);
    s4 = V::sub(V::add(t0, t4), s6);

    s1 = y[1]; s3 = y[3]; s5 = y[5]; s7 = y[7];

    s3 = V::muls(s3, 0.86602540f);
    t0 = V::muls(V::add(s5, s1), 0.98480775f);
    t4 = V::muls(V::sub(s5, s7), 0.34202014f);
    t2 = V::muls(V::add(s1, s7), 0.64278761f);
    s1 = V::muls(V::sub(V::sub(s1, s5), s7), 0.86602540f);

    s5 = V::sub(V::sub(t0, s3), t2);
    s7 = V::sub(V::sub(t4, s3), t0);
    s3 = V::sub(V::add(t4, s3), t2);

    y[0] = V::sub(s4, s7);
    y[1] = V::add(s2, s1);
    y[2] = V::sub(s0, s3);
    y[3] = V::add(s8, s5);
    y[5] = V::sub(s8, s5);
    y[6] = V::add(s0, s3);
    y[7] = V::sub(s2, s1);
    y[8] = V::add(s4, s7);
}

// L3_imdct36 for one band (minimp3.d:1062-1100).  wrow = 18 * window row (0 normal, 1 stop), the same for every channel in T
// (a granule whose channels need different rows goes through imdct_split): the weights are scalars, broadcast by the packed
// multiply itself.
template <class V>
__device__ __forceinline__ void imdct36_band(const typename V::T* x, typename V::T* ovl, int wrow, typename V::T* out) {
    typedef typename V::T T;
    T co[9], si[9];
    co[0] = V::neg(x[0]);
    si[0] = x[17];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        si[8 - 2 * i] = V::sub(x[4 * i + 1], x[4 * i + 2]);
        co[1 + 2 * i] = V::add(x[4 * i + 1], x[4 * i + 2]);
        si[7 - 2 * i] = V::sub(x[4 * i + 4], x[4 * i + 3]);
        co[2 + 2 * i] = V::neg(V::add(x[4 * i + 3], x[4 * i + 4]));
    }
    dct3_9<V>(co);
    dct3_9<V>(si);
    si[1] = V::neg(si[1]);
    si[3] = V::neg(si[3]);
    si[5] = V::neg(si[5]);
    si[7] = V::neg(si[7]);
    const float* w = c_mdctw + wrow;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        T o = ovl[i];
        T sum = V::mm_add(co[i], c_twid9[9 + i], si[i], c_twid9[0 + i]);
        ovl[i] = V::mm_sub(co[i], c_twid9[0 + i], si[i], c_twid9[9 + i]);
        out[i] = V::mm_sub(o, w[0 + i], sum, w[9 + i]);
        out[17 - i] = V::mm_add(o, w[9 + i], sum, w[0 + i]);
    }
}

// L3_idct3 / L3_imdct12 (minimp3.d:1102-1129); X(k) = x[OFF + 3k]
template <class V, int OFF>
__device__ __forceinline__ void imdct12(const typename V::T* x, typename V::T* dst, typename V::T* overlap) {
    typedef typename V::T T;
    T co[3], si[3];
    {
        T x0 = V::neg(x[OFF + 0]), x1 = V::add(x[OFF + 6], x[OFF + 3]), x2 = V::add(x[OFF + 12], x[OFF + 9]);
        T a1 = V::msc(x0, x2, 0.5f);
        co[1] = V::add(x0, x2); co[0] = V::mac(a1, x1, 0.86602540f); co[2] = V::msc(a1, x1, 0.86602540f);
    }
    {
        T x0 = x[OFF + 15], x1 = V::sub(x[OFF + 12], x[OFF + 9]), x2 = V::sub(x[OFF + 6], x[OFF + 3]);
        T a1 = V::msc(x0, x2, 0.5f);
        si[1] = V::add(x0, x2); si[0] = V::mac(a1, x1, 0.86602540f); si[2] = V::msc(a1, x1, 0.86602540f);
    }
    si[1] = V::neg(si[1]);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        T o = overlap[i];
        T sum = V::mm_add(co[i], c_twid3[3 + i], si[i], c_twid3[0 + i]);
        overlap[i] = V::mm_sub(co[i], c_twid3[0 + i], si[i], c_twid3[3 + i]);
        dst[i] = V::mm_sub(o, c_twid3[2 - i], sum, c_twid3[5 - i]);
        dst[5 - i] = V::mm_add(o, c_twid3[5 - i], sum, c_twid3[2 - i]);
    }
}

// L3_imdct_short for one band (minimp3.d:1131-1142)
template <class V>
__device__ __forceinline__ void imdct_short_band(const typename V::T* x, typename V::T* ovl, typename V::T* out) {
    typedef typename V::T T;
#pragma unroll
    for (int i = 0; i < 6; i++) out[i] = ovl[i];
    imdct12<V, 0>(x, out + 6, ovl + 6);
    imdct12<V, 1>(x, out + 12, ovl + 6);
    T nd[6];
    imdct12<V, 2>(x, nd, ovl + 6);
#pragma unroll
    for (int i = 0; i < 6; i++) ovl[i] = nd[i];
}

constexpr int kDStride = 33;  // T elements per row (odd: the DCT's per-slot column writes hit distinct banks)

// ---- mbarrier + TMA (1-D bulk copy) helpers ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 32-bit load through a shared-window address (register + constant; a generic pointer to a static table would have its
// cluster-window base re-derived and added in front of every use)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// per-warp shared memory
template <int NCH>
struct __align__(16) WarpSmem {
    // DCT-32 outputs, 33 rows x 33: rows 0..14 are the history (qmf_state), rows 15..32 belong to the current
    // granule.  The current-granule rows ALIAS the spectrum buffer `xr` (natural layout while requantising,
    // x19 padded layout between IMDCT and DCT): each stage has consumed its input before the next one writes.
    typename VT<NCH, false>::T Dbuf[1 + 15 * kDStride + kXrStride];  // D = Dbuf + 1, so that row 15 (= xr) is 16-byte aligned
    uint4 st_is[NCH * kIsChunks];              // TMA-staged inputs of the next granule: quantised spectra (only the chunks
                                               //   that hold anything: `nz_chunks` of each channel),
    uint4 st_rec[NCH * kSfRecBytes / 16];      //   scalefactor records,
    float gains[40][NCH];                      // band gains of this granule (minimp3.d:714-719), the channels of a band adjacent
    uint8_t sfbpair[3][288];
    uint8_t ist[40];
    uint8_t smode[40];
    float kl[40], kr[40];
    uint64_t mbar;
    // delivery window of the tile (see the kernel): read back once per granule by the window stage, parked here rather than
    // in registers (the allocator would spill them to local memory, which misses L1 with the shared-memory carve-out at its
    // maximum)
    char* tbase;
    int rel_lo0, rel_hi0;
    int kstart, kend;   // first halo granule / end of the tile, relative to the tile's first granule (same reason)
};

// L3_intensity_stereo (minimp3.d:898-982) on channel 0's band layout, for the whole warp; X = the granule's spectrum (natural
// order, both channels).  Out of line on purpose: see the call.
__device__ __noinline__ void intensity_stereo(WarpSmem<2>& W, float2* X, int kind0, bool mpeg1, int hb, int mpeg2_sh, const uint8_t* sfbw,
                                              const uint16_t* sfbo, int lane) {
    const int n_long_sfb0 = kind0 == 0 ? 22 : (kind0 == 1 ? 0 : (mpeg1 ? 8 : 6));
    const int n_sfb0 = n_long_sfb0 + (kind0 == 0 ? 0 : (kind0 == 1 ? 39 : 30));
    int mb0 = -1, mb1 = -1, mb2 = -1;
    for (int i = lane; i < n_sfb0; i += 32) {  // L3_stereo_top_band
        const int off = __ldg(sfbo + i), wdt = __ldg(sfbw + i);
        bool nz = false;
#pragma unroll 1
        for (int k = 0; k < wdt; k++) nz |= (X[off + k].y != 0.0f);
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
            W.ist[itop] = (uint8_t)(mb[i] >= prev ? default_pos : W.ist[prev]);
        }
    }
    __syncwarp();
    const unsigned max_pos = mpeg1 ? 7u : 64u;
    for (int i = lane; i < n_sfb0; i += 32) {  // L3_stereo_process: per-sfb decision and gains
        const unsigned ipos = W.ist[i];
        const int mbc = (i % 3) == 0 ? mb0 : ((i % 3) == 1 ? mb1 : mb2);
        uint8_t md = 0;
        if (i > mbc && ipos < max_pos) {
            float kl, kr, s = (hb & 2) ? 1.41421356f : 1.0f;
            if (mpeg1) {
                kl = c_pan[2 * ipos];
                kr = c_pan[2 * ipos + 1];
            } else {
                kl = 1.0f;
                kr = ldexp_q2(1.0f, (int)((ipos + 1) >> 1 << mpeg2_sh));
                if (ipos & 1) { kl = kr; kr = 1.0f; }
            }
            W.kl[i] = kl * s;
            W.kr[i] = kr * s;
            md = 1;
        } else if (hb & 2) {
            md = 2;
        }
        W.smode[i] = md;
    }
    __syncwarp();
#pragma unroll 2
    for (int m = 0; m < 18; m++) {
        const int k = lane + 32 * m;
        const int sfb = W.sfbpair[kind0][k >> 1];
        const int md = W.smode[sfb];
        const float2 v = X[k];
        if (md == 1) X[k] = make_float2(__fmul_rn(v.x, W.kl[sfb]), __fmul_rn(v.x, W.kr[sfb]));
        else if (md == 2) X[k] = make_float2(__fadd_rn(v.x, v.y), __fsub_rn(v.x, v.y));
    }
    __syncwarp();
}

// Requantisation through the general path (minimp3.d:813-816, 737-745, 874-878 + MS stereo :885-896) for the trips of this
// lane flagged in `bigmask`: a value outside the shared table's range.  tap_xr: the granule's row of the xr snapshot or null.
template <int NCH, bool TAPS>
__device__ __noinline__ void requant_wide(WarpSmem<NCH>& W, typename VT<NCH, false>::T* xr, const float* pow43s, const float* pow43g, uint32_t bigmask, int nch0, int nch1,
                                          int kind0, int kind1, bool ms_now, int lane, float* tap_xr) {
    const uint32_t* isw0 = reinterpret_cast<const uint32_t*>(W.st_is);
    const uint32_t* isw1 = isw0 + (NCH - 1) * (kIsChunks * 4);
#pragma unroll 1
    for (; bigmask; bigmask &= bigmask - 1u) {
        const int pi = lane + 32 * (__ffs((int)bigmask) - 1);
        const uint32_t va = (pi >> 2) < nch0 ? isw0[pi] : 0u;
        const float sa = W.gains[W.sfbpair[kind0][pi]][0];
        float a0 = requant(pow43s, pow43g, (int)(int16_t)(va & 0xFFFFu), sa);
        float a1 = requant(pow43s, pow43g, (int)(int16_t)(va >> 16), sa);
        if (NCH == 2) {
            const uint32_t vb = (pi >> 2) < nch1 ? isw1[pi] : 0u;
            const float sb = W.gains[W.sfbpair[kind1][pi]][NCH - 1];
            float b0 = requant(pow43s, pow43g, (int)(int16_t)(vb & 0xFFFFu), sb);
            float b1 = requant(pow43s, pow43g, (int)(int16_t)(vb >> 16), sb);
            if (TAPS && tap_xr) {
                tap_xr[2 * pi] = a0; tap_xr[2 * pi + 1] = a1;
                tap_xr[576 + 2 * pi] = b0; tap_xr[576 + 2 * pi + 1] = b1;
            }
            if (ms_now) {
                const float l0 = __fadd_rn(a0, b0), r0 = __fsub_rn(a0, b0);
                const float l1 = __fadd_rn(a1, b1), r1 = __fsub_rn(a1, b1);
                a0 = l0; b0 = r0; a1 = l1; b1 = r1;
            }
            *reinterpret_cast<float4*>(&xr[2 * pi]) = make_float4(a0, b0, a1, b1);
        } else {
            if (TAPS && tap_xr) { tap_xr[2 * pi] = a0; tap_xr[2 * pi + 1] = a1; }
            *reinterpret_cast<float2*>(&xr[2 * pi]) = make_float2(a0, a1);
        }
    }
}

// The (rare) band where the two channels of a granule use different transforms: one channel at a time.
template <bool FUSED>
__device__ __noinline__ void imdct_split(float2* x, float2* ovl, float2* y, bool sh0, bool sh1, int ws0, int ws1) {
    typedef VT<1, FUSED> V1;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        float xs[18], os[9], ys[18];
#pragma unroll
        for (int i = 0; i < 18; i++) xs[i] = c ? x[i].y : x[i].x;
#pragma unroll
        for (int i = 0; i < 9; i++) os[i] = c ? ovl[i].y : ovl[i].x;
        if (c ? sh1 : sh0) imdct_short_band<V1>(xs, os, ys);
        else imdct36_band<V1>(xs, os, 18 * (c ? ws1 : ws0), ys);
#pragma unroll
        for (int i = 0; i < 18; i++) { if (c) y[i].y = ys[i]; else y[i].x = ys[i]; }
#pragma unroll
        for (int i = 0; i < 9; i++) { if (c) ovl[i].y = os[i]; else ovl[i].x = os[i]; }
    }
}
// mono: the same out-of-line route for the one rare case it has (a per-lane window row, see below)
template <bool FUSED>
__device__ __noinline__ void imdct_split(float* x, float* ovl, float* y, bool sh0, bool, int ws0, int) {
    typedef VT<1, FUSED> V1;
    if (sh0) imdct_short_band<V1>(x, ovl, y);
    else imdct36_band<V1>(x, ovl, 18 * ws0, y);
}

constexpr int kWinStride = 20;                // floats per lane in s_win: 16 weights + 4 (rows of 80 bytes: conflict-free 16-byte loads)
constexpr int kCtaTableBytes = 2048 + 32 * kWinStride * 4;   // s_pow43 (512 floats) | s_win (32 lanes x 20 floats)

// L12: the Layer I / II instance -- a granule is 12 slots x 32 subbands whose samples arrive dequantised and scaled from
// l12_parse_kernel (p.l12_x); only the synthesis half of the pipeline runs (minimp3.d:1567 calls mp3d_synth_granule with 12).
template <int NCH, int WARPS, bool FUSED, bool TAPS, bool S16, bool L12>
__global__ void __launch_bounds__(32 * WARPS, 16 / WARPS) l3_granule_kernel(BatchParams p, const Tile* tiles, uint32_t n_tiles) {
    typedef VT<NCH, FUSED> V;
    typedef typename V::T T;
    constexpr int NS = L12 ? 12 : 18;   // time slots per granule
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // tables shared by the CTA, statically allocated (their shared-memory addresses are compile-time constants, so a lookup
    // is LDS [offset register + constant])
    __shared__ __align__(16) float s_pow43[512];   // 512 signed entries
    __shared__ __align__(16) float s_win[32 * kWinStride];   // synthesis window weights per lane: [lane][tap k][w0 / w1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem<NCH>& W = *reinterpret_cast<WarpSmem<NCH>*>(smem_raw + (size_t)warp * sizeof(WarpSmem<NCH>));
    T* const D = W.Dbuf + 1;
    T* const xr = D + 15 * kDStride;

    // s_pow43[v & 511] = sign(v) * |v|^(4/3) for -256 <= v <= 255: the reference's mirrored table (minimp3.d:722-725, 816)
    // extended to a 9-bit range (table values up to 128, L3_pow_43's own interpolation above) and laid out so that the low
    // nine bits of v are the index
    for (int i = threadIdx.x; i < 512; i += 32 * WARPS) {
        const int a = i < 256 ? i : 512 - i;
        const float pw = a <= 128 ? __ldg(p.t.pow43 + a) : pow43_big(p.t.pow43, a);
        s_pow43[i] = i < 256 ? pw : -pw;
    }

    const uint32_t tile_idx = blockIdx.x * WARPS + warp;
    const bool have_tile = tile_idx < n_tiles;
    const Tile tile = have_tile ? tiles[tile_idx] : Tile{0, 0, 0};
    const l3b_stream_desc_t S = p.streams[tile.stream];
    const bool mpeg1 = S.mpeg1 != 0;
    const int row = S.sr_idx;

    for (int i = lane; i < 3 * 288 / 4; i += 32)
        reinterpret_cast<uint32_t*>(&W.sfbpair[0][0])[i] = reinterpret_cast<const uint32_t*>(p.t.sfb_of_pair + row * 3 * 288)[i];
    for (int i = lane; i < 15 * kDStride; i += 32) D[i] = V::zero();
    for (int i = lane; i < 40; i += 32) W.ist[i] = 0;
    if (lane == 0) mbar_init(&W.mbar, 1);

    // synthesis window weights of a lane: inner index ii = lane & 15 (0..14 active), slot parity par.  Kept in shared memory
    // and fetched at the start of every window stage: held in registers across the IMDCT they cost 16 registers there.
    const int ii = lane & 15, par = lane >> 4;
    for (int i = threadIdx.x; i < 16 * 32; i += 32 * WARPS) {
        const int l = i >> 4, kc = i & 15;
        s_win[l * kWinStride + kc] = (l & 15) < 15 ? __ldg(p.t.win + kc * 15 + (l & 15)) : 0.0f;
    }
#ifdef L3B_EXP_W_REGS   // A/B only: weights held in registers for the whole kernel
    float w0[8], w1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        w0[k] = ii < 15 ? __ldg(p.t.win + (k * 2 + 0) * 15 + ii) : 0.0f;
        w1[k] = ii < 15 ? __ldg(p.t.win + (k * 2 + 1) * 15 + ii) : 0.0f;
    }
#endif

    // recompute halo: up to two granules before the tile, unless decoder state was zeroed in between
    int start = (int)tile.g0;
    for (int depth = 0; depth < 2 && start > 0 && have_tile; depth++) {
        const l3b_grch_desc_t* dp = p.grch + S.first_grch + (uint64_t)start * NCH;
        if (__ldg(&dp->w3) >> 31) break;
        start--;
    }
    // The intensity positions of a frame are per-FRAME scratch in the reference (minimp3.d:1497): what granule 1 does not
    // transmit keeps what granule 0 left there.  A halo that starts on granule 1 of an intensity-stereo frame therefore
    // starts one granule earlier, so that granule 0's intensity pass has run (it only matters when the two channels use
    // different block types, but the extra granule is cheap: it happens once per tile of a stream with an odd delay).
    if (!L12 && NCH == 2 && have_tile && start > 0 && start < (int)tile.g0) {
        const Desc ds = load_desc(p.grch + S.first_grch + (uint64_t)start * NCH);
        if (ds.second_granule() && !ds.reset_before() && (ds.hdr_bits() & 1)) start--;
    }
    // The granules of the tile are k = 0 .. ng-1; the halo is k = kstart .. -1 (two granules, three in the case above).
    const int kstart = start - (int)tile.g0, kend = have_tile ? (int)tile.ng : kstart;
    T ovl[9];
#pragma unroll
    for (int i = 0; i < 9; i++) ovl[i] = V::zero();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int n_long_bands_mixed = 2 << (row == 1 ? 1 : 0);  // minimp3.d:1218
    // Delivery: float samples, or (S16) 16-bit ones at the same element offsets.  Frames [skip, skip + count) of the decoded
    // signal are delivered; relative to the tile's first granule that is [r0, r0 + count), kept as two 32-bit numbers clamped
    // far outside the tile (a tile spans 64 * 576 frames), and tbase points at frame 0 of granule g0 in the delivered signal
    // (only ever dereferenced for delivered frames).
    constexpr long long kGranFrames = 32 * NS;   // 576 (Layer III) / 384 (Layer I / II)
    constexpr int kFrameBytes = NCH * (S16 ? 2 : 4);
    const long long r0 = (long long)(S.pcm_skip / NCH) - (long long)tile.g0 * kGranFrames;
    if (lane == 0) {
        W.kstart = kstart;
        W.kend = kend;
        W.rel_lo0 = (int)max(-(1ll << 30), min(1ll << 30, r0));
        W.rel_hi0 = (int)max(-(1ll << 30), min(1ll << 30, r0 + (long long)(S.pcm_count / NCH)));
        W.tbase = (S16 ? reinterpret_cast<char*>(p.pcm16 + S.pcm_off) : reinterpret_cast<char*>(p.pcm + S.pcm_off)) - r0 * kFrameBytes;
    }
    // descriptor index of granule k, channel 0 (a batch holds fewer than 2^32 granule-channels: l3b_batch_create checks)
    const uint32_t di0 = (uint32_t)S.first_grch + (uint32_t)tile.g0 * NCH;

    // lane 0: TMA bulk copies of granule g's inputs into the staging buffers.  Only the chunks of the spectra that
    // hold anything are fetched: n0 / n1 = nz_chunks of the two channels (from p.nzc, read a granule ahead).
    auto prefetch = [&](uint32_t di32, uint32_t n0, uint32_t n1) {
        const uint64_t di = di32;
        mbar_expect_tx(&W.mbar, (n0 + n1) * 16u + NCH * kSfRecBytes);
        if (n0) tma_load_1d(W.st_is, p.is + di * kIsChunks, n0 * 16u, &W.mbar);
        if (NCH == 2 && n1) tma_load_1d(W.st_is + kIsChunks, p.is + (di + 1) * kIsChunks, n1 * 16u, &W.mbar);
        tma_load_1d(W.st_rec, p.sf + di * kSfRecBytes, NCH * kSfRecBytes, &W.mbar);
    };
    auto load_nz = [&](uint32_t di32, uint32_t& n0, uint32_t& n1) {
        const uint8_t* q = p.nzc + di32;
        n0 = __ldg(q);
        n1 = NCH == 2 ? __ldg(q + 1) : 0u;
    };
    uint32_t nzn0 = 0, nzn1 = 0;   // nz_chunks of the NEXT granule
    if (!L12 && lane == 0 && kstart < kend) {
        // The staging barrier's phase parity is k & 1: a tile whose first iteration is odd completes one empty phase first
        // (the barrier expects one arrival per phase; nothing is in flight yet).
        if (kstart & 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&W.mbar)) : "memory");
        load_nz(di0 + kstart * NCH, nzn0, nzn1);
        prefetch(di0 + kstart * NCH, nzn0, nzn1);
    }
    __syncwarp();
#pragma unroll 1
    for (int k = kstart; k < W.kend; k++) {
        const uint32_t phase = k & 1;
        // 2: a granule of the tile; 1: halo granule whose DCT outputs are history for the window; 0: halo granule that only
        // contributes IMDCT overlap.  Layer I / II: 12 slots per granule, so the 15 history slots span both halo granules.
        const int mode = k >= 0 ? 2 : ((L12 || k == -1) ? 1 : 0);
        const uint32_t di32 = di0 + (uint32_t)(k * NCH);
        const uint64_t di = di32;
        GranFlags d0, d1;   // the descriptor bits of the two channels (from the records)
        d0.v = d1.v = 0;
        int kind0 = 0, kind1 = 0, hb = 0;
        bool ms_frame = false, istereo = false;
        if (L12) {
            {
                const Desc dd = load_desc(p.grch + di);
                if (dd.reset_before() && k != W.kstart)
                    for (int i = lane; i < 15 * kDStride; i += 32) D[i] = V::zero();
                // lane = subband: its 12 samples of both channels, into the padded layout the DCT reads
                const float4* x0 = reinterpret_cast<const float4*>(p.l12_x + di * 384 + lane * 12);
                const float4* x1 = reinterpret_cast<const float4*>(p.l12_x + (di + NCH - 1) * 384 + lane * 12);
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const float4 a = __ldg(x0 + q), b = NCH == 2 ? __ldg(x1 + q) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    xr[lane * 19 + 4 * q + 0] = V::pack(a.x, b.x);
                    xr[lane * 19 + 4 * q + 1] = V::pack(a.y, b.y);
                    xr[lane * 19 + 4 * q + 2] = V::pack(a.z, b.z);
                    xr[lane * 19 + 4 * q + 3] = V::pack(a.w, b.w);
                }
            }
        } else {
            if (lane == 0 && k + 1 < W.kend) load_nz(di32 + NCH, nzn0, nzn1);   // used when the next granule is fetched
            mbar_wait(&W.mbar, phase);
            d0.v = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(W.st_rec) + kSfFlagsOff);
            d1.v = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(W.st_rec) + (NCH - 1) * kSfRecBytes + kSfFlagsOff);
            if (d0.reset_before() && k != W.kstart) {
#pragma unroll
                for (int i = 0; i < 9; i++) ovl[i] = V::zero();
                for (int i = lane; i < 15 * kDStride; i += 32) D[i] = V::zero();
            }
            kind0 = d0.kind(); kind1 = d1.kind();
            hb = d0.hdr_bits();
            ms_frame = (hb & 0xE) == 0x6;       // HDR_IS_MS_STEREO
            istereo = NCH == 2 && (hb & 1);     // HDR_TEST_I_STEREO
            const uint8_t* rec0 = reinterpret_cast<const uint8_t*>(W.st_rec);
            const uint8_t* rec1 = rec0 + (NCH - 1) * kSfRecBytes;

            // band gains (minimp3.d:714-719): scf[i] = gain * 2^(-(iscf[i] << shift)/4); `gain` = 2^(gain_exp/4) comes with
            // the record (l3_scf_kernel), the per-band factor is one table step of L3_ldexp_q2 (two or more when the
            // exponent exceeds 120, which takes a scalefactor above 30).
            {
                const int nsf0 = kind0 == 0 ? 22 : (kind0 == 1 ? 39 : (mpeg1 ? 38 : 36));
                const int nsf1 = kind1 == 0 ? 22 : (kind1 == 1 ? 39 : (mpeg1 ? 38 : 36));
                const int sh0 = d0.scalefac_scale() + 1, sh1 = d1.scalefac_scale() + 1;
                const float g0 = *reinterpret_cast<const float*>(rec0 + kSfGainOff), g1 = *reinterpret_cast<const float*>(rec1 + kSfGainOff);
                // lane = band, both channels; a second round only for the 33rd..40th band of short / mixed blocks
                for (int b = lane; b < max(nsf0, NCH == 2 ? nsf1 : 0); b += 32) {
                    int e0 = (int)rec0[b] << sh0, e1 = NCH == 2 ? (int)rec1[b] << sh1 : 0;
                    float y0 = g0, y1 = g1;
                    if (max(e0, e1) > 120) {   // rare: the loop of L3_ldexp_q2 takes more than one step
                        for (; e0 > 120; e0 -= 120) y0 = __fmul_rn(y0, ldexp_step(120));
                        for (; e1 > 120; e1 -= 120) y1 = __fmul_rn(y1, ldexp_step(120));
                    }
                    W.gains[b][0] = b < nsf0 ? __fmul_rn(y0, ldexp_step(e0)) : 0.0f;
                    if (NCH == 2) W.gains[b][NCH - 1] = b < nsf1 ? __fmul_rn(y1, ldexp_step(e1)) : 0.0f;
                }
            }
#define L3B_SCF0(b) W.gains[b][0]
#define L3B_SCF1(b) W.gains[b][NCH - 1]
            if (istereo) {
                // ist_pos is per-FRAME scratch in the reference (zeroed at frame start, minimp3.d:1497): what granule 1 of
                // channel 1 does not transmit keeps granule 0's values, including the top-band entries that granule 0's
                // L3_intensity_stereo wrote (minimp3.d:974-980).  W.ist still holds exactly that state.
                const int n_sent = !d1.second_granule() ? 0 : (kind1 == 0 ? 21 : (kind1 == 1 ? 36 : 35));   // MPEG-1 partition totals
                for (int i = lane; i < 40; i += 32)
                    if (!d1.second_granule() || i < n_sent) W.ist[i] = rec1[40 + i];
            }
            __syncwarp();

            // ---------------- requantisation (minimp3.d:813-816, 846, 874-878) + MS stereo (:885-896) -------
            {
                const int nch0 = *reinterpret_cast<const uint16_t*>(rec0 + 80);
                const int nch1 = *reinterpret_cast<const uint16_t*>(rec1 + 80);
                const uint32_t* isw0 = reinterpret_cast<const uint32_t*>(W.st_is);
                const uint32_t* isw1 = isw0 + (NCH - 1) * (kIsChunks * 4);
                const bool ms_now = NCH == 2 && ms_frame && !istereo;
                const int nz_hi = max(nch0, NCH == 2 ? nch1 : 0);   // chunks (8 coefficients) holding anything non-zero
                // Fast path without a branch per value: every lane looks its four values up in the mirrored table through
                // an offset that is masked into the table (almost all values lie in [-256, 255]), so the lookups of a trip
                // are independent of each other.  A lane that meets a larger value redoes the trips concerned afterwards
                // through the general path.  The loop stays rolled (instruction-cache footprint) and runs on pointers.
                // byte offset of s_pow43[v & 511] is (4v) & 0x7FC; a value outside [-256, 255] lands somewhere inside the
                // table (and is redone below)
                const char* const tab = reinterpret_cast<const char*>(s_pow43);
                uint32_t bigmask = 0;
                // trips m >= mhi hold nothing in either channel (+0.0, like the memset grbuf): their own short loop, so that
                // the loop over the trips that hold something has no test in it
                const int mhi = min(9, (nz_hi + 7) >> 3);
#pragma unroll 1
                for (int m = mhi; m < 9; m++) {
                    const int pi = lane + 32 * m;
                    if (NCH == 2) *reinterpret_cast<float4*>(&xr[2 * pi]) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    else *reinterpret_cast<float2*>(&xr[2 * pi]) = make_float2(0.0f, 0.0f);
                    if (TAPS && mode == 2)
                        for (int c = 0; c < NCH; c++) p.tap_xr[(di + c) * 576 + 2 * pi] = p.tap_xr[(di + c) * 576 + 2 * pi + 1] = 0.0f;
                }
#pragma unroll kRqUnroll
                for (int m = 0; m < mhi; m++) {
                    const int pi = lane + 32 * m;
                    const uint32_t va = (pi >> 2) < nch0 ? isw0[pi] : 0u;   // chunks past nz_chunks were never fetched
                    const uint32_t vb = (NCH == 2 && (pi >> 2) < nch1) ? isw1[pi] : 0u;
                    const float sa = L3B_SCF0(W.sfbpair[kind0][pi]);
                    const float sb = NCH == 2 ? L3B_SCF1(W.sfbpair[kind1][pi]) : 0.0f;
                    // a 16-bit value lies in [-256, 255] iff its bits 15..8 are all equal
                    const uint32_t wide = ((va ^ (va << 1)) | (vb ^ (vb << 1))) & 0xFE00FE00u;
                    bigmask |= (wide ? 1u : 0u) << m;
                    float a0 = *reinterpret_cast<const float*>(tab + ((va << 2) & 0x7FCu));
                    float a1 = *reinterpret_cast<const float*>(tab + ((va >> 14) & 0x7FCu));
                    if (NCH == 2) {
                        float b0 = *reinterpret_cast<const float*>(tab + ((vb << 2) & 0x7FCu));
                        float b1 = *reinterpret_cast<const float*>(tab + ((vb >> 14) & 0x7FCu));
                        {   // both channels of a coefficient in one packed multiply (two independent roundings, like two FMULs)
                            const float2 u0 = __fmul2_rn(make_float2(a0, b0), make_float2(sa, sb)), u1 = __fmul2_rn(make_float2(a1, b1), make_float2(sa, sb));
                            a0 = u0.x; b0 = u0.y; a1 = u1.x; b1 = u1.y;
                        }
                        if (TAPS && mode == 2) {
                            p.tap_xr[di * 576 + 2 * pi] = a0; p.tap_xr[di * 576 + 2 * pi + 1] = a1;
                            p.tap_xr[(di + 1) * 576 + 2 * pi] = b0; p.tap_xr[(di + 1) * 576 + 2 * pi + 1] = b1;
                        }
                        if (ms_now) {
                            const float l0 = __fadd_rn(a0, b0), r0 = __fsub_rn(a0, b0);
                            const float l1 = __fadd_rn(a1, b1), r1 = __fsub_rn(a1, b1);
                            a0 = l0; b0 = r0; a1 = l1; b1 = r1;
                        }
                        *reinterpret_cast<float4*>(&xr[2 * pi]) = make_float4(a0, b0, a1, b1);
                    } else {
                        a0 = __fmul_rn(a0, sa); a1 = __fmul_rn(a1, sa);
                        if (TAPS && mode == 2) { p.tap_xr[di * 576 + 2 * pi] = a0; p.tap_xr[di * 576 + 2 * pi + 1] = a1; }
                        *reinterpret_cast<float2*>(&xr[2 * pi]) = make_float2(a0, a1);
                    }
                }
                // rare: a value outside [-256, 255] somewhere in this lane's coefficients -- those trips again, through the general
                // path (out of line, like the intensity stereo pass)
                if (bigmask)
                    requant_wide<NCH, TAPS>(W, xr, s_pow43, p.t.pow43, bigmask, nch0, nch1, kind0, kind1, ms_now, lane, TAPS && mode == 2 ? p.tap_xr + di * 576 : nullptr);
            }
            __syncwarp();
            // the staging buffers are free again: fetch the next granule while this one is transformed
            if (lane == 0 && k + 1 < W.kend) {
                fence_proxy_async();   // generic -> async proxy ordering of the staging buffers
                prefetch(di32 + NCH, nzn0, nzn1);
            }

            // ---------------- intensity stereo (minimp3.d:898-982), on channel 0's band layout ----------------
            // (out of line: the granule loop's instruction stream stays short and contiguous for the streams without it)
            if constexpr (NCH == 2) {
                if (istereo)
                    intensity_stereo(W, reinterpret_cast<float2*>(xr), kind0, mpeg1, hb, d1.scalefac_compress_lsb(),
                                     p.t.sfb_width + (row * 3 + kind0) * 40, p.t.sfb_start + (row * 3 + kind0) * 40, lane);
            }
            // A MONO frame whose header has the intensity bit set (mode_extension is "don't care" outside joint stereo,
            // but HDR_TEST_I_STEREO looks at the bit in every mode, minimp3.d:100, 1207): the reference runs
            // L3_intensity_stereo on (this channel, a zeroed second channel, zeroed ist_pos): every band is "intensity"
            // with position 0, so the spectrum is multiplied by kl*s -- g_pan[0] = 0 in MPEG-1 (the frame goes silent,
            // signed zeros included), 1 in MPEG-2 -- with s = sqrt(2) when the MS bit is set too (minimp3.d:928-961).
            if (NCH == 1 && hb & 1) {
                const float s = (hb & 2) ? 1.41421356f : 1.0f;
                const float k = __fmul_rn(mpeg1 ? c_pan[0] : 1.0f, s);
                float* X = reinterpret_cast<float*>(xr);
#pragma unroll 6
                for (int m = 0; m < 18; m++) X[lane + 32 * m] = __fmul_rn(X[lane + 32 * m], k);
                __syncwarp();
            }
            if (TAPS && mode == 2) {   // spectrum after stereo processing (after minimp3.d:1213)
                for (int m = 0; m < 18; m++) {
                    const int k = lane + 32 * m;
                    for (int c = 0; c < NCH; c++) p.tap_st[(di + c) * 576 + k] = V::ch(xr[k], c);
                }
            }
        }

        // ---------------- reorder + antialias + IMDCT + frequency inversion (minimp3.d:1215-1229) ----------
        if (!L12) {
            T x[18], y[18];
            const int bt0 = d0.block_type(), bt1 = d1.block_type();
            const int nlb0 = kind0 == 2 ? n_long_bands_mixed : 0, nlb1 = kind1 == 2 ? n_long_bands_mixed : 0;
            if (kind0 == 0 && kind1 == 0) {
#pragma unroll
                for (int i = 0; i < 18; i++) x[i] = xr[lane * 18 + i];
                // alias reduction across all 31 band boundaries (minimp3.d:1002-1020): the neighbours' elements are read
                // from the buffer (eight adjacent elements each: 16-byte loads for stereo) instead of being shuffled in one
                // by one.  Lanes 0 / 31 read eight elements outside the spectrum (inside the warp's buffer) and drop them.
                T dnv[8], upv[8];   // band-1: elements 10..17, band+1: elements 0..7
                if (NCH == 2) {
                    const float4* dq = reinterpret_cast<const float4*>(xr + (lane - 1) * 18 + 10);
                    const float4* uq = reinterpret_cast<const float4*>(xr + (lane + 1) * 18);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float4 a = dq[q], b = uq[q];
                        dnv[2 * q] = V::pack(a.x, a.y); dnv[2 * q + 1] = V::pack(a.z, a.w);
                        upv[2 * q] = V::pack(b.x, b.y); upv[2 * q + 1] = V::pack(b.z, b.w);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) { dnv[i] = xr[(lane - 1) * 18 + 10 + i]; upv[i] = xr[(lane + 1) * 18 + i]; }
                }
                const bool lo = lane >= 1, up = lane < 31;
                T nlo[8], nhi[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    nlo[i] = V::mm_sub(x[i], c_aa[i], dnv[7 - i], c_aa[8 + i]);
                    nhi[i] = V::mm_add(upv[i], c_aa[8 + i], x[17 - i], c_aa[i]);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    x[i] = V::sel(lo, lo, nlo[i], x[i]);
                    x[17 - i] = V::sel(up, up, nhi[i], x[17 - i]);
                }
                __syncwarp();  // every lane has its inputs in registers: the buffer may be rewritten in the padded (x19) layout
            } else {
                // short / mixed blocks: L3_reorder folded into the load through the permutation table
                const uint16_t* pm0 = p.t.perm + (row * 2 + (kind0 == 2 ? 1 : 0)) * 576 + lane * 18;
                const uint16_t* pm1 = p.t.perm + (row * 2 + (kind1 == 2 ? 1 : 0)) * 576 + lane * 18;
                const float* xf = reinterpret_cast<const float*>(xr);
#pragma unroll
                for (int i = 0; i < 18; i++) {
                    const int i0 = kind0 == 0 ? lane * 18 + i : (int)__ldg(pm0 + i);
                    if (NCH == 2) {
                        const int i1 = kind1 == 0 ? lane * 18 + i : (int)__ldg(pm1 + i);
                        x[i] = V::pack(xf[2 * i0], xf[2 * i1 + 1]);
                    } else {
                        x[i] = V::pack(xf[i0], 0.0f);
                    }
                }
                __syncwarp();  // every lane has its inputs in registers: the buffer may be rewritten in the padded (x19) layout,
                               // and the stores below are free to interleave with the transform
                const int aa0 = kind0 == 0 ? 31 : nlb0 - 1, aa1 = kind1 == 0 ? 31 : nlb1 - 1;
                if (aa0 > 0 || aa1 > 0) {
                    const bool lo0 = lane >= 1 && lane - 1 < aa0, up0 = lane < aa0;
                    const bool lo1 = lane >= 1 && lane - 1 < aa1, up1 = lane < aa1;
                    T nlo[8], nhi[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const T dn = V::shfl_up(x[17 - i]);   // band-1, element 17-i
                        const T up = V::shfl_down(x[i]);      // band+1, element i
                        nlo[i] = V::mm_sub(x[i], c_aa[i], dn, c_aa[8 + i]);
                        nhi[i] = V::mm_add(up, c_aa[8 + i], x[17 - i], c_aa[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        x[i] = V::sel(lo0, lo1, nlo[i], x[i]);
                        x[17 - i] = V::sel(up0, up1, nhi[i], x[17 - i]);
                    }
                }
            }
            const bool sh0 = bt0 == 2 && lane >= nlb0, sh1 = bt1 == 2 && lane >= nlb1;
            // window row: the stop window, except in the long bands of a granule whose mixed_block_flag is set --
            // the reference honours the flag on every block type (minimp3.d:1212, 1158-1167), so a STOP block that
            // closes a mixed run keeps the normal window in its lowest bands.  Only then does the row differ from lane to
            // lane; everywhere else it is warp-uniform, and the window weights stay uniform constant-bank operands.
            const bool stop_mixed = (bt0 == 3 && d0.mixed()) || (NCH == 2 && bt1 == 3 && d1.mixed());
            const bool same_row = NCH == 1 || (bt0 == 3) == (bt1 == 3);
            const int wrow = bt0 == 3 ? 18 : 0;   // window row of the long transform: 0 normal / start, 1 stop
            if (!stop_mixed && (NCH == 1 || sh0 == sh1) && (sh0 || same_row)) {
                if (sh0) imdct_short_band<V>(x, ovl, y);
                else imdct36_band<V>(x, ovl, wrow, y);
            } else {
                // Rare: the channels use different transforms or window rows in this band, or the row is per lane.  The
                // out-of-line helper works on COPIES so that x / ovl / y themselves never have their address taken
                // (they must stay in registers).
                const int ws0 = (bt0 == 3 && lane >= (d0.mixed() ? n_long_bands_mixed : 0)) ? 1 : 0;
                const int ws1 = (bt1 == 3 && lane >= (d1.mixed() ? n_long_bands_mixed : 0)) ? 1 : 0;
                T xc[18], oc[9], yc[18];
#pragma unroll
                for (int i = 0; i < 18; i++) xc[i] = x[i];
#pragma unroll
                for (int i = 0; i < 9; i++) oc[i] = ovl[i];
                imdct_split<FUSED>(xc, oc, yc, sh0, sh1, ws0, ws1);
#pragma unroll
                for (int i = 0; i < 18; i++) y[i] = yc[i];
#pragma unroll
                for (int i = 0; i < 9; i++) ovl[i] = oc[i];
            }
            if (mode >= 1) {
                const uint32_t fm = (lane & 1) ? 0x80000000u : 0u;   // L3_change_sign: odd samples of odd bands
#pragma unroll
                for (int i = 1; i < 18; i += 2) y[i] = V::flip(y[i], fm);
#pragma unroll
                for (int i = 0; i < 18; i++) xr[lane * 19 + i] = y[i];
                if (TAPS && mode == 2) {
                    for (int i = 0; i < 18; i++)
                        for (int c = 0; c < NCH; c++) p.tap_im[(di + c) * 576 + lane * 18 + i] = V::ch(y[i], c);
                }
            }
        }

        __syncwarp();   // the transform's rows (lane = subband) are read by columns (lane = time slot) from here on

        // ---------------- DCT-32 matrixing across bands, one time slot per lane (minimp3.d:1232-1298) -------
        if (mode >= 1 && lane < NS) {
            T t[4][8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                T x0 = xr[i * 19 + lane];
                T x1 = xr[(15 - i) * 19 + lane];
                T x2 = xr[(16 + i) * 19 + lane];
                T x3 = xr[(31 - i) * 19 + lane];
                T t0 = V::add(x0, x3);
                T t1 = V::add(x1, x2);
                T t2 = V::muls(V::sub(x1, x2), c_sec[3 * i + 0]);
                T t3 = V::muls(V::sub(x0, x3), c_sec[3 * i + 1]);
                t[0][i] = V::add(t0, t1);
                t[1][i] = V::muls(V::sub(t0, t1), c_sec[3 * i + 2]);
                t[2][i] = V::add(t3, t2);
                t[3][i] = V::muls(V::sub(t3, t2), c_sec[3 * i + 2]);
            }
            __syncwarp((1u << NS) - 1u);  // all slots have read their column: the buffer may now take the output rows
#pragma unroll
            for (int r = 0; r < 4; r++) {
                T x0 = t[r][0], x1 = t[r][1], x2 = t[r][2], x3 = t[r][3], x4 = t[r][4], x5 = t[r][5], x6 = t[r][6], x7 = t[r][7], xt;
                xt = V::sub(x0, x7); x0 = V::add(x0, x7);
                x7 = V::sub(x1, x6); x1 = V::add(x1, x6);
                x6 = V::sub(x2, x5); x2 = V::add(x2, x5);
                x5 = V::sub(x3, x4); x3 = V::add(x3, x4);
                x4 = V::sub(x0, x3); x0 = V::add(x0, x3);
                x3 = V::sub(x1, x2); x1 = V::add(x1, x2);
                t[r][0] = V::add(x0, x1);
                t[r][4] = V::muls(V::sub(x0, x1), 0.70710677f);
                x5 = V::add(x5, x6);
                x6 = V::muls(V::add(x6, x7), 0.70710677f);
                x7 = V::add(x7, xt);
                x3 = V::muls(V::add(x3, x4), 0.70710677f);
                x5 = V::msc(x5, x7, 0.198912367f);
                x7 = V::mac(x7, x5, 0.382683432f);
                x5 = V::msc(x5, x7, 0.198912367f);
                x0 = V::sub(xt, x6); xt = V::add(xt, x6);
                t[r][1] = V::muls(V::add(xt, x7), 0.50979561f);
                t[r][2] = V::muls(V::add(x4, x3), 0.54119611f);
                t[r][3] = V::muls(V::sub(x0, x5), 0.60134488f);
                t[r][5] = V::muls(V::add(x0, x5), 0.89997619f);
                t[r][6] = V::muls(V::sub(x4, x3), 1.30656302f);
                t[r][7] = V::muls(V::sub(xt, x7), 2.56291556f);
            }
            T* out = D + (15 + lane) * kDStride;
#pragma unroll
            for (int i = 0; i < 7; i++) {
                out[4 * i + 0] = t[0][i];
                out[4 * i + 1] = V::add(V::add(t[2][i], t[3][i]), t[3][i + 1]);
                out[4 * i + 2] = V::add(t[1][i], t[1][i + 1]);
                out[4 * i + 3] = V::add(V::add(t[2][i + 1], t[3][i]), t[3][i + 1]);
            }
            out[28] = t[0][7];
            out[29] = V::add(t[2][7], t[3][7]);
            out[30] = t[1][7];
            out[31] = t[3][7];
            if (TAPS && mode == 2) {   // the reference's in-place layout: output j of slot k at grbuf[j*18 + k]
                __syncwarp((1u << NS) - 1u);
                for (int j = 0; j < 32; j++)
                    for (int c = 0; c < NCH; c++) p.tap_dct[(di + c) * 576 + j * 18 + lane] = V::ch(out[j], c);
            }
        }
        __syncwarp();   // (a CTA barrier here measured slower; the warp's own writes must still be ordered before the window's reads)

        // ---------------- 512-tap window (minimp3.d:1305-1406) ----------------
        if (mode == 2) {
            // samples [lo, hi) of this granule are delivered (all of them except at the edges of the stream's PCM range)
            const int dlo = max(0, min((int)kGranFrames, W.rel_lo0 - k * (int)kGranFrames));
            const int dhi = max(0, min((int)kGranFrames, W.rel_hi0 - k * (int)kGranFrames));
            const unsigned span = (unsigned)(dhi - dlo);
#define L3B_DELIVER(f) ((unsigned)((f) - dlo) < span)
            // frame 0 of this granule in the delivered signal (only dereferenced for delivered frames); one byte pointer per
            // granule, compile-time offsets from there: a store is a range test and a predicated STG
            char* const gbase = W.tbase + (long long)k * (kGranFrames * kFrameBytes);
            const float scale = 1.0f / 32768.0f;
            // 16-bit delivery: q = clamp(lrintf(x * 32768), -32768, 32767) of the float sample x the float path would
            // have written (the conversion SURVEY 8c defines; un-dithered, wav.d:475-700)
            auto store = [&](char* q, T v) {
                const T s = V::muls(v, scale);
                if (S16) {
                    if (NCH == 2) {
                        const int q0 = max(-32768, min(32767, __float2int_rn(__fmul_rn(V::ch(s, 0), 32768.0f))));
                        const int q1 = max(-32768, min(32767, __float2int_rn(__fmul_rn(V::ch(s, 1), 32768.0f))));
                        *reinterpret_cast<uint32_t*>(q) = (uint32_t)(q0 & 0xFFFF) | ((uint32_t)q1 << 16);
                    } else {
                        *reinterpret_cast<int16_t*>(q) = (int16_t)max(-32768, min(32767, __float2int_rn(__fmul_rn(V::ch(s, 0), 32768.0f))));
                    }
                } else {
                    *reinterpret_cast<T*>(q) = s;
                }
            };
            if (ii < 15) {
#ifndef L3B_EXP_W_REGS
                float w0[8], w1[8];
#pragma unroll
                for (int k = 0; k < 4; k++) {   // four 16-byte loads
                    const float4 w = *reinterpret_cast<const float4*>(&s_win[lane * kWinStride + 4 * k]);
                    w0[2 * k] = w.x; w1[2 * k] = w.y; w0[2 * k + 1] = w.z; w1[2 * k + 1] = w.w;
                }
#endif
                // lane (par, ii) produces samples 15-ii and 17+ii of slots s = 2q + par.
                // V[j] = D[row par + j][ j odd ? 31-ii : 1+ii ]  -- row r of D is slot r-15 ("window")
                // Sliding window of 16 taps in registers; three slots per loop trip (the window then moves by 6
                // registers), which keeps the unrolled body ~3x smaller than a full 9-slot unroll (instruction cache).
                T Vw[22];
                const T* base_lo = D + par * kDStride + (1 + ii);
                const T* base_hi = D + par * kDStride + (31 - ii);
#pragma unroll
                for (int j = 0; j < 16; j++) Vw[j] = (j & 1) ? base_hi[j * kDStride] : base_lo[j * kDStride];
                int fa = 32 * par + 15 - ii, fb = 32 * par + 17 + ii;   // this lane's two samples of slot `par`
                char* pa = gbase + fa * kFrameBytes;
                char* pb = gbase + fb * kFrameBytes;
#pragma unroll 1
                for (int q3 = 0;; q3++, fa += 192, fb += 192, pa += 192 * kFrameBytes, pb += 192 * kFrameBytes) {
                    const T* lo = base_lo + q3 * 6 * kDStride;
                    const T* hi = base_hi + q3 * 6 * kDStride;
#pragma unroll
                    for (int qq = 0; qq < 3; qq++) {
                        if (q3 + qq > 0) {   // slot 0 of the granule has all its taps from the initial load
                            Vw[2 * qq + 14] = lo[(2 * qq + 14) * kDStride];
                            Vw[2 * qq + 15] = hi[(2 * qq + 15) * kDStride];
                        }
                        T a, b;
                        {
                            const T vz = Vw[2 * qq + 15], vy = Vw[2 * qq + 0];
                            b = V::mm_add(vz, w1[0], vy, w0[0]);
                            a = V::mm_sub(vz, w0[0], vy, w1[0]);
                        }
#pragma unroll
                        for (int k = 1; k < 8; k++) {
                            const T vz = Vw[2 * qq + 15 - k], vy = Vw[2 * qq + k];
                            if (FUSED) {
                                b = V::mac(V::mac(b, vz, w1[k]), vy, w0[k]);
                                if (k & 1) a = V::msc(V::mac(a, vy, w1[k]), vz, w0[k]);
                                else a = V::msc(V::mac(a, vz, w0[k]), vy, w1[k]);
                            } else {
                                b = V::add(b, V::mm_add(vz, w1[k], vy, w0[k]));
                                if (k & 1) a = V::add(a, V::mm_sub(vy, w1[k], vz, w0[k]));
                                else a = V::add(a, V::mm_sub(vz, w0[k], vy, w1[k]));
                            }
                        }
                        // slot s = 2 * (3 * q3 + qq) + par: samples 32 s + 15 - ii and 32 s + 17 + ii
                        if (L3B_DELIVER(fa + 64 * qq)) store(pa + 64 * qq * kFrameBytes, a);
                        if (L3B_DELIVER(fb + 64 * qq)) store(pb + 64 * qq * kFrameBytes, b);
                    }
                    if (q3 == NS / 6 - 1) break;   // (nothing to slide after the last trip)
#pragma unroll
                    for (int j = 0; j < 16; j++) Vw[j] = Vw[j + 6];   // slide by three slots
                }
            }
            // samples 0 and 16 of every slot (mp3d_synth_pair), one slot per lane
            if (lane < NS) {
                const T* col = D + lane * kDStride;   // row lane + k is slot lane - 15 + k
                T z[15];
#pragma unroll
                for (int k = 0; k < 15; k++) z[k] = col[k * kDStride + 16];
                T a;
                a = V::muls(V::sub(z[14], z[0]), 29.0f);
                a = V::mac(a, V::add(z[1], z[13]), 213.0f);
                a = V::mac(a, V::sub(z[12], z[2]), 459.0f);
                a = V::mac(a, V::add(z[3], z[11]), 2037.0f);
                a = V::mac(a, V::sub(z[10], z[4]), 5153.0f);
                a = V::mac(a, V::add(z[5], z[9]), 6574.0f);
                a = V::mac(a, V::sub(z[8], z[6]), 37489.0f);
                a = V::mac(a, z[7], 75038.0f);
                const int fa = 32 * lane, fb = 32 * lane + 16;
                if (L3B_DELIVER(fa)) store(gbase + fa * kFrameBytes, a);
#pragma unroll
                for (int k = 0; k < 15; k += 2) z[k] = col[k * kDStride];
                a = V::muls(z[14], 104.0f);
                a = V::mac(a, z[12], 1567.0f);
                a = V::mac(a, z[10], 9727.0f);
                a = V::mac(a, z[8], 64019.0f);
                a = V::mac(a, z[6], -9975.0f);
                a = V::mac(a, z[4], -45.0f);
                a = V::mac(a, z[2], 146.0f);
                a = V::mac(a, z[0], -5.0f);
                if (L3B_DELIVER(fb)) store(gbase + fb * kFrameBytes, a);
            }
        }
#undef L3B_DELIVER
        // slide the history: the last 15 slots become rows 0..14 (qmf_state, minimp3.d:1423-1433)
        if (mode >= 1) {
            __syncwarp();
            if (NCH == 2) {
                // 15 x 33 float2 = 495 elements moved down by 18 rows.  D is 8 bytes past a 16-byte boundary, and so is
                // D + 18 rows: element 0 goes alone, elements 1..494 as 247 16-byte vectors.
                const float4* src = reinterpret_cast<const float4*>(D + NS * kDStride + 1);
                float4* dst = reinterpret_cast<float4*>(D + 1);
                const T first = D[NS * kDStride];
                float4 tmp[8];
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const int e = lane + 32 * m;
                    if (e < 247) tmp[m] = src[e];
                }
                __syncwarp();
                if (lane == 0) D[0] = first;
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const int e = lane + 32 * m;
                    if (e < 247) dst[e] = tmp[m];
                }
            } else {
                T tmp[16];
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const int e = lane + 32 * m;
                    if (e < 15 * kDStride) tmp[m] = D[NS * kDStride + e];
                }
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const int e = lane + 32 * m;
                    if (e < 15 * kDStride) D[e] = tmp[m];
                }
            }
            __syncwarp();
        }
    }
}

template <int NCH, int WARPS, bool FUSED, bool TAPS, bool S16, bool L12>
static cudaError_t launch_granule_t(const BatchParams& p, const Tile* tiles, uint32_t n, cudaStream_t s) {
    if (!n) return cudaSuccess;
    const size_t smem = (size_t)WARPS * sizeof(WarpSmem<NCH>);   // + kCtaTableBytes of static tables
    // the attribute is per device and a process may hold contexts on several: set it every time (a cheap driver call)
    cudaError_t e = cudaFuncSetAttribute(l3_granule_kernel<NCH, WARPS, FUSED, TAPS, S16, L12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    l3_granule_kernel<NCH, WARPS, FUSED, TAPS, S16, L12><<<(n + WARPS - 1) / WARPS, 32 * WARPS, smem, s>>>(p, tiles, n);
    return cudaGetLastError();
}

template <bool FUSED, bool TAPS, bool S16, bool L12>
static cudaError_t launch_both(const BatchParams& p, const Tile* ts, uint32_t ns, const Tile* tm, uint32_t nm, cudaStream_t s) {
    cudaError_t e = launch_granule_t<2, kGranuleWarpsStereo, FUSED, TAPS, S16, L12>(p, ts, ns, s);
    if (e == cudaSuccess) e = launch_granule_t<1, kGranuleWarpsMono, FUSED, TAPS, S16, L12>(p, tm, nm, s);
    return e;
}

cudaError_t launch_granule(const BatchParams& p, const Tile* tiles_stereo, uint32_t n_stereo, const Tile* tiles_mono,
                           uint32_t n_mono, cudaStream_t s, bool fused, bool taps) {
    const bool s16 = p.pcm16 != nullptr;
    // float taps exist in the bit-exact float-delivery mode only (they are compared bitwise)
    if (taps) return launch_both<false, true, false, false>(p, tiles_stereo, n_stereo, tiles_mono, n_mono, s);
    if (fused) return s16 ? launch_both<true, false, true, false>(p, tiles_stereo, n_stereo, tiles_mono, n_mono, s)
                          : launch_both<true, false, false, false>(p, tiles_stereo, n_stereo, tiles_mono, n_mono, s);
    return s16 ? launch_both<false, false, true, false>(p, tiles_stereo, n_stereo, tiles_mono, n_mono, s)
               : launch_both<false, false, false, false>(p, tiles_stereo, n_stereo, tiles_mono, n_mono, s);
}

cudaError_t launch_granule_l12(const BatchParams& p, const Tile* tiles_stereo, uint32_t n_stereo, const Tile* tiles_mono,
                               uint32_t n_mono, cudaStream_t s, bool fused) {
    const bool s16 = p.pcm16 != nullptr;
    if (fused) return s16 ? launch_both<true, false, true, true>(p, tiles_stereo, n_stereo, tiles_mono, n_mono, s)
                          : launch_both<true, false, false, true>(p, tiles_stereo, n_stereo, tiles_mono, n_mono, s);
    return s16 ? launch_both<false, false, true, true>(p, tiles_stereo, n_stereo, tiles_mono, n_mono, s)
               : launch_both<false, false, false, true>(p, tiles_stereo, n_stereo, tiles_mono, n_mono, s);
}

}  // namespace l3b
