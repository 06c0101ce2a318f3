// Micro-benchmark: the DCT-32 "matrixing" of the polyphase synthesis (mp3d_DCT_II, minimp3.d:1232-1298) done two ways on
// shared-memory-resident data, as it would sit inside the fused granule kernel:
//   (a) the reference's fast algorithm on the FP32 pipe, packed float2 (both channels per register), every product and sum
//       rounded separately (bit-exact), 296 flops per 32-point transform;
//   (b) the same linear map as a dense 32x32 matrix product on the tensor cores with the FP32-accurate 3xTF32 split
//       (x = x_hi + x_lo, c = c_hi + c_lo, D = c_hi x_hi + c_hi x_lo + c_lo x_hi, FP32 accumulate), mma.sync.m16n8k8.tf32.
// One "granule" = 18 time slots x 2 channels = 36 columns of 32 subband samples (padded to 40 for the n=8 tiles).
// Prints granules per microsecond per SM for both and the deviation of (b) from (a).
// nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o dct32_tc dct32_tc.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

static const float SEC[24] = {10.19000816f, 0.50060302f, 0.50241929f, 3.40760851f, 0.50547093f, 0.52249861f, 2.05778098f, 0.51544732f,
                              0.56694406f, 1.48416460f, 0.53104258f, 0.64682180f, 1.16943991f, 0.55310392f, 0.78815460f, 0.97256821f,
                              0.58293498f, 1.06067765f, 0.83934963f, 0.62250412f, 1.72244716f, 0.74453628f, 0.67480832f, 5.10114861f};
__constant__ float c_sec[24];

// the reference's algorithm on one column x[32] -> y[32] (natural output order), T = float / double / float2-like
template <class T, class S>
__host__ __device__ inline void dct32(const T* x, T* y, const S* sec) {
    T t[4][8];
    for (int i = 0; i < 8; i++) {
        T x0 = x[i], x1 = x[15 - i], x2 = x[16 + i], x3 = x[31 - i];
        T t0 = x0 + x3, t1 = x1 + x2, t2 = (x1 - x2) * sec[3 * i + 0], t3 = (x0 - x3) * sec[3 * i + 1];
        t[0][i] = t0 + t1; t[1][i] = (t0 - t1) * sec[3 * i + 2]; t[2][i] = t3 + t2; t[3][i] = (t3 - t2) * sec[3 * i + 2];
    }
    for (int r = 0; r < 4; r++) {
        T x0 = t[r][0], x1 = t[r][1], x2 = t[r][2], x3 = t[r][3], x4 = t[r][4], x5 = t[r][5], x6 = t[r][6], x7 = t[r][7], xt;
        xt = x0 - x7; x0 = x0 + x7; x7 = x1 - x6; x1 = x1 + x6; x6 = x2 - x5; x2 = x2 + x5; x5 = x3 - x4; x3 = x3 + x4;
        x4 = x0 - x3; x0 = x0 + x3; x3 = x1 - x2; x1 = x1 + x2;
        t[r][0] = x0 + x1; t[r][4] = (x0 - x1) * (S)0.70710677f;
        x5 = x5 + x6; x6 = (x6 + x7) * (S)0.70710677f; x7 = x7 + xt; x3 = (x3 + x4) * (S)0.70710677f;
        x5 = x5 - x7 * (S)0.198912367f; x7 = x7 + x5 * (S)0.382683432f; x5 = x5 - x7 * (S)0.198912367f;
        x0 = xt - x6; xt = xt + x6;
        t[r][1] = (xt + x7) * (S)0.50979561f; t[r][2] = (x4 + x3) * (S)0.54119611f; t[r][3] = (x0 - x5) * (S)0.60134488f;
        t[r][5] = (x0 + x5) * (S)0.89997619f; t[r][6] = (x4 - x3) * (S)1.30656302f; t[r][7] = (xt - x7) * (S)2.56291556f;
    }
    for (int i = 0; i < 7; i++) {
        y[4 * i + 0] = t[0][i]; y[4 * i + 1] = t[2][i] + t[3][i] + t[3][i + 1];
        y[4 * i + 2] = t[1][i] + t[1][i + 1]; y[4 * i + 3] = t[2][i + 1] + t[3][i] + t[3][i + 1];
    }
    y[28] = t[0][7]; y[29] = t[2][7] + t[3][7]; y[30] = t[1][7]; y[31] = t[3][7];
}

struct F2 {   // both channels in one register pair; packed FMUL2 / FFMA2-by-one like the product kernel (exactly rounded)
    float2 v;
};
__constant__ float2 c_one2;
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { F2 r; r.v = __ffma2_rn(a.v, c_one2, b.v); return r; }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { F2 r; r.v = __ffma2_rn(a.v, c_one2, make_float2(-b.v.x, -b.v.y)); return r; }
__device__ __forceinline__ F2 operator*(F2 a, float s) { F2 r; r.v = __fmul2_rn(a.v, make_float2(s, s)); return r; }

constexpr int kCols = 40, kStride = 41;   // columns per granule (36 used), padded row stride in floats

// (a) FP32 pipe: lane = time slot (18 lanes), float2 = (ch0, ch1); data [band][slot] float2 in shared memory
__global__ void __launch_bounds__(128) k_fma(float* out, int iters) {
    __shared__ float2 buf[4][32 * 19];
    float2* b = buf[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < 32 * 19; i += 32) b[i] = make_float2(sinf(i * 0.37f + threadIdx.x), cosf(i * 0.11f));
    __syncwarp();
    for (int it = 0; it < iters; it++) {
        if (lane < 18) {
            F2 x[32], y[32];
#pragma unroll
            for (int i = 0; i < 32; i++) x[i].v = b[i * 19 + lane];
            dct32<F2, float>(x, y, c_sec);
#pragma unroll
            for (int i = 0; i < 32; i++) b[i * 19 + lane] = make_float2(y[i].v.x * 0.1f, y[i].v.y * 0.1f);
        }
        __syncwarp();
    }
    if (lane == 0) out[blockIdx.x * 4 + (threadIdx.x >> 5)] = b[5].x;
}

__device__ __forceinline__ uint32_t tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (b) tensor cores, 3xTF32: D[32 x 40] = C[32 x 32] . X[32 x 40]; X and D in shared memory as [band][col] floats
__global__ void __launch_bounds__(128) k_tc(const float* cmat, float* out, int iters, float* dump, const float* xin) {
    __shared__ float buf[4][32 * kStride];
    float* X = buf[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = lane; i < 32 * kStride; i += 32) X[i] = xin ? xin[i] : sinf(i * 0.37f + threadIdx.x);
    // A fragments of C (hi and lo parts): 2 row tiles x 4 k tiles x 4 registers
    uint32_t ahi[2][4][4], alo[2][4][4];
    for (int m = 0; m < 2; m++)
        for (int k = 0; k < 4; k++)
            for (int r = 0; r < 4; r++) {
                const int row = 16 * m + g + 8 * (r & 1), col = 8 * k + t + 4 * (r >> 1);
                const float c = cmat[row * 32 + col];
                ahi[m][k][r] = tf32(c);
                alo[m][k][r] = tf32(c - __uint_as_float(ahi[m][k][r]));
            }
    __syncwarp();
    for (int it = 0; it < iters; it++) {
        float d[2][5][4];
        for (int m = 0; m < 2; m++) for (int n = 0; n < 5; n++) for (int r = 0; r < 4; r++) d[m][n][r] = 0.0f;
#pragma unroll
        for (int n = 0; n < 5; n++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float x0 = X[(8 * k + t) * kStride + 8 * n + g], x1 = X[(8 * k + t + 4) * kStride + 8 * n + g];
                const uint32_t h0 = tf32(x0), h1 = tf32(x1);
                const uint32_t l0 = tf32(x0 - __uint_as_float(h0)), l1 = tf32(x1 - __uint_as_float(h1));
#pragma unroll
                for (int m = 0; m < 2; m++) {
                    mma_tf32(d[m][n], alo[m][k], h0, h1);
                    mma_tf32(d[m][n], ahi[m][k], l0, l1);
                    mma_tf32(d[m][n], ahi[m][k], h0, h1);
                }
            }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int n = 0; n < 5; n++)
#pragma unroll
                for (int r = 0; r < 4; r++)
                    X[(16 * m + g + 8 * (r >> 1)) * kStride + 8 * n + 2 * t + (r & 1)] = d[m][n][r] * (dump ? 1.0f : 0.1f);
        __syncwarp();
    }
    if (dump && blockIdx.x == 0 && threadIdx.x < 32)
        for (int i = lane; i < 32 * kStride; i += 32) dump[i] = X[i];
    if (lane == 0) out[blockIdx.x * 4 + (threadIdx.x >> 5)] = X[5];
}

int main() {
    cudaMemcpyToSymbol(c_sec, SEC, sizeof SEC);
    const float2 one = make_float2(1.0f, 1.0f);
    cudaMemcpyToSymbol(c_one2, &one, sizeof one);
    // the 32x32 matrix of the linear map, from the algorithm itself in double precision
    std::vector<float> cm(32 * 32);
    double secd[24];
    for (int i = 0; i < 24; i++) secd[i] = SEC[i];
    for (int i = 0; i < 32; i++) {
        double x[32] = {0}, y[32];
        x[i] = 1.0;
        dct32<double, double>(x, y, secd);
        for (int j = 0; j < 32; j++) cm[j * 32 + i] = (float)y[j];
    }
    float *d_c, *d_out, *d_dump;
    cudaMalloc(&d_c, sizeof(float) * 1024);
    cudaMalloc(&d_out, sizeof(float) * 148 * 16 * 4);
    cudaMalloc(&d_dump, sizeof(float) * 32 * kStride);
    cudaMemcpy(d_c, cm.data(), sizeof(float) * 1024, cudaMemcpyHostToDevice);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int blocks = 148 * 4, iters = 4000;   // 16 warps per SM, like the granule kernel
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_fma, ms_tc;
    k_fma<<<blocks, 128>>>(d_out, 10);
    cudaEventRecord(e0); k_fma<<<blocks, 128>>>(d_out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_fma, e0, e1);
    k_tc<<<blocks, 128>>>(d_c, d_out, 10, nullptr, nullptr);
    cudaEventRecord(e0); k_tc<<<blocks, 128>>>(d_c, d_out, iters, nullptr, nullptr); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_tc, e0, e1);
    const double granules = (double)blocks * 4 * iters;
    printf("DCT-32 of one granule (18 slots x 2 channels), data in shared memory, 16 warps per SM, %d MHz\n", clk_khz / 1000);
    printf("  FP32 pipe, reference algorithm, packed + exactly rounded : %8.3f ms  %7.1f clk per granule per SM-quarter  %6.2f granules/us/SM\n",
           ms_fma, ms_fma * 1e-3 * clk_khz * 1e3 / (granules / (148.0 * 4)), granules / (ms_fma * 1e3) / 148);
    printf("  tensor cores, 3xTF32 mma.sync.m16n8k8 (120 MMAs/granule)  : %8.3f ms  %7.1f clk per granule per SM-quarter  %6.2f granules/us/SM\n",
           ms_tc, ms_tc * 1e-3 * clk_khz * 1e3 / (granules / (148.0 * 4)), granules / (ms_tc * 1e3) / 148);
    printf("  ratio tensor / FP32-pipe time: %.2f\n", ms_tc / ms_fma);
    // deviation of 3xTF32 from the exactly rounded algorithm, one pass over the same data
    std::vector<float> xin(32 * kStride);
    for (size_t i = 0; i < xin.size(); i++) xin[i] = sinf((float)i * 0.37f) * 0.3f;   // PCM-like magnitudes
    float* d_x;
    cudaMalloc(&d_x, xin.size() * 4);
    cudaMemcpy(d_x, xin.data(), xin.size() * 4, cudaMemcpyHostToDevice);
    k_tc<<<1, 128>>>(d_c, d_out, 1, d_dump, d_x);
    std::vector<float> got(32 * kStride);
    cudaMemcpy(got.data(), d_dump, got.size() * 4, cudaMemcpyDeviceToHost);
    double max_rel = 0, sum2 = 0, ref2 = 0;
    int n_diff = 0, n = 0;
    for (int col = 0; col < 36; col++) {
        float x[32], y[32];
        for (int b = 0; b < 32; b++) x[b] = xin[b * kStride + col];
        dct32<float, float>(x, y, SEC);   // host float with -fmad=false: the reference's rounding sequence
        for (int j = 0; j < 32; j++) {
            const double e = (double)got[j * kStride + col] - (double)y[j];
            sum2 += e * e; ref2 += (double)y[j] * y[j];
            max_rel = fmax(max_rel, fabs(e));
            n_diff += got[j * kStride + col] != y[j];
            n++;
        }
    }
    printf("  3xTF32 vs exactly rounded: %d of %d outputs differ, max |delta| = %.3e, rms delta / rms value = %.3e (2^%.1f)\n", n_diff, n, max_rel,
           sqrt(sum2 / ref2), log2(sqrt(sum2 / ref2)));
    printf("  (the bit-exact contract needs 0 differing outputs; the tolerance mode's FMA contraction deviates by a similar 2^-24 .. 2^-22)\n");
    return 0;
}
