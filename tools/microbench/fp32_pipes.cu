// Micro-benchmark: issue throughput of scalar FMUL / FADD, packed FADD2 / FMUL2 / FFMA2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o fp32_pipes fp32_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s0, float s1) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 m = make_float2(s0, s1), nz = make_float2(-0.0f, -0.0f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) { a[i].x = __fmul_rn(a[i].x, s0); a[i].y = __fmul_rn(a[i].y, s1); }          // 2 FMUL
                if (MODE == 1) { a[i].x = __fadd_rn(a[i].x, s0); a[i].y = __fadd_rn(a[i].y, s1); }          // 2 FADD
                if (MODE == 2) { a[i] = __fadd2_rn(a[i], m); }                                              // 1 FADD2
                if (MODE == 3) { a[i] = __ffma2_rn(a[i], m, nz); }                                          // 1 FFMA2 (exact packed multiply)
                if (MODE == 4) { a[i] = __fmul2_rn(a[i], m); }                                              // 1 FMUL2
                if (MODE == 5) { a[i].x = __fmul_rn(a[i].x, s0); a[i].y = __fmul_rn(a[i].y, s1); a[i] = __fadd2_rn(a[i], m); }   // 2 FMUL + FADD2
                if (MODE == 6) { a[i] = __ffma2_rn(a[i], m, nz); a[i] = __fadd2_rn(a[i], m); }              // FFMA2 + FADD2
                if (MODE == 7) { a[i].x = __fmul_rn(a[i].x, 1.0001f); a[i].y = __fmul_rn(a[i].y, 0.9999f); } // 2 FMUL imm
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int instr_per_elem, float* d) {
    const int iters = 2000, blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 10, 1.0001f, 0.9999f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 1.0001f, 0.9999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double elems = (double)blocks * 256 * iters * 4 * 8;     // float2 element updates
    const double winst = elems / 32 * instr_per_elem;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-22s %8.3f ms  %6.3f warp-inst/clk/SM  %6.2f float2-updates/clk/SM\n", name, ms, winst / cycles / 148, elems / cycles / 148);
}

int main() {
    float* d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("2xFMUL (reg)", 2, d);
    run<7>("2xFMUL (imm)", 2, d);
    run<1>("2xFADD", 2, d);
    run<2>("FADD2", 1, d);
    run<3>("FFMA2(a,b,-0)", 1, d);
    run<4>("FMUL2", 1, d);
    run<5>("2xFMUL+FADD2", 3, d);
    run<6>("FFMA2+FADD2", 2, d);
    return 0;
}
