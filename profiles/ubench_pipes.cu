// Pipe-rate microbenchmark for the epilogue design decisions of b2_glm_tc.cu / b2_hier.cu (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_pipes profiles/ubench_pipes.cu
// Prints thread-operations per clock per SM for: FFMA (3-reg), FFMA (immediate addend, Horner form), FFMA2
// (fma.rn.f32x2, counted as 2 ops), MUFU ex2 / rcp / lg2, and two mixes that mirror the GLM epilogue
// (3 MUFU + 12 FP32 per element vs 2 MUFU + a degree-9 FFMA2 polynomial).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void k(float* out, float seed, long long* clk) {
    float x[ILP];
    float2 x2[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = seed + 0.001f * (threadIdx.x + i); x2[i] = make_float2(x[i], x[i] + 0.5f); }
    const float a = seed * 0.999f, b = seed * 1e-3f;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);
            if (MODE == 1) x[i] = fmaf(x[i], x[i], 0.25f);                 // immediate addend (Horner step shape)
            if (MODE == 2) x2[i] = __ffma2_rn(x2[i], a2, b2);
            if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (MODE == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (MODE == 6) {                                                // today's epilogue shape: 3 MUFU + ~12 FP32
                float e, inv, l;
                const float eta = x[i];
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.442695f * fabsf(eta)));
                const float w1 = 1.f + e;
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(w1));
                asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(w1));
                const float sig = eta >= 0.f ? inv : e * inv;
                x[i] = fmaf(a, eta, -fmaf(0.6931472f, l, fmaxf(eta, 0.f))) + (b - sig);
            }
            if (MODE == 7 && (i & 1) == 0) {                                // 2 MUFU + degree-9 log1p polynomial, two elements packed
                float e0, e1, i0, i1;
                const float eta0 = x[i], eta1 = x[i + 1];
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-1.442695f * fabsf(eta0)));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-1.442695f * fabsf(eta1)));
                const float2 e = make_float2(e0, e1);
                const float2 w1 = __fadd2_rn(e, make_float2(1.f, 1.f));
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(i0) : "f"(w1.x));
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(w1.y));
                float2 p = make_float2(0.0101f, 0.0101f);
#pragma unroll
                for (int d = 0; d < 9; ++d) p = __ffma2_rn(p, e, make_float2(0.1f * d - 0.5f, 0.1f * d - 0.5f));
                const float s0 = eta0 >= 0.f ? i0 : e0 * i0, s1 = eta1 >= 0.f ? i1 : e1 * i1;
                const float2 l = __ffma2_rn(p, e, make_float2(fmaxf(eta0, 0.f), fmaxf(eta1, 0.f)));
                const float2 r = __ffma2_rn(a2, make_float2(eta0, eta1), make_float2(-l.x, -l.y));
                x[i] = r.x + (b - s0);
                x[i + 1] = r.y + (b - s1);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + x2[i].x + x2[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, double ops_per_iter_elem, int blocks_per_sm, int threads) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * blocks_per_sm;
    float* out; long long* clk;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&clk, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, clk);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096];
    cudaMemcpy(h, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; ++i) mean += (double)h[i];
    mean /= blocks;
    const double elems = (double)ITERS * ILP * threads * blocks_per_sm;      // per SM
    printf("%-28s %2d blk/SM x %4d thr: %8.1f cycles/block  %7.2f elem/clk/SM  %7.2f ops/clk/SM  (%.3f ms)\n", name, blocks_per_sm,
           threads, mean, elems / mean, elems * ops_per_iter_elem / mean, ms);
    cudaFree(out); cudaFree(clk);
}

int main() {
    for (int bps = 1; bps <= 2; ++bps) {
        run<0>("FFMA 3-reg", 1, bps, 512);
        run<1>("FFMA imm addend", 1, bps, 512);
        run<2>("FFMA2 (x2 counted)", 2, bps, 512);
        run<3>("MUFU ex2", 1, bps, 512);
        run<4>("MUFU rcp", 1, bps, 512);
        run<5>("MUFU lg2", 1, bps, 512);
        run<6>("epilogue: 3 MUFU", 1, bps, 512);
        run<7>("epilogue: 2 MUFU + poly9 x2", 1, bps, 512);
    }
    return 0;
}
