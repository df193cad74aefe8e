// Reference position of the tensor-core GLM kernels (b2_glm_tc.cu, b2_glm_tcw.cu; see TcWorkspace::q_ref):
// the GEMMs work on dq = q - q_ref, the epilogue adds eta_ref = Xa . q_ref back.
#pragma once
#include "b2_engine.cuh"

// q_ref[k] = mean over the chains that will be evaluated (st == null: all n) of their pending position; one block
// per feature, fixed-order tree reduction (the result must be identical on every rank of a sharded run)
static __global__ void __launch_bounds__(256) k_glm_ref_mean(const float* qA, const float* qB, int ld, const B2ChainState* st,
                                                             int first, int n, int K1, float* __restrict__ q_ref, int kp) {
    __shared__ double s_sum[256];
    __shared__ int s_cnt[256];
    const int k = blockIdx.x;
    if (k >= kp) return;
    double acc = 0.0;
    int cnt = 0;
    if (k < K1) {
        for (int i = threadIdx.x; i < n; i += 256) {
            const int c = first + i;
            int sel = 0;
            if (st) {
                if (!b2_needs_grad(st[c].phase)) continue;
                sel = st[c].sel;
            }
            const float v = (sel ? qB : qA)[(size_t)c * ld + k];
            if (v - v == 0.f) { acc += (double)v; ++cnt; }           // finite positions only
        }
    }
    s_sum[threadIdx.x] = acc; s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s_sum[threadIdx.x] += s_sum[threadIdx.x + o]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) q_ref[k] = s_cnt[0] > 0 ? (float)(s_sum[0] / s_cnt[0]) : 0.f;
}

// eta_ref[i] = q_ref[0] * has_intercept + sum_k X[i, k] q_ref[k + off]   (fp64 accumulation, one warp per row)
static __global__ void k_glm_ref_eta(const float* __restrict__ X, int N, int K, const float* __restrict__ q_ref, int off,
                              float* __restrict__ eta_ref, int n_pad) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    double acc = 0.0;
    if (row < N) {
        const float* x = X + (size_t)row * K;
        for (int k = lane; k < K; k += 32) acc += (double)x[k] * (double)q_ref[k + off];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) eta_ref[row] = row < N ? (float)(acc + (off ? (double)q_ref[0] : 0.0)) : 0.f;
}

