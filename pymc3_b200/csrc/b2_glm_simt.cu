// Chain-batched Bernoulli-logit GLM likelihood, SIMT (FP32/FP64 FMA) version.
//
// logp_c = sum_i [y_i eta_ic - softplus(eta_ic)] + prior,  eta = b0_c + X beta_c
// grad_c = X^T (y - sigmoid(eta_c)) - tau beta_c           (SURVEY appendix C, C2/C5;
// reference: pymc3/glm/linear.py:49-101, glm/families.py:115-119, discrete.py:104,350)
//
// One launch evaluates every chain: a block owns a 64-chain x (row range) slab, streams 64-row
// tiles of X through shared memory and runs two register-tiled products per tile
//   phase 1  Eta[64 rows, 64 chains] = Xtile . B        (4x4 micro-tiles)
//   phase 2  G[K1, 64 chains]       += Xtile^T . R       (R = y - sigmoid(Eta), kept on chip)
// so X is read from HBM/L2 once per leapfrog for all chains and R never leaves the SM.
// This is the fp64 check-build path and the fallback for shapes the tcgen05 kernel
// (b2_glm_tc.cu) does not take.  Partial sums per row split are reduced in a fixed order by
// k_glm_finalize, so results are bit-reproducible run to run.
#include "b2_engine.cuh"

#define GT_ROWS 64
#define GT_CHAINS 64
#define GT_STRIDE 68          // k-major row stride (elements): 16B-aligned, spreads banks

template <typename T> struct Vec4 { T x, y, z, w; };
template <typename T> __device__ __forceinline__ Vec4<T> ld4(const T* p) {
    Vec4<T> v; v.x = p[0]; v.y = p[1]; v.z = p[2]; v.w = p[3]; return v;
}
template <> __device__ __forceinline__ Vec4<float> ld4<float>(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    Vec4<float> v; v.x = t.x; v.y = t.y; v.z = t.z; v.w = t.w; return v;
}
template <> __device__ __forceinline__ Vec4<double> ld4<double>(const double* p) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    Vec4<double> v; v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y; return v;
}

__device__ __forceinline__ float b2_exp(float x) { return expf(x); }
__device__ __forceinline__ double b2_exp(double x) { return exp(x); }
__device__ __forceinline__ float b2_log1p(float x) { return log1pf(x); }
__device__ __forceinline__ double b2_log1p(double x) { return log1p(x); }

// y*eta - softplus(eta) and y - sigmoid(eta), overflow-safe
template <typename T>
__device__ __forceinline__ void logit_terms(T eta, T y, T& ll, T& r) {
    const T e = b2_exp(-fabs(eta));
    const T inv = (T)1 / ((T)1 + e);
    const T sig = eta >= (T)0 ? inv : e * inv;
    ll = y * eta - ((eta > (T)0 ? eta : (T)0) + b2_log1p(e));
    r = y - sig;
}

template <typename T, int KPT>
__global__ void __launch_bounds__(256)
k_glm_simt(const float* __restrict__ X, const float* __restrict__ y, int N, int K,
           const T* qA, const T* qB, int ld, const B2ChainState* st, int n_chains,
           int rows_per_split, double* __restrict__ gpart, double* __restrict__ lpart) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K1 = K + 1;
    T* Bs = reinterpret_cast<T*>(smem_raw);            // [K1][GT_STRIDE]   (k, chain)
    T* Xt = Bs + (size_t)K1 * GT_STRIDE;               // [K1][GT_STRIDE]   (k, row)
    T* Rs = Xt + (size_t)K1 * GT_STRIDE;               // [GT_ROWS][GT_STRIDE] (row, chain)
    T* ys = Rs + (size_t)GT_ROWS * GT_STRIDE;          // [GT_ROWS]
    T* lps = ys + GT_ROWS;                             // [16][GT_CHAINS]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int cb = blockIdx.x * GT_CHAINS;
    const int split = blockIdx.y;
    const int row_begin = split * rows_per_split;
    const int row_end = min(N, row_begin + rows_per_split);

    // coefficients of this block's chains: Bs[k][c]; chains that need no gradient get zeros
    for (int idx = tid; idx < GT_CHAINS * K1; idx += 256) {
        const int c = idx / K1, k = idx - c * K1;
        const int chain = cb + c;
        T v = (T)0;
        if (chain < n_chains) {
            int sel = 0;
            bool live = true;
            if (st) { const int ph = st[chain].phase; live = ph <= B2_PHASE_HMC; sel = st[chain].sel; }
            if (live) v = (sel ? qB : qA)[(size_t)chain * ld + k];
        }
        Bs[k * GT_STRIDE + c] = v;
    }
    T gacc[KPT][4];
#pragma unroll
    for (int m = 0; m < KPT; ++m) { gacc[m][0] = gacc[m][1] = gacc[m][2] = gacc[m][3] = (T)0; }
    T lacc[4] = {(T)0, (T)0, (T)0, (T)0};

    for (int r0 = row_begin; r0 < row_end; r0 += GT_ROWS) {
        __syncthreads();                               // previous tile fully consumed (and Bs ready)
        for (int idx = tid; idx < GT_ROWS * K; idx += 256) {
            const int r = idx / K, kk = idx - r * K;
            const int row = r0 + r;
            Xt[(kk + 1) * GT_STRIDE + r] = row < row_end ? (T)X[(size_t)row * K + kk] : (T)0;
        }
        if (tid < GT_ROWS) {
            const int row = r0 + tid;
            Xt[tid] = row < row_end ? (T)1 : (T)0;     // intercept column
            ys[tid] = row < row_end ? (T)y[row] : (T)0;
        }
        __syncthreads();
        // phase 1: eta micro-tile rows 4ty.. x chains 4tx..
        T acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = (T)0; }
        for (int k = 0; k < K1; ++k) {
            const Vec4<T> a = ld4<T>(Xt + k * GT_STRIDE + 4 * ty);
            const Vec4<T> b = ld4<T>(Bs + k * GT_STRIDE + 4 * tx);
            const T av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = 4 * ty + i;
            const bool valid = (r0 + r) < row_end;
            const T yy = ys[r];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                T ll, rr;
                logit_terms<T>(acc[i][j], yy, ll, rr);
                if (!valid) { ll = (T)0; rr = (T)0; }
                lacc[j] += ll;
                Rs[r * GT_STRIDE + 4 * tx + j] = rr;
            }
        }
        __syncthreads();
        // phase 2: G[k][chain] += sum_row X[row][k] R[row][chain],  k = ty + 16 m
        for (int r = 0; r < GT_ROWS; ++r) {
            const Vec4<T> rv = ld4<T>(Rs + r * GT_STRIDE + 4 * tx);
#pragma unroll
            for (int m = 0; m < KPT; ++m) {
                const int k = ty + 16 * m;
                if (k < K1) {
                    const T x = Xt[k * GT_STRIDE + r];
                    gacc[m][0] += x * rv.x; gacc[m][1] += x * rv.y; gacc[m][2] += x * rv.z; gacc[m][3] += x * rv.w;
                }
            }
        }
    }
    // write partials
#pragma unroll
    for (int m = 0; m < KPT; ++m) {
        const int k = ty + 16 * m;
        if (k < K1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chain = cb + 4 * tx + j;
                if (chain < n_chains) gpart[((size_t)split * n_chains + chain) * K1 + k] = (double)gacc[m][j];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) lps[ty * GT_CHAINS + 4 * tx + j] = lacc[j];
    __syncthreads();
    if (tid < GT_CHAINS) {
        double s = 0.0;
        for (int t = 0; t < 16; ++t) s += (double)lps[t * GT_CHAINS + tid];
        const int chain = cb + tid;
        if (chain < n_chains) lpart[(size_t)split * n_chains + chain] = s;
    }
}

// fixed-order reduction over row splits + prior; writes grad into the chain's pending plane
template <typename T>
__global__ void k_glm_finalize(const double* __restrict__ gpart, const double* __restrict__ lpart, int n_splits,
                               int n_chains, int K1, double prior_tau, const T* qA, const T* qB, T* gA, T* gB,
                               int ld, const B2ChainState* st, double* logp) {
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (chain >= n_chains) return;
    int sel = 0;
    if (st) {
        if (st[chain].phase > B2_PHASE_HMC) return;
        sel = st[chain].sel;
    }
    const T* q = (sel ? qB : qA) + (size_t)chain * ld;
    T* g = (sel ? gB : gA) + (size_t)chain * ld;
    double prior = 0.0;
    for (int k = lane; k < K1; k += 32) {
        double s = 0.0;
        for (int sp = 0; sp < n_splits; ++sp) s += gpart[((size_t)sp * n_chains + chain) * K1 + k];
        if (k > 0) {
            const double b = (double)q[k];
            s -= prior_tau * b;
            prior += 0.5 * (-prior_tau * b * b + log(prior_tau) - B2_LOG_2PI);
        }
        g[k] = (T)s;
    }
    double lp = 0.0;
    for (int sp = lane; sp < n_splits; sp += 32) lp += lpart[(size_t)sp * n_chains + chain];
    for (int o = 16; o > 0; o >>= 1) {
        prior += __shfl_xor_sync(0xffffffffu, prior, o);
        lp += __shfl_xor_sync(0xffffffffu, lp, o);
    }
    if (lane == 0) logp[chain] = lp + prior;
}

template <typename T>
int b2_glm_simt_launch(b2_engine* e, const T* qA, const T* qB, T* gA, T* gB, int ld,
                       const B2ChainState* st, int n, double* logp, cudaStream_t stream) {
    const int K = e->md.G, K1 = K + 1, N = e->md.N;
    const int chain_tiles = (n + GT_CHAINS - 1) / GT_CHAINS;
    int splits = (2 * e->sm_count + chain_tiles - 1) / chain_tiles;
    const int max_splits = (N + GT_ROWS - 1) / GT_ROWS;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rows_per_split = (N + splits - 1) / splits;
    rows_per_split = ((rows_per_split + GT_ROWS - 1) / GT_ROWS) * GT_ROWS;
    splits = (N + rows_per_split - 1) / rows_per_split;
    const size_t need = ((size_t)splits * e->C * K1 + (size_t)splits * e->C) * sizeof(double);
    if (e->glm_ws_bytes < need) {
        if (e->glm_ws) cudaFree(e->glm_ws);
        e->glm_ws = nullptr; e->glm_ws_bytes = 0;
        B2_CUDA_OK(cudaMalloc(&e->glm_ws, need));
        e->glm_ws_bytes = need;
    }
    double* gpart = (double*)e->glm_ws;
    double* lpart = gpart + (size_t)splits * n * K1;
    const size_t smem = ((size_t)2 * K1 * GT_STRIDE + (size_t)GT_ROWS * GT_STRIDE + GT_ROWS + 16 * GT_CHAINS) * sizeof(T);
    if (smem > 227 * 1024) { b2_set_error("GLM SIMT kernel: too many regressors for shared memory"); return -8; }
    dim3 grid(chain_tiles, splits);
#define B2_LAUNCH_GLM(KPT)                                                                              \
    do {                                                                                                \
        B2_CUDA_OK(cudaFuncSetAttribute(k_glm_simt<T, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_glm_simt<T, KPT><<<grid, 256, smem, stream>>>(e->md.X, e->md.yf, N, K, qA, qB, ld, st, n,     \
                                                        rows_per_split, gpart, lpart);                 \
    } while (0)
    if (K1 <= 16 * 2) B2_LAUNCH_GLM(2);
    else if (K1 <= 16 * 8) B2_LAUNCH_GLM(8);
    else if (K1 <= 16 * 17) B2_LAUNCH_GLM(17);
    else { b2_set_error("GLM SIMT kernel supports at most 271 regressors"); return -8; }
#undef B2_LAUNCH_GLM
    B2_CUDA_OK(cudaGetLastError());
    k_glm_finalize<T><<<(n + 3) / 4, 128, 0, stream>>>(gpart, lpart, splits, n, K1, e->md.hp[0], qA, qB, gA, gB, ld, st, logp);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 2;
    return 0;
}

template int b2_glm_simt_launch<float>(b2_engine*, const float*, const float*, float*, float*, int,
                                       const B2ChainState*, int, double*, cudaStream_t);
template int b2_glm_simt_launch<double>(b2_engine*, const double*, const double*, double*, double*, int,
                                        const B2ChainState*, int, double*, cudaStream_t);
