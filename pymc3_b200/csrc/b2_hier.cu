// Chain-batched hierarchical linear (radon NCP) likelihood.
//
// Model: benchmarks/benchmarks/benchmarks.py:25-45:
//   A_g = mu_a + sigma_a a_g,  B_g = mu_b + sigma_b b_g,  y_i ~ N(A_{g_i} + B_{g_i} floor_i, eps)
// Gradient needs S0_g = sum_{i in g} r_i, S1_g = sum_{i in g} r_i floor_i, SS = sum r_i^2.
//
// Layout: observations are sorted by group once at model build (CSR offsets grp_off), so the
// reference's fancy-index gather / AdvancedIncSubtensor scatter becomes a segmented
// reduction with no atomics.  One thread owns one chain; a block of 128 chains walks a slab
// of observations staged through shared memory, so each observation byte is read from
// HBM/L2 once per 128 chain-gradients (bytes = 6 N / Cb per chain-grad, Cb=128)
// and every (chain, observation) pair still costs its ~6 flops: all N observations are
// streamed for every chain -- the Gaussian sufficient-statistics shortcut is NOT used.
#include "b2_engine.cuh"

#define HT_CHAINS 128
#define HT_TILE 2048

// Launch geometry, derived on the device from the number of chains that still need a gradient: the live
// chains are packed into dense 128-chain tiles (k_hier_compact) and the fixed grid is re-divided into as many
// observation slabs per tile as fit, so a launch costs what its live chains cost (early in tuning and at the
// end of a run most chains wait for a few deep trees).
struct HierGeom { int n_act, nt, sp, rows_per_split, stride; };
__device__ __forceinline__ HierGeom hier_geom(const int* cnt, int grid_ctas, int N) {
    HierGeom g;
    g.n_act = *cnt;
    g.nt = (g.n_act + HT_CHAINS - 1) / HT_CHAINS;
    int sp = g.nt > 0 ? grid_ctas / g.nt : 1;
    const int max_sp = (N + 255) / 256;
    if (sp > max_sp) sp = max_sp;
    if (sp < 1) sp = 1;
    g.rows_per_split = (N + sp - 1) / sp;
    g.sp = (N + g.rows_per_split - 1) / g.rows_per_split;
    g.stride = g.nt * HT_CHAINS;
    return g;
}

// slot <-> chain map of the chains that need a gradient (st == null: parity hook, every chain, identity order)
__global__ void k_hier_compact(const B2ChainState* st, int n_chains, int* cnt, int* chain_of_slot) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    if (!st) { chain_of_slot[c] = c; if (c == 0) *cnt = n_chains; return; }
    if (st[c].phase > B2_PHASE_HMC) return;
    chain_of_slot[atomicAdd(cnt, 1)] = c;
}

// Per-thread running sums of one (chain, group) segment: S0 = sum r, S1 = sum r f, SS = sum r^2.
// fp64 check build: plain double.  fp32 production build: the FP32 pipe is the bound of this kernel (5 flops per
// (chain, observation)), so the inner loop uses Blackwell's packed fp32 math (FADD2 / FFMA2: two
// observations per instruction, 2.5 issue slots per pair instead of 5) into four independent lanes per sum, and
// the lanes are folded into DOUBLE totals at the end of every staged tile (<= HT_TILE rows, <= 512 addends per
// lane) and at every group end -- a slab of 27 k rows (C3) is therefore never summed in fp32 end to end.
template <typename T> struct HierAcc;
template <> struct HierAcc<double> {
    double s0, s1, ss;
    __device__ __forceinline__ void clear() { s0 = 0.0; s1 = 0.0; ss = 0.0; }
    __device__ __forceinline__ void set_group(double, double) {}
    __device__ __forceinline__ void one(double A, double B, float y, float f) {
        const double r = (double)y - (A + B * (double)f);
        s0 += r; s1 += r * (double)f; ss += r * r;
    }
    __device__ __forceinline__ void four(double A, double B, const float4& y4, const float4& f4) {
        one(A, B, y4.x, f4.x); one(A, B, y4.y, f4.y); one(A, B, y4.z, f4.z); one(A, B, y4.w, f4.w);
    }
    __device__ __forceinline__ void fold(double& d0, double& d1, double& dss) { d0 += s0; d1 += s1; dss += ss; clear(); }
};
template <> struct HierAcc<float> {
    float2 a0, a1, a2, b0, b1, b2;            // lanes {x, y} of pair a (observations 0,1) and pair b (2,3)
    float2 nA, nB;                            // (-A, -A), (-B, -B) of the current group
    __device__ __forceinline__ void clear() {
        a0 = a1 = a2 = b0 = b1 = b2 = make_float2(0.f, 0.f);
    }
    __device__ __forceinline__ void set_group(float A, float B) { nA = make_float2(-A, -A); nB = make_float2(-B, -B); }
    __device__ __forceinline__ void one(float A, float B, float y, float f) {
        const float r = y - (A + B * f);
        a0.x += r; a1.x = fmaf(r, f, a1.x); a2.x = fmaf(r, r, a2.x);
    }
    __device__ __forceinline__ void four(float, float, const float4& y4, const float4& f4) {
        const float2 fa = make_float2(f4.x, f4.y), fb = make_float2(f4.z, f4.w);
        const float2 ra = __ffma2_rn(nB, fa, __fadd2_rn(make_float2(y4.x, y4.y), nA));     // r = (y - A) - B f
        const float2 rb = __ffma2_rn(nB, fb, __fadd2_rn(make_float2(y4.z, y4.w), nA));
        a0 = __fadd2_rn(a0, ra); a1 = __ffma2_rn(ra, fa, a1); a2 = __ffma2_rn(ra, ra, a2);
        b0 = __fadd2_rn(b0, rb); b1 = __ffma2_rn(rb, fb, b1); b2 = __ffma2_rn(rb, rb, b2);
    }
    __device__ __forceinline__ void fold(double& d0, double& d1, double& dss) {
        d0 += (double)((a0.x + a0.y) + (b0.x + b0.y));
        d1 += (double)((a1.x + a1.y) + (b1.x + b1.y));
        dss += (double)((a2.x + a2.y) + (b2.x + b2.y));
        clear();
    }
};

template <typename T>
__global__ void __launch_bounds__(HT_CHAINS)
k_hier_slab(const float* __restrict__ y, const unsigned char* __restrict__ fl, const int* __restrict__ grp_off,
            int N, int NG, const T* qA, const T* qB, int ld, const B2ChainState* st, const int* cnt,
            const int* __restrict__ chain_of_slot, T* __restrict__ part /* [split][2*NG+1][stride] */) {
    __shared__ __align__(16) float ys[HT_TILE];
    __shared__ __align__(16) float fs[HT_TILE];     // covariate converted u8 -> float once per tile, not per (chain, obs)
    __shared__ int s_g0;
    const HierGeom gm = hier_geom(cnt, gridDim.x, N);
    if ((int)blockIdx.x >= gm.nt * gm.sp) return;
    const int tid = threadIdx.x;
    const int ctile = blockIdx.x / gm.sp, split = blockIdx.x % gm.sp;
    const int slot = ctile * HT_CHAINS + tid;
    const int rows_per_split = gm.rows_per_split;
    const int row_begin = split * rows_per_split;
    const int row_end = min(N, row_begin + rows_per_split);
    const bool live = slot < gm.n_act;
    const int chain = live ? chain_of_slot[slot] : 0;
    const int n_chains = gm.stride;                  // partials are indexed by slot
    int sel = 0;
    if (live && st) sel = st[chain].sel;
    const T* q = (sel ? qB : qA) + (size_t)chain * ld;
    const T mu_a = q[0], mu_b = q[2];
    const T sa = (T)exp((double)q[1]), sb = (T)exp((double)q[3]);
    if (tid == 0) {                                    // first group intersecting the slab
        int lo = 0, hi = NG - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (grp_off[mid] <= row_begin) lo = mid; else hi = mid - 1; }
        s_g0 = lo;
    }
    __syncthreads();
    int g = s_g0;
    int g_end = grp_off[g + 1];
    T A = mu_a + sa * q[4 + g], B = mu_b + sb * q[4 + NG + g];
    HierAcc<T> acc;
    acc.clear();
    acc.set_group(A, B);
    double d0 = 0.0, d1 = 0.0, dss = 0.0;             // this group's S0, S1 so far; the slab's SS
    bool dirty = false;                                // rows accumulated since the last flush (block-uniform)
    T* pbase = part + (size_t)split * (2 * NG + 1) * n_chains;
    for (int t0 = row_begin; t0 < row_end; t0 += HT_TILE) {
        const int cnt = min(HT_TILE, row_end - t0);
        __syncthreads();
        for (int i = tid; i < cnt; i += HT_CHAINS) { ys[i] = y[t0 + i]; fs[i] = (float)fl[t0 + i]; }
        __syncthreads();
        int i = 0;
        while (i < cnt) {
            const int stop = min(cnt, g_end - t0);      // block-uniform
            if (i < stop) dirty = true;
            // scalar head up to a 4-aligned index, then 4 observations per shared-memory transaction
            // (one LDS.128 of y + one LDS.128 of the covariate, both broadcast to the warp)
            for (; i < stop && (i & 3); ++i) acc.one(A, B, ys[i], fs[i]);
#pragma unroll 2
            for (; i + 4 <= stop; i += 4)
                acc.four(A, B, *reinterpret_cast<const float4*>(ys + i), *reinterpret_cast<const float4*>(fs + i));
            for (; i < stop; ++i) acc.one(A, B, ys[i], fs[i]);
            acc.fold(d0, d1, dss);                      // end of this (tile, group) segment: lanes -> double totals
            if (t0 + i == g_end) {                      // group complete: flush and move on
                if (live && dirty) {
                    pbase[(size_t)g * n_chains + slot] = (T)d0;
                    pbase[(size_t)(NG + g) * n_chains + slot] = (T)d1;
                }
                d0 = 0.0; d1 = 0.0; dirty = false;
                ++g;
                if (g < NG) {
                    g_end = grp_off[g + 1];
                    A = mu_a + sa * q[4 + g]; B = mu_b + sb * q[4 + NG + g];
                    acc.set_group(A, B);
                } else {
                    g_end = 0x7fffffff;
                }
            }
        }
    }
    if (live) {
        if (dirty && g < NG) {                         // group cut by the slab boundary
            pbase[(size_t)g * n_chains + slot] = (T)d0;
            pbase[(size_t)(NG + g) * n_chains + slot] = (T)d1;
        }
        pbase[(size_t)(2 * NG) * n_chains + slot] = (T)dss;
    }
}

// warp per chain: combine slab partials (fixed order), add priors, chain rule to the free variables
template <typename T>
__global__ void k_hier_finalize(const T* __restrict__ part, const int* __restrict__ grp_off, int N, int NG,
                                const int* cnt, const int* __restrict__ chain_of_slot, int grid_ctas,
                                double mu_sd, double hc_beta,
                                const T* qA, const T* qB, T* gA, T* gB, int ld, const B2ChainState* st, double* logp) {
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const HierGeom gm = hier_geom(cnt, grid_ctas, N);
    if (slot >= gm.n_act) return;
    const int chain = chain_of_slot[slot];
    const int n_splits = gm.sp, rows_per_split = gm.rows_per_split, n_chains = gm.stride;
    const int sel = st ? st[chain].sel : 0;
    const T* q = (sel ? qB : qA) + (size_t)chain * ld;
    T* gr = (sel ? gB : gA) + (size_t)chain * ld;
    const double mu_a = (double)q[0], ua = (double)q[1], mu_b = (double)q[2], ub = (double)q[3];
    const double ue = (double)q[4 + 2 * NG];
    const double sa = exp(ua), sb = exp(ub), eps = exp(ue);
    const double inv_e2 = 1.0 / (eps * eps);
    const size_t pstride = (size_t)(2 * NG + 1) * n_chains;
    double acc[6] = {0, 0, 0, 0, 0, 0};               // sumS0, sumS1, S0.a, S1.b, prior(a,b), ss
    for (int g = lane; g < NG; g += 32) {
        const int lo = grp_off[g], hi = grp_off[g + 1];
        double s0 = 0.0, s1 = 0.0;
        if (hi > lo) {
            const int sp_lo = lo / rows_per_split, sp_hi = (hi - 1) / rows_per_split;
            for (int sp = sp_lo; sp <= sp_hi; ++sp) {
                s0 += (double)part[sp * pstride + (size_t)g * n_chains + slot];
                s1 += (double)part[sp * pstride + (size_t)(NG + g) * n_chains + slot];
            }
        }
        s0 *= inv_e2; s1 *= inv_e2;
        const double a = (double)q[4 + g], b = (double)q[4 + NG + g];
        acc[0] += s0; acc[1] += s1; acc[2] += s0 * a; acc[3] += s1 * b;
        acc[4] += -0.5 * (a * a + b * b) - B2_LOG_2PI;
        gr[4 + g] = (T)(-a + sa * s0);
        gr[4 + NG + g] = (T)(-b + sb * s1);
    }
    for (int sp = lane; sp < n_splits; sp += 32) acc[5] += (double)part[sp * pstride + (size_t)(2 * NG) * n_chains + slot];
#pragma unroll
    for (int k = 0; k < 6; ++k)
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0) {
        double dua, dub, due;
        double lp = acc[4] + b2_normal_logp(mu_a, 0.0, mu_sd) + b2_normal_logp(mu_b, 0.0, mu_sd);
        lp += b2_halfcauchy_log_logp(ua, hc_beta, &dua) + b2_halfcauchy_log_logp(ub, hc_beta, &dub) +
              b2_halfcauchy_log_logp(ue, hc_beta, &due);
        lp += -0.5 * inv_e2 * acc[5] + (double)N * (-ue - 0.5 * B2_LOG_2PI);
        const double pm = 1.0 / (mu_sd * mu_sd);
        gr[0] = (T)(-mu_a * pm + acc[0]);
        gr[1] = (T)(dua + sa * acc[2]);
        gr[2] = (T)(-mu_b * pm + acc[1]);
        gr[3] = (T)(dub + sb * acc[3]);
        gr[4 + 2 * NG] = (T)(due - (double)N + acc[5] * inv_e2);
        logp[chain] = lp;
    }
}

template <typename T>
int b2_hier_launch(b2_engine* e, const T* qA, const T* qB, T* gA, T* gB, int ld,
                   const B2ChainState* st, int n, double* logp, cudaStream_t stream) {
    const int N = e->md.N, NG = e->md.G;
    const int chain_tiles = (n + HT_CHAINS - 1) / HT_CHAINS;
    // fixed grid: ~8 resident 128-thread blocks per SM, a whole number of waves (148 SMs); the kernels divide
    // it into (live chain tiles) x (observation slabs) themselves
    int grid_ctas = 8 * e->sm_count;
    if (grid_ctas < chain_tiles) grid_ctas = chain_tiles;
    const size_t part_bytes = (size_t)grid_ctas * (2 * NG + 1) * HT_CHAINS * sizeof(T);
    const size_t map_off = (part_bytes + 255) & ~(size_t)255;
    const size_t need = map_off + ((size_t)e->C + 64) * sizeof(int);
    if (e->hier_ws_bytes < need) {
        if (e->hier_ws) cudaFree(e->hier_ws);
        e->hier_ws = nullptr; e->hier_ws_bytes = 0;
        B2_CUDA_OK(cudaMalloc(&e->hier_ws, need));
        e->hier_ws_bytes = need;
    }
    int* cnt = reinterpret_cast<int*>((char*)e->hier_ws + map_off);
    int* chain_of_slot = cnt + 64;
    B2_CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(int), stream));
    k_hier_compact<<<(n + 255) / 256, 256, 0, stream>>>(st, n, cnt, chain_of_slot);
    B2_CUDA_OK(cudaGetLastError());
    k_hier_slab<T><<<grid_ctas, HT_CHAINS, 0, stream>>>(e->md.yf, e->md.floor_u8, e->md.grp_off, N, NG, qA, qB, ld, st,
                                                        cnt, chain_of_slot, (T*)e->hier_ws);
    B2_CUDA_OK(cudaGetLastError());
    k_hier_finalize<T><<<(n + 3) / 4, 128, 0, stream>>>((const T*)e->hier_ws, e->md.grp_off, N, NG, cnt, chain_of_slot,
                                                         grid_ctas, e->md.hp[0], e->md.hp[1], qA, qB, gA, gB, ld, st, logp);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 1;
    e->launches += 2;
    return 0;
}

template int b2_hier_launch<float>(b2_engine*, const float*, const float*, float*, float*, int,
                                   const B2ChainState*, int, double*, cudaStream_t);
template int b2_hier_launch<double>(b2_engine*, const double*, const double*, double*, double*, int,
                                    const B2ChainState*, int, double*, cudaStream_t);
