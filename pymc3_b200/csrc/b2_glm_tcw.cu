// Wide variant of the tensor-core Bernoulli-logit GLM likelihood: 128 <= K <= 256 feature columns
// (config C5: K = 256 features + intercept; reference: pymc3/glm/linear.py:49-101, glm/families.py:115-119,
// evaluated by ValueGradFunction.__call__, model.py:645-666).  Same algorithm as b2_glm_tc.cu
//
//   GEMM1  S[128 chains, 32 obs]    = Q . Xtile^T        (K up to 256: 16 K steps x 3 bf16 split passes)
//   epi    R = y - sigmoid(S + q0),  logp += ...          (the intercept q0 is added here, not in the GEMM)
//   GEMM2  G[128 chains, 128 feat] += R . Xtile[:, half]  (MN-major read of the SAME shared-memory tile)
//
// with the tensor-memory budget re-cut for twice the features.  512 TMEM columns hold S 2x32, R 2x32
// (bf16 hi | lo), Q 256 (bf16 hi | lo of 256 features) and ONE 128-feature half of G, so a CTA owns
// (chain tile, observation slab, feature half): the two CTAs of a pair walk the same X tiles (second reader
// hits L2), both run GEMM1 and the epilogue, each accumulates its half of the gradient.  Tensor work per
// 32-observation tile and CTA: GEMM1 768 + GEMM2 384 cycles against 768 + 768 for an (impossible) unsplit
// CTA, i.e. 2/3 efficiency from the duplicated GEMM1.
// A tile is 32 observations so that a pipeline stage stays one 32 KB bulk copy (+128 B of y) and S / R can be
// double-buffered; its 54 MMAs are small (N = 32: 16 cycles of tensor work), so the issue loop must be cheap:
// descriptors are `per-tile base + constant` (see mma_ts_split), the K-step count is a template parameter.
// (Measured, 3 M rows x 256 chains: 3.74 ms with 64-bit descriptors rebuilt per MMA (~35 cycles per MMA whatever
// its N), 2.89 ms for a 64-observation single-buffered variant, 2.67 ms for this one.)
// The intercept never enters the GEMMs (K = 256 would become 257): eta = S + q0 in the epilogue, and its
// gradient is the row sum of R, accumulated next to logp.  Zero-padded rows carry y = -1 and are masked.
#include <cstring>
#include <cstdlib>
#include "b2_engine.cuh"
#include "b2_tc_ptx.cuh"
#include "b2_glm_ref.cuh"

#define TW_CHAINS 128
#define TW_OBS 32
#define TW_KP 256
#define TW_STAGES 6
#define TW_XPART_BYTES (TW_OBS * TW_KP * 2)             // 16384: one of {hi, lo}, four 64-column atoms of 32 rows
#define TW_Y_BYTES (TW_OBS * 4)                         // 128
#define TW_STAGE_DATA (2 * TW_XPART_BYTES + TW_Y_BYTES) // 32896 in global memory: Xhi | Xlo | y
#define TW_STAGE_BYTES (2 * TW_XPART_BYTES)
#define TW_YS_BYTES (2 * TW_Y_BYTES)                    // 256 in shared memory: y | eta_ref (b2_glm_tc.cu, TcWorkspace::q_ref)
#define TW_SMEM_BYTES (1024 + TW_STAGES * (TW_STAGE_BYTES + TW_YS_BYTES) + 512)
#define TW_EPI_GROUPS 4                                 // epilogue warpgroups; group g takes tiles t = g (mod 4)
#define TW_EPI_WARPS (4 * TW_EPI_GROUPS)
#define TW_THREADS (128 + 32 * TW_EPI_WARPS)
#define TW_TMEM_COLS 512
#define TW_COL_S 0                                      // S[b] at 32 b
#define TW_COL_P 64                                     // R[b] at 64 + 32 b: hi 16 cols | lo 16 cols
#define TW_COL_G 128                                    // 128 columns (this CTA's feature half)
#define TW_COL_Q 256                                    // hi 128 cols | lo 128 cols

struct TwWorkspace {
    unsigned char* xt;       // [n_tiles][TW_STAGE_DATA]
    float* gpart;            // [slab][slot][TW_KP]
    double* lpart;           // [slab][TW_EPI_GROUPS][slot]
    double* rpart;           // [slab][TW_EPI_GROUPS][slot]   row sums of R (intercept gradient)
    int n_tiles, chain_tiles, grid_ctas;
    int* err;
    const float* qA; const float* qB; int ld; const B2ChainState* st; int n_chains; int K;
    int* counter;            // live chains of this launch
    int* chain_of_slot;
    const float* q_ref;      // [1 + TW_KP] reference position (intercept first): the GEMM sees q - q_ref ...
    const float* eta_ref;    // [n_tiles * TW_OBS] ... and the epilogue adds eta_ref = q_ref[0] + X . q_ref[1:] back
    int* x_flags;            // [0] != 0: some X element has a non-zero bf16 low part (set by the tiling kernel)
    int n_pass;              // 3 split passes, or 2 when X is bf16-representable (Xlo == 0: config C5's generator)
    int flush_tiles;         // the gradient accumulator is drained into the fp32 partials every flush_tiles tiles
};

struct TwGeom { int n_act, nt, sp, tps, stride; };
__device__ __forceinline__ TwGeom tw_geom(const TwWorkspace& ws) {
    TwGeom g;
    g.n_act = *ws.counter;
    g.nt = (g.n_act + TW_CHAINS - 1) / TW_CHAINS;
    g.sp = g.nt > 0 ? (ws.grid_ctas / 2) / g.nt : 1;
    if (g.sp > ws.n_tiles) g.sp = ws.n_tiles;
    if (g.sp < 1) g.sp = 1;
    g.tps = (ws.n_tiles + g.sp - 1) / g.sp;
    g.sp = (ws.n_tiles + g.tps - 1) / g.tps;
    g.stride = g.nt * TW_CHAINS;
    return g;
}

__global__ void k_glm_tcw_prep_x(const float* __restrict__ X, const float* __restrict__ y, int N, int K,
                                 unsigned char* __restrict__ xt, int* __restrict__ x_flags) {
    const int tile = blockIdx.x;
    unsigned char* blob = xt + (size_t)tile * TW_STAGE_DATA;
    bool any_lo = false;
    for (int idx = threadIdx.x; idx < TW_OBS * TW_KP; idx += blockDim.x) {
        const int r = idx / TW_KP, c = idx - r * TW_KP;
        const int row = tile * TW_OBS + r;
        const float v = (row < N && c < K) ? X[(size_t)row * K + c] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int off = (c >> 6) * (TW_OBS * 128) + tc_swz(r, c & 63);
        *reinterpret_cast<__nv_bfloat16*>(blob + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(blob + TW_XPART_BYTES + off) = lo;
        any_lo = any_lo || (__bfloat162float(lo) != 0.f);
    }
    if (__syncthreads_or(any_lo) && threadIdx.x == 0) atomicOr(x_flags, 1);
    for (int r = threadIdx.x; r < TW_OBS; r += blockDim.x) {
        const int row = tile * TW_OBS + r;
        reinterpret_cast<float*>(blob + 2 * TW_XPART_BYTES)[r] = row < N ? y[row] : -1.f;   // -1: padding row
    }
}

__global__ void k_glm_tcw_compact(const B2ChainState* st, int n_chains, int* cnt, int* chain_of_slot) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    if (!st) { chain_of_slot[c] = c; if (c == 0) *cnt = n_chains; return; }
    if (!b2_needs_grad(st[c].phase)) return;
    chain_of_slot[atomicAdd(cnt, 1)] = c;
}

// GEMM1 of one tile for a compile-time number of live K steps and split passes, fully unrolled, one uniform add per
// MMA (descriptor low word = per-tile base + constant).  The tensor core adds into its fp32 accumulator with
// truncation (each add loses up to one ulp of the RUNNING sum, towards zero -- measured: a relative bias of
// ~6e-7 on eta at 21 adds, 0.03 nats on a logp of -7e4), so the small cross terms go first, while the
// accumulator is still ~2^-8 of its final size, and only the Qhi.Xhi adds run at full magnitude.
// NPASS = 3: Qlo.Xhi, Qhi.Xlo, Qhi.Xhi;  NPASS = 2 (X is bf16-representable, Xlo == 0): Qlo.Xhi, Qhi.Xhi.
template <int KS, int NPASS>
__device__ __forceinline__ void tw_issue_gemm1(uint32_t d, uint32_t tmem, uint32_t dlo, uint32_t dhi, uint32_t idesc) {
#pragma unroll
    for (int pi = 0; pi < NPASS; ++pi) {
        const bool q_lo = (pi == 0);                           // A operand: Q low part in the first pass only
        const bool x_lo = (NPASS == 3 && pi == 1);             // B operand: X low part in the middle pass of three
        const uint32_t qa = tmem + TW_COL_Q + (q_lo ? 128 : 0);
#pragma unroll
        for (int j = 0; j < KS; ++j)
            mma_ts_split(d, qa + j * 8, dlo + (((x_lo ? TW_XPART_BYTES : 0) + (j >> 2) * (TW_OBS * 128) + (j & 3) * 32) >> 4),
                         dhi, idesc, (pi > 0 || j > 0) ? 1u : 0u);
    }
}

template <int NPASS>
__device__ __forceinline__ void tw_issue_gemm1_ks(int ks, uint32_t d, uint32_t tmem, uint32_t dlo, uint32_t dhi, uint32_t idesc) {
    switch ((ks + 1) >> 1) {                             // 8..16 live K steps, rounded up to even (zero columns)
        case 8: tw_issue_gemm1<16, NPASS>(d, tmem, dlo, dhi, idesc); break;
        case 7: tw_issue_gemm1<14, NPASS>(d, tmem, dlo, dhi, idesc); break;
        case 6: tw_issue_gemm1<12, NPASS>(d, tmem, dlo, dhi, idesc); break;
        case 5: tw_issue_gemm1<10, NPASS>(d, tmem, dlo, dhi, idesc); break;
        default: tw_issue_gemm1<8, NPASS>(d, tmem, dlo, dhi, idesc); break;
    }
}

__global__ void __launch_bounds__(TW_THREADS, 1) k_glm_tcw_main(TwWorkspace ws) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* x_s = smem;
    unsigned char* y_s = x_s + TW_STAGES * TW_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(y_s + TW_STAGES * TW_YS_BYTES);
    uint64_t* q_full = bars;                       // 1
    uint64_t* x_full = bars + 1;                   // TW_STAGES
    uint64_t* x_empty = x_full + TW_STAGES;        // TW_STAGES
    // S and R are double-buffered (b = t & 1) but their barriers are per tile-mod-4 (= per epilogue group): a
    // barrier must only ever be waited on by one agent walking its phases in order -- a group that starts on a
    // fresh barrier with parity 1 falls straight through (the phase "before" phase 0 counts as complete).
    uint64_t* s_full = x_empty + TW_STAGES;        // 4
    uint64_t* s_empty = s_full + 4;                // 4
    uint64_t* p_full = s_empty + 4;                // 4
    uint64_t* p_empty = p_full + 4;                // 4
    uint64_t* g_full = p_empty + 4;                // 1: completes once per flush chunk (GEMM2 of the chunk's last tile)
    uint64_t* g_empty = g_full + 1;                // 1: every epilogue warp has drained the chunk's G
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TwGeom gm = tw_geom(ws);
    const int pair = blockIdx.x >> 1, half = blockIdx.x & 1;
    if (gm.nt == 0 || pair >= gm.nt * gm.sp) return;
    const int ctile = pair / gm.sp, split = pair % gm.sp;
    const int t_begin = split * gm.tps;
    const int t_end = min(ws.n_tiles, t_begin + gm.tps);
    const int T = t_end - t_begin;                 // >= 1

    if (warp == 1 && lane == 0) {
        mbar_init(q_full, TW_EPI_WARPS);
        for (int i = 0; i < TW_STAGES; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, 4); mbar_init(p_full + i, 4); mbar_init(p_empty + i, 1); }
        mbar_init(g_full, 1);
        mbar_init(g_empty, TW_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TW_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===== producer: the X tile ring (one 32 KB + one 128 B bulk copy per stage)
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                const int s = t % TW_STAGES;
                if (t >= TW_STAGES) mbar_wait(x_empty + s, ((t / TW_STAGES) - 1) & 1, ws.err, 1);
                const unsigned char* src = ws.xt + (size_t)(t_begin + t) * TW_STAGE_DATA;
                // Xlo == 0 (two split passes): only the hi half of the tile is read, which halves the HBM / L2
                // traffic of a launch -- the bound when few chains are live (C5: 6.4 -> 3.2 GB per leapfrog)
                const uint32_t x_bytes = ws.n_pass == 2 ? TW_XPART_BYTES : TW_STAGE_BYTES;
                mbar_expect_tx(x_full + s, x_bytes + 2 * TW_Y_BYTES);
                bulk_g2s(x_s + s * TW_STAGE_BYTES, src, x_bytes, x_full + s);
                bulk_g2s(y_s + s * TW_YS_BYTES, src + TW_STAGE_BYTES, TW_Y_BYTES, x_full + s);
                bulk_g2s(y_s + s * TW_YS_BYTES + TW_Y_BYTES, ws.eta_ref + (size_t)(t_begin + t) * TW_OBS, TW_Y_BYTES, x_full + s);
            }
        }
    } else if (warp == 1) {
        // ===== GEMM1 issuer (whole warp runs the uniform loop, one elected lane issues)
        mbar_wait(q_full, 0, ws.err, 2);
        tc_fence_after();
        const int ks = (ws.K + 15) >> 4;                           // <= 16
        const uint32_t idesc_g1 = TC_IDESC_BASE | ((TW_OBS >> 3) << 17) | ((TW_CHAINS >> 4) << 24);   // N = 32, K-major B
        for (int t = 0; t < T; ++t) {
            const int s = t % TW_STAGES, b = t & 1;
            mbar_wait(x_full + s, (t / TW_STAGES) & 1, ws.err, 3);
            if (t >= 2) mbar_wait(s_empty + ((t - 2) & 3), ((t - 2) >> 2) & 1, ws.err, 4);   // S[b] of tile t-2 is in registers
            tc_fence_after();
            const uint32_t x_addr = smem_u32(x_s + s * TW_STAGE_BYTES);
            const uint32_t d = tmem + TW_COL_S + 32 * b;
            if (elect_one()) {
                // descriptors: low word = per-tile base + compile-time offset, high word constant
                const uint32_t dlo = tc_desc_lo(x_addr, 16), dhi = tc_desc_hi(1024);
                if (ws.n_pass == 2) tw_issue_gemm1_ks<2>(ks, d, tmem, dlo, dhi, idesc_g1);
                else tw_issue_gemm1_ks<3>(ks, d, tmem, dlo, dhi, idesc_g1);
                tc_commit(s_full + (t & 3));
            }
            __syncwarp();
        }
    } else if (warp == 3) {
        // ===== GEMM2 issuer: this CTA's feature half, N = its live features rounded up to 16
        int kh = ws.K - 128 * half;
        kh = kh > 128 ? 128 : (kh < 1 ? 1 : kh);                 // K = 128: the second half is all padding, N = 16 of zeros
        const uint32_t n2 = (uint32_t)((kh + 15) & ~15);
        const uint32_t idesc_g2 = TC_IDESC_BASE | (1u << 16) | ((n2 >> 3) << 17) | ((TW_CHAINS >> 4) << 24);
        // The gradient accumulator is drained every F tiles (see the epilogue): thousands of truncating adds into
        // one fp32 accumulator biased the gradient by 2e-4 at 3 M rows (round 2 parity test at C5's shard size).
        const int F = ws.flush_tiles;
        const int np = ws.n_pass;
        for (int u = 0; u < T; ++u) {
            const int s = u % TW_STAGES, b = u & 1;
            const bool chunk_first = (u % F) == 0, chunk_last = ((u + 1) % F) == 0 || u == T - 1;
            mbar_wait(p_full + (u & 3), (u >> 2) & 1, ws.err, 5);
            if (chunk_first && u > 0) mbar_wait(g_empty, ((u / F) - 1) & 1, ws.err, 10);   // the previous chunk's G has been read out
            tc_fence_after();
            const uint32_t x_addr = smem_u32(x_s + s * TW_STAGE_BYTES) + half * 2 * (TW_OBS * 128);
            const uint32_t p_base = tmem + TW_COL_P + 32 * b;
            const uint32_t d = tmem + TW_COL_G;
            if (elect_one()) {
                // MN-major B: 64-feature atoms LBO = 4096 B apart, 8-row groups SBO = 1024 B apart
                const uint32_t dlo = tc_desc_lo(x_addr, TW_OBS * 128), dhi = tc_desc_hi(1024);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {              // Rhi.Xhi, Rlo.Xhi, Rhi.Xlo (the last only if Xlo != 0)
                    if (pass == 2 && np == 2) break;
                    const uint32_t pa = p_base + (pass == 1 ? 16 : 0);
#pragma unroll
                    for (int j = 0; j < TW_OBS / 16; ++j)
                        mma_ts_split(d, pa + j * 8, dlo + (((pass == 2 ? TW_XPART_BYTES : 0) + j * 2048) >> 4), dhi, idesc_g2,
                                     (!chunk_first || pass > 0 || j > 0) ? 1u : 0u);
                }
                tc_commit(x_empty + s);
                tc_commit(p_empty + (u & 3));
                if (chunk_last) tc_commit(g_full);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===== epilogue warpgroups: thread == chain row; group cg owns every 4th tile (32 observation columns)
        const int wq = warp & 3;                                     // TMEM lane quarter this warp may touch
        const int cg = (warp - 4) >> 2;
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const int slot = ctile * TW_CHAINS + row;
        const bool live = slot < gm.n_act;
        float q0 = 0.f;
        {   // Q (A operand of GEMM1): this chain's row, features 64cg..64cg+63, bf16 hi/lo split
            const int chain = live ? ws.chain_of_slot[slot] : 0;
            int sel = 0;
            if (live && ws.st) sel = ws.st[chain].sel;
            const float* q = (sel ? ws.qB : ws.qA) + (size_t)chain * ws.ld;
            if (live) q0 = q[0] - ws.q_ref[0];
#pragma unroll
            for (int rnd = 0; rnd < 2; ++rnd) {
                uint32_t qh[16], ql[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k = 64 * cg + 32 * rnd + 2 * i;
                    const float a = (live && k < ws.K) ? q[1 + k] - ws.q_ref[1 + k] : 0.f;      // dq = q - q_ref
                    const float b2 = (live && k + 1 < ws.K) ? q[2 + k] - ws.q_ref[2 + k] : 0.f;
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b2);
                    const float2 back = __bfloat1622float2(h2);
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - back.x, b2 - back.y);
                    qh[i] = *reinterpret_cast<const uint32_t*>(&h2);
                    ql[i] = *reinterpret_cast<const uint32_t*>(&l2);
                }
                TC_ST16(tmem + lane_addr + TW_COL_Q + 32 * cg + 16 * rnd, qh);
                TC_ST16(tmem + lane_addr + TW_COL_Q + 128 + 32 * cg + 16 * rnd, ql);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_full);
        }
        float lp_sum = 0.f, lp_comp = 0.f, r_sum = 0.f, r_comp = 0.f;   // Kahan: no fp64 in the tile loop
        // Chunked gradient accumulation: after GEMM2 of a chunk's last tile (g_full) every epilogue warp adds its
        // 32 columns of G into the slab's fp32 partial in global memory (round-to-nearest adds; the same thread
        // owns the same elements every time, so plain load-add-store) and releases the accumulator (g_empty).
        const int F = ws.flush_tiles;
        const int n_chunks = (T + F - 1) / F;
        int k_drain = 0;                                             // next chunk this warp has to drain
        float* gout = ws.gpart + ((size_t)split * gm.stride + slot) * TW_KP + 128 * half + 32 * cg;
        auto drain = [&](int k) {
            mbar_wait(g_full, k & 1, ws.err, 8);
            tc_fence_after();
            uint32_t g32[32];
            TC_LD32(tmem + lane_addr + TW_COL_G + 32 * cg, g32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(g_empty);                     // G is in registers: the next chunk may start
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 a = make_float4(__uint_as_float(g32[4 * i]), __uint_as_float(g32[4 * i + 1]),
                                       __uint_as_float(g32[4 * i + 2]), __uint_as_float(g32[4 * i + 3]));
                if (k > 0) {
                    const float4 o = reinterpret_cast<const float4*>(gout)[i];
                    a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
                }
                reinterpret_cast<float4*>(gout)[i] = a;
            }
        };
        for (int t = cg; t < T; t += TW_EPI_GROUPS) {
            const int s = t % TW_STAGES, b = t & 1;
            const float4* ys4 = reinterpret_cast<const float4*>(y_s + s * TW_YS_BYTES);
            const float4* rs4 = ys4 + TW_OBS / 4;                    // eta_ref of the tile's observations
            mbar_wait(s_full + cg, (t >> 2) & 1, ws.err, 6);
            // y values of this stage: the phase is already complete (GEMM1 of this tile consumed the stage) and the
            // next one cannot complete before this tile's R is handed to GEMM2, so the parity test is unambiguous
            mbar_wait(x_full + s, (t / TW_STAGES) & 1, ws.err, 9);
            tc_fence_after();
            uint32_t v[2][16];
            TC_LD16(tmem + lane_addr + TW_COL_S + 32 * b, v[0]);
            TC_LD16(tmem + lane_addr + TW_COL_S + 32 * b + 16, v[1]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + cg);                // S(t) is in registers
            uint32_t hi[2][8], lo[2][8];
            float lsum = 0.f, rsum = 0.f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float yv[16], rv[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 y4 = ys4[4 * hh + i];
                    yv[4 * i] = y4.x; yv[4 * i + 1] = y4.y; yv[4 * i + 2] = y4.z; yv[4 * i + 3] = y4.w;
                    const float4 r4 = rs4[4 * hh + i];
                    rv[4 * i] = r4.x; rv[4 * i + 1] = r4.y; rv[4 * i + 2] = r4.z; rv[4 * i + 3] = r4.w;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float r2[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float eta = (__uint_as_float(v[hh][2 * i + h]) + q0) + rv[2 * i + h];   // dq . x + dq0 + eta_ref
                        const float yy = yv[2 * i + h];
                        const bool valid = yy >= 0.f;
                        const float e = tc_ex2(-1.4426950408889634f * fabsf(eta));     // exp(-|eta|)
                        const float w1 = 1.f + e;
                        const float inv = tc_rcp(w1);
                        const float sig = eta >= 0.f ? inv : e * inv;
                        if (half == 0) {                             // logp is counted by one CTA of the pair
                            const float term = fmaf(yy, eta, -fmaf(0.6931471805599453f, tc_lg2(w1), fmaxf(eta, 0.f)));
                            lsum += valid ? term : 0.f;
                        }
                        r2[h] = valid ? yy - sig : 0.f;
                        rsum += r2[h];
                    }
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(r2[0], r2[1]);
                    const float2 back = __bfloat1622float2(h2);
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(r2[0] - back.x, r2[1] - back.y);
                    hi[hh][i] = *reinterpret_cast<const uint32_t*>(&h2);
                    lo[hh][i] = *reinterpret_cast<const uint32_t*>(&l2);
                }
            }
            // Chunks that ended before this tile are drained here: after this tile's math (useful work while GEMM2
            // of the chunk's last tile finishes), but BEFORE storing its R -- that store waits for GEMM2 of tile t-2,
            // which in a new chunk waits for all sixteen warps to have drained the old one.
            while (k_drain < n_chunks - 1 && (k_drain + 1) * F - 1 < t) { drain(k_drain); ++k_drain; }
            if (t >= 2) mbar_wait(p_empty + ((t - 2) & 3), ((t - 2) >> 2) & 1, ws.err, 7);   // GEMM2 of tile t-2 has read R[b]
            tc_fence_after();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                TC_ST8(tmem + lane_addr + TW_COL_P + 32 * b + 8 * hh, hi[hh]);
                TC_ST8(tmem + lane_addr + TW_COL_P + 32 * b + 16 + 8 * hh, lo[hh]);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full + cg);
            // this warp's tile closes a chunk: drain it as soon as GEMM2 has consumed the R just stored
            if (k_drain < n_chunks - 1 && (k_drain + 1) * F - 1 == t) { drain(k_drain); ++k_drain; }
            float ky = lsum - lp_comp, kt = lp_sum + ky;
            lp_comp = (kt - lp_sum) - ky; lp_sum = kt;
            ky = rsum - r_comp; kt = r_sum + ky;
            r_comp = (kt - r_sum) - ky; r_sum = kt;
        }
        // what is left: at least the last chunk, G[chain row][this half's 128 features] -> global partials
        while (k_drain < n_chunks) { drain(k_drain); ++k_drain; }
        if (half == 0) {
            const size_t o = ((size_t)split * TW_EPI_GROUPS + cg) * gm.stride + slot;
            ws.lpart[o] = (double)lp_sum - (double)lp_comp;
            ws.rpart[o] = (double)r_sum - (double)r_comp;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TW_TMEM_COLS) : "memory");
    }
}

// one warp per live chain: fixed-order reduction over slabs, prior, intercept gradient
__global__ void k_glm_tcw_finalize(TwWorkspace ws, double prior_tau, float* gA, float* gB, double* logp) {
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const TwGeom gm = tw_geom(ws);
    if (slot >= gm.n_act) return;
    const int chain = ws.chain_of_slot[slot];
    const int sel = ws.st ? ws.st[chain].sel : 0;
    const float* q = (sel ? ws.qB : ws.qA) + (size_t)chain * ws.ld;
    float* g = (sel ? gB : gA) + (size_t)chain * ws.ld;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float4* gp = reinterpret_cast<const float4*>(ws.gpart + (size_t)slot * TW_KP) + 2 * lane;
    const size_t stride4 = (size_t)gm.stride * TW_KP / 4;
    for (int sp = 0; sp < gm.sp; ++sp) {
        const float4 a = __ldcg(gp + (size_t)sp * stride4), b = __ldcg(gp + (size_t)sp * stride4 + 1);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    double lp = 0.0, rs = 0.0, prior = 0.0;
    const int n_lp = gm.sp * TW_EPI_GROUPS;
    for (int i = lane; i < n_lp; i += 32) {
        lp += ws.lpart[(size_t)i * gm.stride + slot];
        rs += ws.rpart[(size_t)i * gm.stride + slot];
    }
    const double prior_const = 0.5 * (log(prior_tau) - B2_LOG_2PI);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = 8 * lane + j;
        if (k < ws.K) {
            const double b = (double)q[1 + k];
            g[1 + k] = (float)((double)acc[j] - prior_tau * b);
            prior += -0.5 * prior_tau * b * b + prior_const;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lp += __shfl_xor_sync(0xffffffffu, lp, o);
        rs += __shfl_xor_sync(0xffffffffu, rs, o);
        prior += __shfl_xor_sync(0xffffffffu, prior, o);
    }
    if (lane == 0) {
        g[0] = (float)rs;                                   // intercept: Flat prior (glm/linear.py:50)
        logp[chain] = lp + prior;
    }
}

// ---------------------------------------------------------------------------------- host
struct TwHostState { TwWorkspace ws; };

bool b2_glm_tcw_supported(const b2_engine* e) {
    return e->md.family == B2_FAMILY_GLM_LOGIT && e->dtype == B2_F32 && e->md.G >= 128 && e->md.G <= TW_KP && e->md.N >= 1;
}

static int tw_setup(b2_engine* e, cudaStream_t stream) {
    TwHostState* hs = new TwHostState();
    memset(hs, 0, sizeof(*hs));
    TwWorkspace& w = hs->ws;
    const int N = e->md.N;
    w.n_tiles = (N + TW_OBS - 1) / TW_OBS;
    w.chain_tiles = (e->C + TW_CHAINS - 1) / TW_CHAINS;
    w.grid_ctas = 2 * (e->sm_count / 2);                     // CTA pairs (feature halves), one CTA per SM
    if (w.grid_ctas < 2 * w.chain_tiles) w.grid_ctas = 2 * w.chain_tiles;
    const size_t slabs = (size_t)w.grid_ctas / 2 + w.chain_tiles;
    B2_CUDA_OK(cudaMalloc(&w.counter, 64 * sizeof(int)));
    B2_CUDA_OK(cudaMalloc(&w.chain_of_slot, (size_t)w.chain_tiles * TW_CHAINS * sizeof(int)));
    B2_CUDA_OK(cudaMalloc(&w.xt, (size_t)w.n_tiles * TW_STAGE_DATA));
    B2_CUDA_OK(cudaMalloc(&w.gpart, slabs * TW_CHAINS * TW_KP * sizeof(float)));
    B2_CUDA_OK(cudaMalloc(&w.lpart, slabs * TW_EPI_GROUPS * TW_CHAINS * sizeof(double)));
    B2_CUDA_OK(cudaMalloc(&w.rpart, slabs * TW_EPI_GROUPS * TW_CHAINS * sizeof(double)));
    B2_CUDA_OK(cudaMalloc(&w.err, TC_ERR_INTS * sizeof(int)));
    B2_CUDA_OK(cudaMemsetAsync(w.err, 0, TC_ERR_INTS * sizeof(int), stream));
    float *q_ref = nullptr, *eta_ref = nullptr;
    B2_CUDA_OK(cudaMalloc(&q_ref, (TW_KP + 64) * sizeof(float)));
    B2_CUDA_OK(cudaMalloc(&eta_ref, (size_t)w.n_tiles * TW_OBS * sizeof(float)));
    B2_CUDA_OK(cudaMemsetAsync(q_ref, 0, (TW_KP + 64) * sizeof(float), stream));
    B2_CUDA_OK(cudaMemsetAsync(eta_ref, 0, (size_t)w.n_tiles * TW_OBS * sizeof(float), stream));
    w.q_ref = q_ref; w.eta_ref = eta_ref;
    B2_CUDA_OK(cudaMalloc(&w.x_flags, 4 * sizeof(int)));
    B2_CUDA_OK(cudaMemsetAsync(w.x_flags, 0, 4 * sizeof(int), stream));
    k_glm_tcw_prep_x<<<w.n_tiles, 256, 0, stream>>>(e->md.X, e->md.yf, N, e->md.G, w.xt, w.x_flags);
    B2_CUDA_OK(cudaGetLastError());
    int has_lo = 1;                                          // one-time: is the Q.Xlo / R.Xlo pass multiplying zeros?
    B2_CUDA_OK(cudaMemcpyAsync(&has_lo, w.x_flags, sizeof(int), cudaMemcpyDeviceToHost, stream));
    B2_CUDA_OK(cudaStreamSynchronize(stream));
    w.n_pass = has_lo ? 3 : 2;
    if (getenv("B2_TCW_PASSES") && atoi(getenv("B2_TCW_PASSES")) == 3) w.n_pass = 3;
    w.flush_tiles = getenv("B2_TCW_FLUSH") ? atoi(getenv("B2_TCW_FLUSH")) : 128;
    if (w.flush_tiles < 8) w.flush_tiles = 8;
    B2_CUDA_OK(cudaFuncSetAttribute(k_glm_tcw_main, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM_BYTES));
    e->launches += 1;
    e->glm_tcw = hs;
    return 0;
}

void b2_glm_tcw_release(b2_engine* e) {
    if (!e->glm_tcw) return;
    TwHostState* hs = (TwHostState*)e->glm_tcw;
    cudaFree(hs->ws.xt); cudaFree(hs->ws.counter); cudaFree(hs->ws.chain_of_slot); cudaFree(hs->ws.gpart);
    cudaFree(hs->ws.lpart); cudaFree(hs->ws.rpart); cudaFree(hs->ws.err); cudaFree(hs->ws.x_flags);
    cudaFree((void*)hs->ws.q_ref); cudaFree((void*)hs->ws.eta_ref);
    delete hs;
    e->glm_tcw = nullptr;
}

// Reference position (b2_glm_tc.cu, TcWorkspace::q_ref): refreshed at the start of every lock-step / stepwise run
// (b2_engine.cu) and for every likelihood-only call -- not per leapfrog: eta_ref = X . q_ref reads the whole fp32 X.
int b2_glm_tcw_refresh(b2_engine* e, const float* qA, const float* qB, int ld, const B2ChainState* st, int n, cudaStream_t stream) {
    if (!e->glm_tcw) { int rc = tw_setup(e, stream); if (rc) return rc; }
    if (getenv("B2_TC_NOREF") && atoi(getenv("B2_TC_NOREF"))) return 0;
    TwWorkspace& w = ((TwHostState*)e->glm_tcw)->ws;
    k_glm_ref_mean<<<TW_KP + 64, 256, 0, stream>>>(qA, qB, ld, st, 0, n, e->md.G + 1, const_cast<float*>(w.q_ref), TW_KP + 64);
    const int n_pad = w.n_tiles * TW_OBS;
    k_glm_ref_eta<<<(n_pad + 7) / 8, 256, 0, stream>>>(e->md.X, e->md.N, e->md.G, w.q_ref, 1, const_cast<float*>(w.eta_ref), n_pad);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 2;
    return 0;
}

int b2_glm_tcw_launch(b2_engine* e, const float* qA, const float* qB, float* gA, float* gB, int ld,
                      const B2ChainState* st, int n, double* logp, cudaStream_t stream) {
    if (!e->glm_tcw) { int rc = tw_setup(e, stream); if (rc) return rc; }
    TwHostState* hs = (TwHostState*)e->glm_tcw;
    TwWorkspace& w = hs->ws;
    w.qA = qA; w.qB = qB; w.ld = ld; w.st = st; w.n_chains = n; w.K = e->md.G;
    if (st == nullptr) {                                     // likelihood-only call: the reference is this batch's mean
        int rc = b2_glm_tcw_refresh(e, qA, qB, ld, st, n, stream);
        if (rc) return rc;
    }
    B2_CUDA_OK(cudaMemsetAsync(w.counter, 0, sizeof(int), stream));
    k_glm_tcw_compact<<<(n + 255) / 256, 256, 0, stream>>>(st, n, w.counter, w.chain_of_slot);
    B2_CUDA_OK(cudaGetLastError());
    k_glm_tcw_main<<<w.grid_ctas, TW_THREADS, TW_SMEM_BYTES, stream>>>(w);
    B2_CUDA_OK(cudaGetLastError());
    k_glm_tcw_finalize<<<(n + 3) / 4, 128, 0, stream>>>(w, e->md.hp[0], gA, gB, logp);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 3;
    return 0;
}

// debugging aid (B2_TC_NOTRAP builds): first stuck mbarrier wait = tag + 100 * warp + 10000 * CTA, 0 if none
extern "C" int b2_debug_tcw_err(b2_engine* e, int* host_out) {
    if (!e || !e->glm_tcw) return -1;
    B2_CUDA_OK(cudaDeviceSynchronize());
    B2_CUDA_OK(cudaMemcpy(host_out, ((TwHostState*)e->glm_tcw)->ws.err, TC_ERR_INTS * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}
