// Chain-batched Bernoulli-logit GLM likelihood on the 5th-generation tensor cores (sm_100a).
//
//   eta[c, i] = sum_k q[c, k] Xa[i, k]                 (Xa = [1 | X], K1 = K + 1 <= 128 columns)
//   logp_c    = sum_i y_i eta - softplus(eta)          (+ prior, added in the finalize kernel)
//   grad_c[k] = sum_i (y_i - sigmoid(eta[c, i])) Xa[i, k]
// (reference: pymc3/glm/linear.py:49-101, glm/families.py:115-119, discrete.py:104,350 evaluated by
//  the Theano-compiled ValueGradFunction.__call__, model.py:645-666 -- two GEMVs per chain there.)
//
// One CTA owns 128 chains (the MMA M dimension) and a contiguous range of 64-observation tiles.
// Per tile, flash-attention style, with nothing but the X tile crossing HBM/L2:
//   GEMM1  S[128 chains, 64 obs]   = Q . Xtile^T     tcgen05.mma TS, A = Q (TMEM, written once per CTA), B = Xtile (smem, K-major)
//   epi    R = y - sigmoid(S), logp += ...           tcgen05.ld -> registers -> bf16 hi/lo -> tcgen05.st (TMEM)
//   GEMM2  G[128 chains, 128 feat] += R . Xtile      tcgen05.mma TS, A = R (TMEM), B = the SAME smem tile, MN-major
// fp32-level accuracy from bf16 tensor cores by two-term splits (x = hi + lo, both bf16):
//   S = Qhi.Xhi + Qlo.Xhi + Qhi.Xlo ,   G = Rhi.Xhi + Rlo.Xhi + Rhi.Xlo      (error ~2^-17 per product)
// X tiles are staged by cp.async.bulk (TMA engine, 1-D bulk copies): X is pre-tiled ONCE at model
// build into the exact 128B-swizzled shared-memory image the UMMA descriptors expect, so a pipeline
// stage is one contiguous 32 KB copy (+256 B of y) and no tensor map is needed.  Q (the positions) is
// split to bf16 hi/lo by the epilogue warps and written straight into TMEM.  Accumulators live in TMEM
// (S double-buffered 2x64 cols, R 2x64 cols, G 128 cols, Q 128 cols).  Warp roles: 0 = bulk-copy producer,
// 1 = GEMM1 issuer, 3 = GEMM2 issuer (the whole warp runs the uniform loop, one elected lane issues),
// 2 = TMEM allocator, 4..19 = four epilogue warpgroups (TMEM lane == chain).  Warpgroups {0,1} take even tiles and
// {2,3} odd tiles, 32 observation columns per warp and tile, so each SM sub-partition always has four epilogue
// warps to interleave and the two pairs run one tile apart (one warp per scheduler was latency-bound, ncu r1;
// all sixteen marching through the same tile left the MUFU and TMEM phases unoverlapped, ncu r1b).
// Only the K steps / N columns that hold real features are issued (7 of 8 K steps, N = 112 at D + 1 = 101).
#include <cuda_bf16.h>
#include <cstring>
#include <cstdlib>
#include "b2_engine.cuh"
#include "b2_tc_ptx.cuh"

#define TC_CHAINS 128                 // MMA M
#define TC_OBS 64                     // observations per tile (GEMM1 N, GEMM2 K)
#define TC_POST_STAGE 6               // stack buffers k_glm_tc_post stages per chain (merge levels 0..5)
#define TC_KP 128                     // padded feature count (GEMM1 K, GEMM2 N)
#define TC_STAGES 6
#define TC_XPART_BYTES (TC_OBS * TC_KP * 2)            // 16384: one of {hi, lo}, two 64-column atoms
#define TC_Y_BYTES (TC_OBS * 4)                        // 256
#define TC_STAGE_DATA (2 * TC_XPART_BYTES + TC_Y_BYTES) // 33024 in global memory: Xhi | Xlo | y
#define TC_STAGE_BYTES (2 * TC_XPART_BYTES)            // 32768 in shared memory (y lives in its own ring)
#define TC_SMEM_BYTES (1024 + TC_STAGES * (TC_STAGE_BYTES + TC_Y_BYTES) + 256)
#define TC_EPI_GROUPS 4               // epilogue warpgroups; group g owns observation columns 16g..16g+15 of a tile
#define TC_EPI_WARPS (4 * TC_EPI_GROUPS)
#define TC_THREADS (128 + 32 * TC_EPI_WARPS)
#define TC_TMEM_COLS 512
#define TC_COL_S 0                    // S[b] at 64 b
#define TC_COL_P 128                  // P[b] at 128 + 64 b   (hi: 32 cols, lo: 32 cols; 2 bf16 per column)
#define TC_COL_G 256                  // 128 columns
#define TC_COL_Q 384                  // Q as the A operand of GEMM1: hi 64 cols | lo 64 cols (2 bf16 per column)

struct TcWorkspace {
    unsigned char* xt;       // [n_tiles][TC_STAGE_DATA] pre-swizzled X tiles (+ y)
    float* gpart;            // [splits][c_pad][TC_KP]
    double* lpart;           // [splits][TC_EPI_GROUPS][c_pad]
    int n_tiles, c_pad, chain_tiles, splits, tiles_per_split, n_pad_rows;
    int* err;                // device watchdog flag
    // per launch: where every chain's pending position lives
    const float* qA; const float* qB; int ld; const B2ChainState* st; int n_chains; int K1;
    // active-chain compaction: chains that still need a gradient are packed into dense 128-row tiles, so a
    // step costs ceil(n_active/128) chain tiles instead of all of them (early tuning and the tail of a run
    // chunk leave most chains idle); the grid is fixed and re-divides its CTAs over the live tiles.
    int* counters;           // [2] number of active chains for step parity p / p^1
    int* chain_of_slot;      // [c_pad]
    int* slot_of_chain;      // [c_pad]
    int parity;              // which counter this launch reads
    int grid_ctas;           // CTAs of the main grid
    long long* dbg;          // optional timeline of CTA (0,0): [event][tile] clock64 stamps (B2_TC_TIMELINE=1)
};
#define TC_DBG_TILES 256
#define TC_STAMP(ev, t) do { if (ws.dbg && blockIdx.x == 0 && (t) < TC_DBG_TILES) ws.dbg[(ev) * TC_DBG_TILES + (t)] = clock64(); } while (0)

// ------------------------------------------------------------------------ one-time X tiling
__global__ void k_glm_tc_prep_x(const float* __restrict__ X, const float* __restrict__ y, int N, int K,
                                unsigned char* __restrict__ xt, int n_tiles) {
    const int tile = blockIdx.x;
    unsigned char* blob = xt + (size_t)tile * TC_STAGE_DATA;
    for (int idx = threadIdx.x; idx < TC_OBS * TC_KP; idx += blockDim.x) {
        const int r = idx / TC_KP, c = idx - r * TC_KP;
        const int row = tile * TC_OBS + r;
        float v = 0.f;
        if (row < N) {
            if (c == 0) v = 1.f;                                   // intercept column
            else if (c <= K) v = X[(size_t)row * K + (c - 1)];
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int off = (c >> 6) * (TC_OBS * 128) + tc_swz(r, c & 63);
        *reinterpret_cast<__nv_bfloat16*>(blob + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(blob + TC_XPART_BYTES + off) = lo;
    }
    for (int r = threadIdx.x; r < TC_OBS; r += blockDim.x) {
        const int row = tile * TC_OBS + r;
        reinterpret_cast<float*>(blob + 2 * TC_XPART_BYTES)[r] = row < N ? y[row] : 0.f;
    }
}

// instruction descriptors (cute::UMMA::InstrDescriptor): D=F32, A=B=BF16
#define TC_IDESC_G1 (TC_IDESC_BASE | ((TC_OBS >> 3) << 17) | ((TC_CHAINS >> 4) << 24))               // N=64,  K-major B
// GEMM2 (MN-major B, bit 16) takes its N from the feature count at run time: see the GEMM2 issuer

struct TcGeom { int n_act, nt, sp, tps, stride; };     // active chains, live tiles, row slabs, tiles per slab, slots
__device__ __forceinline__ TcGeom tc_geom(const TcWorkspace& ws) {
    TcGeom g;
    g.n_act = ws.counters[ws.parity];
    g.nt = (g.n_act + TC_CHAINS - 1) / TC_CHAINS;
    g.sp = g.nt > 0 ? ws.grid_ctas / g.nt : 1;
    if (g.sp > ws.n_tiles) g.sp = ws.n_tiles;
    g.tps = (ws.n_tiles + g.sp - 1) / g.sp;
    g.sp = (ws.n_tiles + g.tps - 1) / g.tps;            // slabs that actually own rows
    g.stride = g.nt * TC_CHAINS;
    return g;
}

// (re)builds the slot <-> chain maps from the chain states (first step of a run / parity hook)
__global__ void k_glm_tc_compact(TcWorkspace ws, const B2ChainState* st, int n_chains) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    const bool live = st ? b2_needs_grad(st[c].phase) : true;
    if (!live) return;
    const int slot = st ? atomicAdd(ws.counters + ws.parity, 1) : c;
    ws.chain_of_slot[slot] = c;
    ws.slot_of_chain[c] = slot;
    if (!st && c == 0) ws.counters[ws.parity] = n_chains;
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_glm_tc_main(TcWorkspace ws) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* x_s = smem;
    unsigned char* y_s = x_s + TC_STAGES * TC_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(y_s + TC_STAGES * TC_Y_BYTES);
    uint64_t* q_full = bars;                       // 1
    uint64_t* x_full = bars + 1;                   // TC_STAGES
    uint64_t* x_empty = x_full + TC_STAGES;        // TC_STAGES
    uint64_t* s_full = x_empty + TC_STAGES;        // 2
    uint64_t* s_empty = s_full + 2;                // 2
    uint64_t* p_full = s_empty + 2;                // 2
    uint64_t* p_empty = p_full + 2;                // 2
    uint64_t* g_full = p_empty + 2;                // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TcGeom gm = tc_geom(ws);
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.counters[ws.parity ^ 1] = 0;   // k_glm_tc_post of this step refills it
    const int ctile = blockIdx.x / gm.sp, split = blockIdx.x % gm.sp;
    if (gm.nt == 0 || ctile >= gm.nt) return;      // whole CTA: no live chain tile for it
    const int t_begin = split * gm.tps;
    const int t_end = min(ws.n_tiles, t_begin + gm.tps);
    const int T = t_end - t_begin;                 // >= 1: gm.sp counts only slabs that own rows

    if (warp == 1 && lane == 0) {
        mbar_init(q_full, TC_EPI_WARPS);
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, TC_EPI_WARPS / 2); mbar_init(p_full + i, TC_EPI_WARPS / 2); mbar_init(p_empty + i, 1); }
        mbar_init(g_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===== producer: Q tile once, then the X tile ring =====
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                const int s = t % TC_STAGES;
                if (t >= TC_STAGES) mbar_wait(x_empty + s, ((t / TC_STAGES) - 1) & 1, ws.err, 1);
                const unsigned char* src = ws.xt + (size_t)(t_begin + t) * TC_STAGE_DATA;
                mbar_expect_tx(x_full + s, TC_STAGE_DATA);
                bulk_g2s(x_s + s * TC_STAGE_BYTES, src, TC_STAGE_BYTES, x_full + s);
                bulk_g2s(y_s + s * TC_Y_BYTES, src + TC_STAGE_BYTES, TC_Y_BYTES, x_full + s);
            }
        }
    } else if (warp == 1) {
        // ===== GEMM1 issuer.  The whole warp runs the (warp-uniform) loop so descriptors and TMEM
        // addresses stay in uniform registers; only tcgen05.mma / tcgen05.commit are issued by one
        // elected lane.  (Under `if (lane == 0)` ptxas wrapped every UTCHMMA in an R2UR + per-thread
        // election loop: ~75 cycles of issue per MMA against 32-64 cycles of tensor work.)
        // GEMM1 and GEMM2 have their own issuing warps (1 and 3): tcgen05.mma issue blocks on a shallow
        // queue, so with a single issuer every barrier poll between the two GEMMs was tensor idle time
        // (timeline r1: 2460 cycles of issuer time per tile for 1536 cycles of tensor work).
        mbar_wait(q_full, 0, ws.err, 2);                           // epilogue warps have written Q into TMEM
        tc_fence_after();
        const int ks = (ws.K1 + 15) >> 4;                          // K steps that hold real features (7 of 8 at D+1 = 101)
        for (int t = 0; t < T; ++t) {
            const int s = t % TC_STAGES, b = t & 1;
            if (lane == 0) TC_STAMP(0, t);
            mbar_wait(x_full + s, (t / TC_STAGES) & 1, ws.err, 3);
            if (t >= 2) mbar_wait(s_empty + b, ((t >> 1) - 1) & 1, ws.err, 4);
            tc_fence_after();
            const uint32_t x_addr = smem_u32(x_s + s * TC_STAGE_BYTES);
            const uint32_t d = tmem + TC_COL_S + 64 * b;
            if (elect_one()) {
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {              // Qhi.Xhi, Qlo.Xhi, Qhi.Xlo
                    const uint32_t qa = tmem + TC_COL_Q + (pass == 1 ? 64 : 0);
                    const uint32_t xa = x_addr + (pass == 2 ? TC_XPART_BYTES : 0);
#pragma unroll
                    for (int j = 0; j < TC_KP / 16; ++j) {
                        if (j >= ks) break;                            // all-zero padding columns: no tensor work spent on them
                        const uint32_t koff = (j & 3) * 32;
                        const uint64_t bd = make_desc(xa + (j >> 2) * (TC_OBS * 128) + koff, 16, 1024);
                        mma_ts(d, qa + j * 8, bd, TC_IDESC_G1, acc);   // A from TMEM: no 4 KB smem read per MMA
                        acc = 1;
                    }
                }
                tc_commit(s_full + b);
                TC_STAMP(1, t);
            }
            __syncwarp();
        }
    } else if (warp == 3) {
        // ===== GEMM2 issuer: G += R(u) . Xtile(u), three split passes; releases the X stage and the R buffer
        // N = features rounded up to 16 (UMMA N granularity at M = 128): 112 instead of 128 at D+1 = 101
        const uint32_t n2 = (uint32_t)((ws.K1 + 15) & ~15);
        const uint32_t idesc_g2 = TC_IDESC_BASE | (1u << 16) | ((n2 >> 3) << 17) | ((TC_CHAINS >> 4) << 24);
        for (int u = 0; u < T; ++u) {
            const int s = u % TC_STAGES, b = u & 1;
            if (lane == 0) TC_STAMP(2, u);
            mbar_wait(p_full + b, (u >> 1) & 1, ws.err, 5);
            tc_fence_after();
            const uint32_t x_addr = smem_u32(x_s + s * TC_STAGE_BYTES);
            const uint32_t p_base = tmem + TC_COL_P + 64 * b;
            const uint32_t d = tmem + TC_COL_G;
            if (elect_one()) {
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {              // Rhi.Xhi, Rlo.Xhi, Rhi.Xlo
                    const uint32_t pa = p_base + (pass == 1 ? 32 : 0);
                    const uint32_t xa = x_addr + (pass == 2 ? TC_XPART_BYTES : 0);
#pragma unroll
                    for (int j = 0; j < TC_OBS / 16; ++j) {
                        // MN-major B: 2 feature atoms LBO = 8192 B apart, 8-row groups SBO = 1024 B apart
                        const uint64_t bd = make_desc(xa + j * 2048, TC_OBS * 128, 1024);
                        mma_ts(d, pa + j * 8, bd, idesc_g2, (u > 0 || pass > 0 || j > 0) ? 1u : 0u);
                    }
                }
                tc_commit(x_empty + s);
                tc_commit(p_empty + b);
                if (u == T - 1) tc_commit(g_full);
                TC_STAMP(3, u);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===== epilogue warpgroups: thread == (chain row, 16-observation column group) =====
        const int wq = warp & 3;                                     // TMEM lane quarter this warp may touch
        const int cg = (warp - 4) >> 2;                              // column group 0..3
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        // Column groups {0,1} take even tiles, {2,3} odd tiles (32 observation columns per warp and
        // tile): the two pairs run one tile apart, so MUFU-heavy math of one pair overlaps the TMEM
        // loads / stores / barrier waits of the other instead of all 16 warps marching in lock-step.
        {   // Q (A operand of GEMM1): this thread's chain row, features 32cg..32cg+31, bf16 hi/lo split
            const int slot = ctile * TC_CHAINS + row;
            const bool live = slot < gm.n_act;
            const int chain = live ? ws.chain_of_slot[slot] : 0;
            int sel = 0;
            if (live && ws.st) sel = ws.st[chain].sel;
            const float* q = (sel ? ws.qB : ws.qA) + (size_t)(live ? chain : 0) * ws.ld;
            uint32_t qh[16], ql[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = 32 * cg + 2 * i;
                const float a = (live && k < ws.K1) ? q[k] : 0.f;
                const float b2 = (live && k + 1 < ws.K1) ? q[k + 1] : 0.f;
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b2);
                const float2 back = __bfloat1622float2(h2);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - back.x, b2 - back.y);
                qh[i] = *reinterpret_cast<const uint32_t*>(&h2);
                ql[i] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            TC_ST16(tmem + lane_addr + TC_COL_Q + 16 * cg, qh);
            TC_ST16(tmem + lane_addr + TC_COL_Q + 64 + 16 * cg, ql);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_full);
        }
        const int pair = cg >> 1, sub = cg & 1;
        float lp_sum = 0.f, lp_comp = 0.f;                         // Kahan: no FP64 adds in the tile loop (ncu r1b)
        for (int t = pair; t < T; t += 2) {
            const int s = t % TC_STAGES, b = pair;
            const float4* ys4 = reinterpret_cast<const float4*>(y_s + s * TC_Y_BYTES) + 8 * sub;
            const bool stamp = (lane == 0) && (sub == 0) && (wq == 0);
            if (stamp) TC_STAMP(4, t);
            mbar_wait(x_full + s, (t / TC_STAGES) & 1, ws.err, 9);    // y values of this stage (async-proxy writes)
            mbar_wait(s_full + b, (t >> 1) & 1, ws.err, 6);
            if (stamp) TC_STAMP(5, t);
            tc_fence_after();
            uint32_t v[2][16];
            TC_LD16(tmem + lane_addr + TC_COL_S + 64 * b + 32 * sub, v[0]);
            TC_LD16(tmem + lane_addr + TC_COL_S + 64 * b + 32 * sub + 16, v[1]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + b);                 // S(t) is in registers: GEMM1(t+2) may overwrite it
            if (stamp) TC_STAMP(6, t);
            if (lane == 0) TC_STAMP(9 + 2 * (warp - 4), t);
            uint32_t hi[2][8], lo[2][8];
            float lsum = 0.f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float yv[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 y4 = ys4[4 * hh + i];
                    yv[4 * i] = y4.x; yv[4 * i + 1] = y4.y; yv[4 * i + 2] = y4.z; yv[4 * i + 3] = y4.w;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float r2[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float eta = __uint_as_float(v[hh][2 * i + h]);
                        const float yy = yv[2 * i + h];
                        const float e = tc_ex2(-1.4426950408889634f * fabsf(eta));     // exp(-|eta|)
                        const float w1 = 1.f + e;
                        const float inv = tc_rcp(w1);
                        const float sig = eta >= 0.f ? inv : e * inv;
                        // y*eta - softplus(eta),  softplus = max(eta,0) + log(1 + exp(-|eta|))
                        lsum += fmaf(yy, eta, -fmaf(0.6931471805599453f, tc_lg2(w1), fmaxf(eta, 0.f)));
                        r2[h] = yy - sig;
                    }
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(r2[0], r2[1]);
                    const float2 back = __bfloat1622float2(h2);
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(r2[0] - back.x, r2[1] - back.y);
                    hi[hh][i] = *reinterpret_cast<const uint32_t*>(&h2);
                    lo[hh][i] = *reinterpret_cast<const uint32_t*>(&l2);
                }
            }
            if (stamp) TC_STAMP(7, t);
            if (t >= 2) mbar_wait(p_empty + b, ((t >> 1) - 1) & 1, ws.err, 7);
            tc_fence_after();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                TC_ST8(tmem + lane_addr + TC_COL_P + 64 * b + 16 * sub + 8 * hh, hi[hh]);
                TC_ST8(tmem + lane_addr + TC_COL_P + 64 * b + 32 + 16 * sub + 8 * hh, lo[hh]);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full + b);
            if (stamp) TC_STAMP(8, t);
            if (lane == 0) TC_STAMP(10 + 2 * (warp - 4), t);
            const float ky = lsum - lp_comp;
            const float kt = lp_sum + ky;
            lp_comp = (kt - lp_sum) - ky;
            lp_sum = kt;
        }
        const double logp = (double)lp_sum - (double)lp_comp;
        // the slab's gradient tile: G[chain row][128 features] -> global partials (32 columns per group)
        mbar_wait(g_full, 0, ws.err, 8);
        tc_fence_after();
        const int slot = ctile * TC_CHAINS + row;
        float* gout = ws.gpart + ((size_t)split * gm.stride + slot) * TC_KP + 32 * cg;
        {
            uint32_t g32[32];
            TC_LD32(tmem + lane_addr + TC_COL_G + 32 * cg, g32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i)
                reinterpret_cast<float4*>(gout)[i] =
                    make_float4(__uint_as_float(g32[4 * i]), __uint_as_float(g32[4 * i + 1]),
                                __uint_as_float(g32[4 * i + 2]), __uint_as_float(g32[4 * i + 3]));
        }
        // logp partial of this column group; the finalize kernel adds the TC_EPI_GROUPS partials
        ws.lpart[((size_t)split * TC_EPI_GROUPS + cg) * gm.stride + slot] = logp;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
    }
}

// fixed-order reduction over slabs + prior + correction for zero-padded rows, for one chain per warp.
// lane l owns features 4l..4l+3 (one float4 per slab partial).
__device__ __forceinline__ double tc_finalize_chain(const TcWorkspace& ws, const TcGeom& gm, int slot, int lane, int K1,
                                                    double prior_tau, const float* q, float* g) {
    // logp slab partials first, so their round trip overlaps the gradient partials'
    const int n_lp = gm.sp * TC_EPI_GROUPS;
    double lpv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        lpv[j] = (lane + 32 * j < n_lp) ? __ldcg(ws.lpart + (size_t)(lane + 32 * j) * gm.stride + slot) : 0.0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* gp = reinterpret_cast<const float4*>(ws.gpart + (size_t)slot * TC_KP) + lane;
    const size_t stride4 = (size_t)gm.stride * TC_KP / 4;
    // all slab partials of this lane in flight at once (one L2 round trip), then a fixed-order sum
    for (int sp0 = 0; sp0 < gm.sp; sp0 += 24) {
        float4 v[24];
#pragma unroll
        for (int j = 0; j < 24; ++j)
            v[j] = (sp0 + j < gm.sp) ? __ldcg(gp + (size_t)(sp0 + j) * stride4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 24; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
    double prior = 0.0;
    const double prior_const = 0.5 * (log(prior_tau) - B2_LOG_2PI);    // once per warp, not per coefficient
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = 4 * lane + j;
        if (k < K1) {
            double sgrad = (double)a4[j];
            if (k > 0) {
                const double b = (double)q[k];
                sgrad -= prior_tau * b;
                prior += -0.5 * prior_tau * b * b + prior_const;
            }
            g[k] = (float)sgrad;
        }
    }
    double lp = (lpv[0] + lpv[1]) + (lpv[2] + lpv[3]);
    for (int sp = lane + 128; sp < n_lp; sp += 32) lp += ws.lpart[(size_t)sp * gm.stride + slot];
    for (int o = 16; o > 0; o >>= 1) {
        prior += __shfl_xor_sync(0xffffffffu, prior, o);
        lp += __shfl_xor_sync(0xffffffffu, lp, o);
    }
    // zero-padded rows have eta = 0 exactly and contributed -log 2 each
    return lp + prior + (double)ws.n_pad_rows * B2_LOG_2;
}

__global__ void k_glm_tc_finalize(TcWorkspace ws, int n_chains, int K1, double prior_tau, const float* qA,
                                  const float* qB, float* gA, float* gB, int ld, const B2ChainState* st, double* logp) {
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (chain >= n_chains) return;
    int sel = 0;
    if (st) {
        if (st[chain].phase > B2_PHASE_HMC) return;
        sel = st[chain].sel;
    }
    const float* q = (sel ? qB : qA) + (size_t)chain * ld;
    float* g = (sel ? gB : gA) + (size_t)chain * ld;
    const TcGeom gm = tc_geom(ws);
    const double lp = tc_finalize_chain(ws, gm, ws.slot_of_chain[chain], lane, K1, prior_tau, q, g);
    if (lane == 0) logp[chain] = lp;
}

// Lock-step companion of k_glm_tc_main: one warp per chain does
//   slab reduction (-> logp, grad)  ->  b2_advance (the whole NUTS/HMC bookkeeping for this leapfrog)
//   ->  bf16 hi/lo re-split of the next pending position into the swizzled Q tile.
// Replaces three launches (finalize, advance, pack) and the round trip of the gradient through HBM.
__global__ void k_glm_tc_post(TcWorkspace ws, B2View<float> w, int K1, double prior_tau) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= w.C) return;
    B2WarpGroup g;
    w.dbg = ws.dbg ? ws.dbg + 48 * TC_DBG_TILES : nullptr;          // post-kernel stamps live behind the main kernel's
    const long long t_start = clock64();
    __shared__ __align__(16) float hot_s[4][B2_V_STACK0 * 128];
    __shared__ double lv_s[4][4 * B2_MAX_LEVELS];
    float* hot = hot_s[threadIdx.x >> 5];
    double* lvh = lv_s[threadIdx.x >> 5];
    // Everything this warp needs from global memory is requested before anything is stored to shared memory
    // (the loads go through generic pointers, so the compiler keeps them behind earlier shared stores):
    // one round trip for the chain state, its 11 hot vector slots, the level scalars and the slot maps.
    const int lane0 = threadIdx.x & 31;
    float4 tmp[B2_V_STACK0];
#pragma unroll
    for (int slot = 0; slot < B2_V_STACK0; ++slot)
        tmp[slot] = (4 * lane0 < w.Dp) ? *reinterpret_cast<const float4*>(w.Vglobal(slot, c) + 4 * lane0) : make_float4(0.f, 0.f, 0.f, 0.f);
    const double* lv_src = w.lv + (size_t)c * 4 * B2_MAX_LEVELS;
    const double lv_a0 = lv_src[lane0];
    const double lv_a1 = (lane0 + 32 < 4 * B2_MAX_LEVELS) ? lv_src[lane0 + 32] : 0.0;
    const int my_slot = ws.slot_of_chain[c];
    const TcGeom gm = tc_geom(ws);
    B2ChainState s = w.st[c];
#pragma unroll
    for (int slot = 0; slot < B2_V_STACK0; ++slot)
        if (4 * lane0 < w.Dp) *reinterpret_cast<float4*>(hot + slot * w.Dp + 4 * lane0) = tmp[slot];
    lvh[lane0] = lv_a0;
    if (lane0 + 32 < 4 * B2_MAX_LEVELS) lvh[lane0 + 32] = lv_a1;
    if (!b2_needs_grad(s.phase)) return;
    // Stack buffers the pending leaf will merge (one per trailing one-bit of its index, nuts.py:347-389 as a
    // binary counter): fetched with cp.async while the slab reduction below runs, so every merge level works
    // out of shared memory instead of paying 4-6 dependent L2 round trips (timeline r1: 5-7k cycles per level,
    // and the kernel lasts as long as its deepest merge chain).
    extern __shared__ __align__(16) unsigned char post_dyn[];
    float* stk = reinterpret_cast<float*>(post_dyn) + (size_t)(threadIdx.x >> 5) * TC_POST_STAGE * B2_S_NVEC * w.Dp;
    int n_merge = 0, n_staged = 0, wb_buf = -1;
    unsigned stk_mask = 0;
    unsigned long long stk_idx = 0;
    if (s.phase == B2_PHASE_TREE) {
        while ((s.leaf_n >> n_merge) & 1) ++n_merge;
        n_staged = n_merge < TC_POST_STAGE ? n_merge : TC_POST_STAGE;
        const int lane = threadIdx.x & 31;
        for (int k = 0; k < n_staged; ++k) {
            const int buf = b2_map_get(s.slot_map, k);
            stk_mask |= 1u << buf;
            stk_idx |= (unsigned long long)k << (4 * buf);
            if (4 * lane < w.Dp) {
#pragma unroll
                for (int which = 0; which < B2_S_NVEC; ++which) {
                    const float* src = w.Vglobal(B2_V_STACK0 + buf * B2_S_NVEC + which, c) + 4 * lane;
                    const uint32_t dst = smem_u32(stk + ((size_t)k * B2_S_NVEC + which) * w.Dp + 4 * lane);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                }
            }
        }
        if (n_staged == n_merge && n_merge > 0) wb_buf = b2_map_get(s.slot_map, n_merge - 1);   // the merged sub-tree ends up here
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    __syncwarp();
    w.hot = hot;
    w.lv_hot = lvh;
    if (w.dbg && c == 0) { w.dbg[(s.n_grad & 4095) * 16 + 0] = t_start; w.dbg[(s.n_grad & 4095) * 16 + 10] = n_merge; w.dbg[(s.n_grad & 4095) * 16 + 11] = s.iter; }
    B2_STAMP(w, c, s, 1);
    const float* q = w.V(B2_V_QE0 + s.sel, c);
    float* gr = w.V(B2_V_GE0 + s.sel, c);
    const double lp = tc_finalize_chain(ws, gm, my_slot, g.lane(), K1, prior_tau, q, gr);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (n_staged > 0) { w.stk_hot = stk; w.stk_mask = stk_mask; w.stk_idx = stk_idx; }
    B2_STAMP(w, c, s, 2);
    const bool active = b2_advance<float, B2WarpGroup>(g, w, c, s, lp);
    __syncwarp();                                      // lanes wrote element i, read back as float4 rows
    {   // write the hot slots back (the likelihood kernel and the next launch read them from HBM/L2)
        const int lane = threadIdx.x & 31;
        if (wb_buf >= 0 && 4 * lane < w.Dp) {          // the one staged stack buffer that is still alive
            const int k = n_merge - 1;
#pragma unroll
            for (int which = 0; which < B2_S_NVEC; ++which)
                *reinterpret_cast<float4*>(w.Vglobal(B2_V_STACK0 + wb_buf * B2_S_NVEC + which, c) + 4 * lane) =
                    *reinterpret_cast<const float4*>(stk + ((size_t)k * B2_S_NVEC + which) * w.Dp + 4 * lane);
        }
        double* dst = w.lv + (size_t)c * 4 * B2_MAX_LEVELS;
        dst[lane] = lvh[lane];
        if (lane + 32 < 4 * B2_MAX_LEVELS) dst[lane + 32] = lvh[lane + 32];
        if (4 * lane < w.Dp) {
#pragma unroll
            for (int slot = 0; slot < B2_V_STACK0; ++slot)
                *reinterpret_cast<float4*>(w.Vglobal(slot, c) + 4 * lane) = *reinterpret_cast<const float4*>(hot + slot * w.Dp + 4 * lane);
        }
    }
    if (g.lane() == 0) {
        w.st[c] = s;
        if (active) {                                  // claim a slot in the next step's dense chain tiles
            const int ns = atomicAdd(ws.counters + (ws.parity ^ 1), 1);
            ws.chain_of_slot[ns] = c;
            ws.slot_of_chain[c] = ns;
        }
    }
    __syncwarp();
    if (w.dbg && c == 0) { __threadfence(); w.dbg[((s.n_grad - 1) & 4095) * 16 + 8] = clock64(); w.dbg[((s.n_grad - 1) & 4095) * 16 + 9] = s.leaf_n * 100 + s.depth; w.dbg[((s.n_grad - 1) & 4095) * 16 + 12] = s.iter; }
}

// ---------------------------------------------------------------------------------- host
struct TcHostState {
    TcWorkspace ws;
    bool ready;
};

bool b2_glm_tc_supported(const b2_engine* e) {
    return e->md.family == B2_FAMILY_GLM_LOGIT && e->dtype == B2_F32 && e->md.G + 1 <= TC_KP && e->md.N >= 1;
}

static int tc_setup(b2_engine* e, cudaStream_t stream) {
    TcHostState* hs = new TcHostState();
    memset(hs, 0, sizeof(*hs));
    TcWorkspace& w = hs->ws;
    const int N = e->md.N;
    w.n_tiles = (N + TC_OBS - 1) / TC_OBS;
    w.n_pad_rows = w.n_tiles * TC_OBS - N;
    w.chain_tiles = (e->C + TC_CHAINS - 1) / TC_CHAINS;
    w.c_pad = w.chain_tiles * TC_CHAINS;
    // fixed grid: a whole number of slabs per chain tile when every tile is live, one CTA per SM at most
    int splits = e->sm_count / w.chain_tiles;
    if (splits < 1) splits = 1;
    if (splits > w.n_tiles) splits = w.n_tiles;
    w.tiles_per_split = (w.n_tiles + splits - 1) / splits;
    w.splits = (w.n_tiles + w.tiles_per_split - 1) / w.tiles_per_split;
    w.grid_ctas = w.chain_tiles * w.splits;
    w.parity = 0;
    B2_CUDA_OK(cudaMalloc(&w.counters, 2 * sizeof(int)));
    B2_CUDA_OK(cudaMemsetAsync(w.counters, 0, 2 * sizeof(int), stream));
    B2_CUDA_OK(cudaMalloc(&w.chain_of_slot, (size_t)w.c_pad * sizeof(int)));
    B2_CUDA_OK(cudaMalloc(&w.slot_of_chain, (size_t)w.c_pad * sizeof(int)));
    B2_CUDA_OK(cudaMalloc(&w.xt, (size_t)w.n_tiles * TC_STAGE_DATA));
    B2_CUDA_OK(cudaMalloc(&w.gpart, (size_t)(w.grid_ctas + w.chain_tiles) * TC_CHAINS * TC_KP * sizeof(float)));
    B2_CUDA_OK(cudaMalloc(&w.lpart, (size_t)(w.grid_ctas + w.chain_tiles) * TC_EPI_GROUPS * TC_CHAINS * sizeof(double)));
    B2_CUDA_OK(cudaMalloc(&w.err, TC_ERR_INTS * sizeof(int)));
    w.dbg = nullptr;
    if (getenv("B2_TC_TIMELINE")) {
        B2_CUDA_OK(cudaMalloc(&w.dbg, (48 * TC_DBG_TILES + 4096 * 16) * sizeof(long long)));
        B2_CUDA_OK(cudaMemsetAsync(w.dbg, 0, (48 * TC_DBG_TILES + 4096 * 16) * sizeof(long long), stream));
    }
    B2_CUDA_OK(cudaMemsetAsync(w.err, 0, TC_ERR_INTS * sizeof(int), stream));
    k_glm_tc_prep_x<<<w.n_tiles, 256, 0, stream>>>(e->md.X, e->md.yf, N, e->md.G, w.xt, w.n_tiles);
    B2_CUDA_OK(cudaGetLastError());
    B2_CUDA_OK(cudaFuncSetAttribute(k_glm_tc_main, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    B2_CUDA_OK(cudaFuncSetAttribute(k_glm_tc_post, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    4 * TC_POST_STAGE * B2_S_NVEC * TC_KP * (int)sizeof(float)));
    e->launches += 1;
    hs->ready = true;
    e->glm_tc = hs;            // owned by the engine; released in b2_glm_tc_release
    return 0;
}

void b2_glm_tc_release(b2_engine* e) {
    if (!e->glm_tc) return;
    TcHostState* hs = (TcHostState*)e->glm_tc;
    cudaFree(hs->ws.dbg); cudaFree(hs->ws.xt); cudaFree(hs->ws.counters); cudaFree(hs->ws.chain_of_slot); cudaFree(hs->ws.slot_of_chain); cudaFree(hs->ws.gpart); cudaFree(hs->ws.lpart); cudaFree(hs->ws.err);
    delete hs;
    e->glm_tc = nullptr;
}

static int tc_ensure(b2_engine* e, cudaStream_t stream) {
    if (!e->glm_tc) return tc_setup(e, stream);
    return 0;
}

int b2_glm_tc_pack(b2_engine* e, const float* qA, const float* qB, int ld, const B2ChainState* st, int n, cudaStream_t stream) {
    // positions are split to bf16 hi/lo inside k_glm_tc_main (written straight into TMEM); only the
    // one-time X tiling has to exist before the first launch
    int rc = tc_ensure(e, stream);
    if (rc) return rc;
    TcWorkspace& w = ((TcHostState*)e->glm_tc)->ws;
    w.qA = qA; w.qB = qB; w.ld = ld; w.st = st; w.n_chains = n; w.K1 = e->md.G + 1;
    // dense slot <-> chain maps for the first launch; afterwards k_glm_tc_post maintains them
    B2_CUDA_OK(cudaMemsetAsync(w.counters, 0, 2 * sizeof(int), stream));
    w.parity = 0;
    k_glm_tc_compact<<<(n + 255) / 256, 256, 0, stream>>>(w, st, n);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 1;
    return 0;
}

int b2_glm_tc_main(b2_engine* e, cudaStream_t stream) {
    TcWorkspace& w = ((TcHostState*)e->glm_tc)->ws;
    k_glm_tc_main<<<w.grid_ctas, TC_THREADS, TC_SMEM_BYTES, stream>>>(w);
    e->launches += 1;
    return 0;
}

// debugging aid: copies the clock64 timeline of CTA (0,0) to the host (9 events x TC_DBG_TILES)
extern "C" int b2_debug_tc_timeline(b2_engine* e, long long* host_out) {
    if (!e || !e->glm_tc) return -1;
    TcWorkspace& w = ((TcHostState*)e->glm_tc)->ws;
    if (!w.dbg) return -2;
    B2_CUDA_OK(cudaDeviceSynchronize());
    B2_CUDA_OK(cudaMemcpy(host_out, w.dbg, 48 * TC_DBG_TILES * sizeof(long long), cudaMemcpyDeviceToHost));
    return 0;
}

// clock64 stamps of chain 0 inside k_glm_tc_post, one row of 16 per leapfrog (ring of 4096)
extern "C" int b2_debug_post_timeline(b2_engine* e, long long* host_out) {
    if (!e || !e->glm_tc) return -1;
    TcWorkspace& w = ((TcHostState*)e->glm_tc)->ws;
    if (!w.dbg) return -2;
    B2_CUDA_OK(cudaDeviceSynchronize());
    B2_CUDA_OK(cudaMemcpy(host_out, w.dbg + 48 * TC_DBG_TILES, 4096 * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    return 0;
}

int b2_glm_tc_post(b2_engine* e, const void* view_f32, cudaStream_t stream) {
    TcWorkspace& w = ((TcHostState*)e->glm_tc)->ws;
    const B2View<float>& v = *reinterpret_cast<const B2View<float>*>(view_f32);
    const size_t dyn = (size_t)4 * TC_POST_STAGE * B2_S_NVEC * e->Dp * sizeof(float);
    k_glm_tc_post<<<(e->C + 3) / 4, 128, dyn, stream>>>(w, v, e->md.G + 1, e->md.hp[0]);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 1;
    w.parity ^= 1;                                 // the next step reads the counter this launch filled
    return 0;
}

int b2_glm_tc_launch(b2_engine* e, const float* qA, const float* qB, float* gA, float* gB, int ld,
                     const B2ChainState* st, int n, double* logp, cudaStream_t stream) {
    int rc = b2_glm_tc_pack(e, qA, qB, ld, st, n, stream);
    if (rc) return rc;
    rc = b2_glm_tc_main(e, stream);
    if (rc) return rc;
    TcWorkspace& w = ((TcHostState*)e->glm_tc)->ws;
    k_glm_tc_finalize<<<(n + 3) / 4, 128, 0, stream>>>(w, n, e->md.G + 1, e->md.hp[0], qA, qB, gA, gB, ld, st, logp);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 1;
    return 0;
}
