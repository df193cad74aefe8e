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
//
// Round 2: (a) FUSED lock-step.  The chains are cut into two halves A | B and one launch runs the likelihood of
// one half (warps 0..19, as above) while four more warps (20..23) run the OTHER half's state machine
// (slab reduction -> b2_advance -> slot claim, the former k_glm_tc_post) on last launch's partials:
//     launch 2k: likelihood(A) + advance(B)      launch 2k+1: likelihood(B) + advance(A)
// so the latency-bound tree logic (26 us per step, 18 % of round 1's step, tensor pipe idle) hides under the
// other half's MMA work instead of running as a kernel of its own; stream order is the only dependency.
// 768 threads at 80 registers (the epilogue fits without spills; the state machine spills ~0.5 KB per thread,
// harmless at its issue rate).  (b) Epilogue v2: the MUFU pipe was the busiest unit (3 transcendentals per
// (chain, observation): ex2, rcp, lg2 -- 1536 of ~1850 cycles per tile).  log1p(e), e in (0, 1], is now a
// degree-7 polynomial evaluated two observations per instruction with Blackwell's packed fp32 math (FFMA2), and
// the rest of the elementwise math is packed too:  logp_i = (y - 1/2) eta - |eta| / 2 - log1p(exp(-|eta|)),
// r_i = (y - 1/2) - copysign(1 / (1 + exp(-|eta|)) - 1/2, eta)   -- 2 MUFU and ~13 issue slots per element.
#include <cuda_bf16.h>
#include <cstring>
#include <cstdlib>
// (the fused launch's state-machine warps would want B2_WELFORD_WIDTH 2 to stay near their 80 registers; the fused
//  schedule is off by default, so the stand-alone state-machine kernel keeps the 4-wide mass-matrix update)
#include "b2_engine.cuh"
#include "b2_tc_ptx.cuh"
#include "b2_glm_ref.cuh"

#define TC_CHAINS 128                 // MMA M
#define TC_OBS 64                     // observations per tile (GEMM1 N, GEMM2 K)
#define TC_POST_STAGE_MAX 6           // stack buffers the state-machine warps stage per chain (merge levels 0..5)
#define TC_KP 128                     // padded feature count (GEMM1 K, GEMM2 N)
#define TC_XPART_BYTES (TC_OBS * TC_KP * 2)            // 16384: one of {hi, lo}, two 64-column atoms
#define TC_Y_BYTES (2 * TC_OBS * 4)                    // 512: y | y - 1/2
#define TC_STAGE_DATA (2 * TC_XPART_BYTES + TC_Y_BYTES) // 33280 in global memory: Xhi | Xlo | y | y - 1/2
#define TC_STAGE_BYTES (2 * TC_XPART_BYTES)            // 32768 in shared memory (y lives in its own ring)
#define TC_YS_BYTES (TC_Y_BYTES + TC_OBS * 4)          // 768 in shared memory: y | y - 1/2 | eta_ref (see TcWorkspace::eta_ref)
#define TC_MAIN_SMEM(stages) (1024 + (stages) * (TC_STAGE_BYTES + TC_YS_BYTES) + 256)
#define TC_EPI_GROUPS 4               // epilogue warpgroups; group g owns observation columns 16g..16g+15 of a tile
#define TC_EPI_WARPS (4 * TC_EPI_GROUPS)
#define TC_THREADS (128 + 32 * TC_EPI_WARPS)          // 640: the likelihood's warps
#define TC_POST_WARPS 4                                // fused launches: warps 20..23 run the other half's state machine
#define TC_THREADS_FUSED (TC_THREADS + 32 * TC_POST_WARPS)
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
    int first, count;        // the chains this workspace covers: [first, first + count)
    int post_levels;         // merge levels whose stack buffers the state-machine warps stage in shared memory
    int x_policy;            // L2 hint of the X tile stream: 0 none, 1 evict_first (state and partials keep their lines), 2 evict_last
    int flush_tiles;         // the gradient accumulator is drained into the fp32 partials every flush_tiles tiles
    // Reference-centred positions.  Q enters GEMM1 as bf16 hi + lo, i.e. to ~17 bits RELATIVE TO |q|; near the
    // posterior mode that rounding of the position (2^-18 |q|) times the Hessian (N/4-ish) was the largest error of
    // the gradient (2.5e-4 of a typical-set gradient at C2, round 2 parity tests).  The chains of a run sit within a
    // few posterior sds of each other, so the GEMM works on dq = q - q_ref (q_ref = mean of the launch's live
    // positions, refreshed at the start of every run chunk) and eta_ref = Xa . q_ref, computed once per refresh in
    // fp64, is added back in the epilogue: the rounding is then relative to |dq|.
    const float* q_ref;      // [TC_KP]
    const float* eta_ref;    // [n_tiles * TC_OBS]
    int* err;                // device watchdog flag
    long long* role_clk;     // optional (B2_TC_ROLE_CLOCKS=1): {sum, count, max} cycles of the state-machine warps, then of the likelihood CTAs
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
        const float yv = row < N ? y[row] : 0.f;
        reinterpret_cast<float*>(blob + 2 * TC_XPART_BYTES)[r] = yv;
        reinterpret_cast<float*>(blob + 2 * TC_XPART_BYTES)[TC_OBS + r] = yv - 0.5f;       // epilogue v2 works on y - 1/2
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
__global__ void k_glm_tc_compact(TcWorkspace ws, const B2ChainState* st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ws.count) return;
    const int c = ws.first + i;
    const bool live = st ? b2_needs_grad(st[c].phase) : true;
    if (!live) return;
    const int slot = st ? atomicAdd(ws.counters + ws.parity, 1) : i;
    ws.chain_of_slot[slot] = c;
    ws.slot_of_chain[c] = slot;
    if (!st && i == 0) ws.counters[ws.parity] = ws.count;
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
    // (12 per round: 48 registers; 24 at once pushed the state-machine warps of the fused launch past their budget)
    for (int sp0 = 0; sp0 < gm.sp; sp0 += 12) {
        float4 v[12];
#pragma unroll
        for (int j = 0; j < 12; ++j)
            v[j] = (sp0 + j < gm.sp) ? __ldcg(gp + (size_t)(sp0 + j) * stride4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 12; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
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

// The chain state machine of the lock-step tensor-core path, for ONE chain on one warp:
//   slab reduction (-> logp, grad)  ->  b2_advance (the whole NUTS/HMC bookkeeping for this leapfrog)
//   ->  claim of a slot in the next step's dense chain tiles.
// Runs either as warps 20..23 of the fused launch (on the half whose likelihood ran in the PREVIOUS launch) or as
// the stand-alone kernel k_glm_tc_post (B2_TC_FUSED=0: round 1's two-kernel step, kept for A/B measurements).
// hot: [B2_V_STACK0][Dp] floats, lvh: [4][B2_MAX_LEVELS] doubles, stk: [levels][B2_S_NVEC][Dp] floats of shared memory
// owned by this warp.
__device__ __forceinline__ void tc_post_chain(const TcWorkspace& ws, B2View<float> w, int K1, double prior_tau, int c,
                                              float* hot, double* lvh, float* stk) {
    B2WarpGroup g;
    const long long t_start = w.dbg ? clock64() : 0;
    // The chain's 11 hot vector slots go global -> shared with cp.async (no register staging: the state-machine
    // warps of the fused launch run on a fixed register budget), in flight together with the level scalars, the
    // chain state and -- second group -- the stack buffers of the pending merges.
    const int lane0 = threadIdx.x & 31;
    if (4 * lane0 < w.Dp) {
#pragma unroll
        for (int slot = 0; slot < B2_V_STACK0; ++slot) {
            const uint32_t dst = smem_u32(hot + slot * w.Dp + 4 * lane0);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(w.Vglobal(slot, c) + 4 * lane0) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* lv_src = w.lv + (size_t)c * 4 * B2_MAX_LEVELS;
    const double lv_a0 = lv_src[lane0];
    const double lv_a1 = (lane0 + 32 < 4 * B2_MAX_LEVELS) ? lv_src[lane0 + 32] : 0.0;
    const int my_slot = ws.slot_of_chain[c];
    const TcGeom gm = tc_geom(ws);
    B2ChainState s = w.st[c];
    if (!b2_needs_grad(s.phase)) {                     // warp-uniform: nothing of this chain was evaluated
        asm volatile("cp.async.wait_all;" ::: "memory");
        return;
    }
    lvh[lane0] = lv_a0;
    if (lane0 + 32 < 4 * B2_MAX_LEVELS) lvh[lane0 + 32] = lv_a1;
    // Stack buffers the pending leaf will merge (one per trailing one-bit of its index, nuts.py:347-389 as a
    // binary counter): fetched with cp.async while the slab reduction below runs, so every merge level works
    // out of shared memory instead of paying 4-6 dependent L2 round trips (timeline r1: 5-7k cycles per level,
    // and the step lasts as long as its deepest merge chain).
    int n_merge = 0, n_staged = 0, wb_buf = -1;
    unsigned stk_mask = 0;
    unsigned long long stk_idx = 0;
    if (s.phase == B2_PHASE_TREE) {
        while ((s.leaf_n >> n_merge) & 1) ++n_merge;
        n_staged = n_merge < ws.post_levels ? n_merge : ws.post_levels;
        for (int k = 0; k < n_staged; ++k) {
            const int buf = b2_map_get(s.slot_map, k);
            stk_mask |= 1u << buf;
            stk_idx |= (unsigned long long)k << (4 * buf);
            if (4 * lane0 < w.Dp) {
#pragma unroll
                for (int which = 0; which < B2_S_NVEC; ++which) {
                    const float* src = w.Vglobal(B2_V_STACK0 + buf * B2_S_NVEC + which, c) + 4 * lane0;
                    const uint32_t dst = smem_u32(stk + ((size_t)k * B2_S_NVEC + which) * w.Dp + 4 * lane0);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                }
            }
        }
        if (n_staged == n_merge && n_merge > 0) wb_buf = b2_map_get(s.slot_map, n_merge - 1);   // the merged sub-tree ends up here
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");            // the hot slots have landed; the stack buffers may not have
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
    {   // write the hot slots back (the likelihood warps of the next launch read them from HBM/L2)
        if (wb_buf >= 0 && 4 * lane0 < w.Dp) {         // the one staged stack buffer that is still alive
            const int k = n_merge - 1;
#pragma unroll
            for (int which = 0; which < B2_S_NVEC; ++which)
                *reinterpret_cast<float4*>(w.Vglobal(B2_V_STACK0 + wb_buf * B2_S_NVEC + which, c) + 4 * lane0) =
                    *reinterpret_cast<const float4*>(stk + ((size_t)k * B2_S_NVEC + which) * w.Dp + 4 * lane0);
        }
        double* dst = w.lv + (size_t)c * 4 * B2_MAX_LEVELS;
        dst[lane0] = lvh[lane0];
        if (lane0 + 32 < 4 * B2_MAX_LEVELS) dst[lane0 + 32] = lvh[lane0 + 32];
        if (4 * lane0 < w.Dp) {
#pragma unroll
            for (int slot = 0; slot < B2_V_STACK0; ++slot)
                *reinterpret_cast<float4*>(w.Vglobal(slot, c) + 4 * lane0) = *reinterpret_cast<const float4*>(hot + slot * w.Dp + 4 * lane0);
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

// bytes of shared memory one state-machine warp needs
__host__ __device__ __forceinline__ size_t tc_post_warp_smem(int Dp, int levels) {
    return (size_t)B2_V_STACK0 * Dp * sizeof(float) + 4 * B2_MAX_LEVELS * sizeof(double) +
           (size_t)levels * B2_S_NVEC * Dp * sizeof(float);
}

// all chains of the workspace's range, `n_warps` warps of this block, block `blk` of `n_blk`
__device__ __forceinline__ void tc_post_role(const TcWorkspace& ws, const B2View<float>& w, int K1, double prior_tau,
                                             unsigned char* smem, int pw, int n_warps, int blk, int n_blk) {
    unsigned char* mine = smem + (size_t)pw * tc_post_warp_smem(w.Dp, ws.post_levels);
    double* lvh = reinterpret_cast<double*>(mine);                                  // 8-byte aligned first
    float* hot = reinterpret_cast<float*>(mine + 4 * B2_MAX_LEVELS * sizeof(double));
    float* stk = hot + (size_t)B2_V_STACK0 * w.Dp;
    for (int i = blk * n_warps + pw; i < ws.count; i += n_blk * n_warps) {
        tc_post_chain(ws, w, K1, prior_tau, ws.first + i, hot, lvh, stk);
        __syncwarp();                                  // the shared buffers are reused by the next chain
    }
}

// bar.sync over the likelihood's 640 threads only (the state-machine warps of a fused launch never join it)
#ifdef TC_V_SYNCTHREADS                               // timing variant (two-kernel build only)
__device__ __forceinline__ void tc_main_sync() { __syncthreads(); }
#else
__device__ __forceinline__ void tc_main_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory"); }
#endif

// degree-7 fit of log1p(e) / e on [0, 1] (Chebyshev nodes; |error| of e * P(e) in fp32 Horner form < 2.4e-7,
// the same as lg2.approx's): coefficients of e^0 .. e^7
#define TC_L1P_C0 9.999998102e-01f
#define TC_L1P_C1 -4.999744938e-01f
#define TC_L1P_C2 3.327617657e-01f
#define TC_L1P_C3 -2.449961172e-01f
#define TC_L1P_C4 1.775702399e-01f
#define TC_L1P_C5 -1.078536792e-01f
#define TC_L1P_C6 4.421419234e-02f
#define TC_L1P_C7 -8.574676205e-03f
__device__ __forceinline__ float2 tc_f2(float a) { return make_float2(a, a); }

// 16 observations of one chain row: S values -> residuals (bf16 hi | lo pairs) and this thread's logp terms.
// ys: 16 floats of y (EPI 0) or y - 1/2 (EPI 1) in shared memory.
template <int EPI>
__device__ __forceinline__ float tc_epilogue16(const uint32_t (&v)[16], uint32_t ys_addr, uint32_t ref_addr, uint32_t (&hi)[8],
                                               uint32_t (&lo)[8]) {
    float yv[16], rv[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                     // ld.shared (a generic pointer here compiled to LD, not LDS)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(yv[4 * i]), "=f"(yv[4 * i + 1]), "=f"(yv[4 * i + 2]), "=f"(yv[4 * i + 3]) : "r"(ys_addr + 16 * i));
#ifdef TC_V_NOETA                                     // timing variant: no reference (round 1's epilogue inputs)
        rv[4 * i] = rv[4 * i + 1] = rv[4 * i + 2] = rv[4 * i + 3] = 0.f;
#else
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(rv[4 * i]), "=f"(rv[4 * i + 1]), "=f"(rv[4 * i + 2]), "=f"(rv[4 * i + 3]) : "r"(ref_addr + 16 * i));
#endif
    }
    if (EPI == 0) {                                   // round 1: 3 MUFU per element (ex2, rcp, lg2)
        float lsum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float r2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float eta = __uint_as_float(v[2 * i + h]) + rv[2 * i + h];      // dq . x + eta_ref
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
            hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
            lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        return lsum;
    }
    // v2: two observations per instruction (packed fp32), 2 MUFU per element
    float2 ls_a = make_float2(0.f, 0.f), ls_b = make_float2(0.f, 0.f);
    float2 ls_c = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 eta = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])),
                                      make_float2(rv[2 * i], rv[2 * i + 1]));           // dq . x + eta_ref
        const float2 ym = make_float2(yv[2 * i], yv[2 * i + 1]);
        const float2 t = __fmul2_rn(eta, tc_f2(1.4426950408889634f));
        const float2 e = make_float2(tc_ex2(-fabsf(t.x)), tc_ex2(-fabsf(t.y)));          // exp(-|eta|) in (0, 1]
        const float2 w1 = __fadd2_rn(e, tc_f2(1.f));
        const float2 inv = make_float2(tc_rcp(w1.x), tc_rcp(w1.y));
        float2 pl = __ffma2_rn(tc_f2(TC_L1P_C7), e, tc_f2(TC_L1P_C6));
        pl = __ffma2_rn(pl, e, tc_f2(TC_L1P_C5));
        pl = __ffma2_rn(pl, e, tc_f2(TC_L1P_C4));
        pl = __ffma2_rn(pl, e, tc_f2(TC_L1P_C3));
        pl = __ffma2_rn(pl, e, tc_f2(TC_L1P_C2));
        pl = __ffma2_rn(pl, e, tc_f2(TC_L1P_C1));
        pl = __ffma2_rn(pl, e, tc_f2(TC_L1P_C0));
        ls_b = __ffma2_rn(e, pl, ls_b);                                                  // + log1p(exp(-|eta|))
        ls_a = __ffma2_rn(ym, eta, ls_a);                                                // + (y - 1/2) eta
        ls_c = __fadd2_rn(ls_c, make_float2(fabsf(t.x), fabsf(t.y)));                    // + |eta| log2(e)
        // sigmoid(eta) - 1/2 = copysign(1 / (1 + exp(-|eta|)) - 1/2, eta):  r = (y - 1/2) + copysign(inv - 1/2, -eta)
        const float2 d = __fadd2_rn(inv, tc_f2(-0.5f));
        const float2 sg = make_float2(__uint_as_float(__float_as_uint(d.x) | (~__float_as_uint(eta.x) & 0x80000000u)),
                                      __uint_as_float(__float_as_uint(d.y) | (~__float_as_uint(eta.y) & 0x80000000u)));
        const float2 r = __fadd2_rn(ym, sg);
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(r.x, r.y);
        const float2 back = __bfloat1622float2(h2);
        const float2 rl = __ffma2_rn(back, tc_f2(-1.f), r);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(rl.x, rl.y);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    // y eta - softplus(eta) = (y - 1/2) eta - |eta| / 2 - log1p(exp(-|eta|))
    return fmaf(-0.34657359027997264f, ls_c.x + ls_c.y, (ls_a.x + ls_a.y) - (ls_b.x + ls_b.y));
}

// STAGES: depth of the X tile ring; EPI: epilogue version (0 = round 1, 1 = packed fp32 + polynomial log1p);
// FUSED: warps 20..23 run the state machine of the chains of `wsp` (the other half) on the side.
template <int STAGES, int EPI, bool FUSED>
__global__ void __launch_bounds__(FUSED ? TC_THREADS_FUSED : TC_THREADS, 1)
k_glm_tc_main(TcWorkspace ws, TcWorkspace wsp, B2View<float> wview, double prior_tau) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* x_s = smem;
    unsigned char* y_s = x_s + STAGES * TC_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(y_s + STAGES * TC_YS_BYTES);
    uint64_t* q_full = bars;                       // 1
    uint64_t* x_full = bars + 1;                   // STAGES
    uint64_t* x_empty = x_full + STAGES;        // STAGES
    uint64_t* s_full = x_empty + STAGES;        // 2
    uint64_t* s_empty = s_full + 2;                // 2
    uint64_t* p_full = s_empty + 2;                // 2
    uint64_t* p_empty = p_full + 2;                // 2
    uint64_t* g_full = p_empty + 2;                // 1: completes once per flush chunk (GEMM2 of the chunk's last tile)
    uint64_t* g_empty = g_full + 1;                // 1: every epilogue warp has read the chunk's G out of TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (FUSED && warp >= TC_THREADS / 32) {
        // ===== state machine of the other half (its likelihood ran in the previous launch) =====
        const long long tp0 = clock64();
        if (wsp.count > 0)
            tc_post_role(wsp, wview, wsp.K1, prior_tau, smem + (TC_MAIN_SMEM(STAGES) - 1024), warp - TC_THREADS / 32,
                         TC_POST_WARPS, blockIdx.x, gridDim.x);
        if (ws.role_clk && lane == 0 && wsp.count > 0 && blockIdx.x * TC_POST_WARPS + warp - TC_THREADS / 32 < wsp.count) {
            const unsigned long long dt = (unsigned long long)(clock64() - tp0);
            atomicAdd((unsigned long long*)ws.role_clk + 0, dt);
            atomicAdd((unsigned long long*)ws.role_clk + 1, 1ull);
            atomicMax((unsigned long long*)ws.role_clk + 2, dt);
        }
        return;
    }
    if (ws.count == 0) return;
    const TcGeom gm = tc_geom(ws);
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.counters[ws.parity ^ 1] = 0;   // the state machine of this half refills it
    const int ctile = blockIdx.x / gm.sp, split = blockIdx.x % gm.sp;
    if (gm.nt == 0 || ctile >= gm.nt) return;      // whole CTA: no live chain tile for it
    const int t_begin = split * gm.tps;
    const int t_end = min(ws.n_tiles, t_begin + gm.tps);
    const int T = t_end - t_begin;                 // >= 1: gm.sp counts only slabs that own rows

    // The two kinds of warps are kept in lexically separate branches (each with its own copy of the two CTA
    // barriers) so that ptxas can give the control branch its own, smaller register budget.
    if (warp < 4) {
        // control warpgroup: 0 = bulk-copy producer, 1 = GEMM1 issuer, 2 = TMEM allocator, 3 = GEMM2 issuer
        const long long tm0 = clock64();
        if (warp == 1 && lane == 0) {
            mbar_init(q_full, TC_EPI_WARPS);
            for (int i = 0; i < STAGES; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, TC_EPI_WARPS / 2); mbar_init(p_full + i, TC_EPI_WARPS / 2); mbar_init(p_empty + i, 1); }
            mbar_init(g_full, 1);
            mbar_init(g_empty, TC_EPI_WARPS);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 2) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        tc_main_sync();
        tc_fence_after();
        const uint32_t tmem = *tmem_slot;
        if (warp == 0) {
            // ===== producer: Q tile once, then the X tile ring =====
            if (lane == 0) {
                const uint64_t x_pol = ws.x_policy == 2 ? l2_policy_evict_last() : l2_policy_evict_first();
                for (int t = 0; t < T; ++t) {
                    const int s = t % STAGES;
                    if (t >= STAGES) mbar_wait(x_empty + s, ((t / STAGES) - 1) & 1, ws.err, 1);
                    const unsigned char* src = ws.xt + (size_t)(t_begin + t) * TC_STAGE_DATA;
                    mbar_expect_tx(x_full + s, TC_STAGE_DATA + TC_OBS * 4);
                    if (ws.x_policy) bulk_g2s_hint(x_s + s * TC_STAGE_BYTES, src, TC_STAGE_BYTES, x_full + s, x_pol);
                    else bulk_g2s(x_s + s * TC_STAGE_BYTES, src, TC_STAGE_BYTES, x_full + s);
                    bulk_g2s(y_s + s * TC_YS_BYTES, src + TC_STAGE_BYTES, TC_Y_BYTES, x_full + s);
                    bulk_g2s(y_s + s * TC_YS_BYTES + TC_Y_BYTES, ws.eta_ref + (size_t)(t_begin + t) * TC_OBS, TC_OBS * 4, x_full + s);
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
                const int s = t % STAGES, b = t & 1;
                if (lane == 0) TC_STAMP(0, t);
                mbar_wait(x_full + s, (t / STAGES) & 1, ws.err, 3);
                if (t >= 2) mbar_wait(s_empty + b, ((t >> 1) - 1) & 1, ws.err, 4);
                tc_fence_after();
                const uint32_t x_addr = smem_u32(x_s + s * TC_STAGE_BYTES);
                const uint32_t d = tmem + TC_COL_S + 64 * b;
                if (elect_one()) {
                    uint32_t acc = 0;
                    // The tensor core adds into its fp32 accumulator with truncation (every add loses up to an ulp of
                    // the running sum, towards zero: measured as a relative bias of ~6e-7 on eta = 0.03 nats on a logp
                    // of -7e4 with the large term first), so the two small cross terms are accumulated first, while
                    // the sum is still ~2^-8 of its final size: Qlo.Xhi, Qhi.Xlo, then Qhi.Xhi.
    #pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t qa = tmem + TC_COL_Q + (pass == 0 ? 64 : 0);
                        const uint32_t xa = x_addr + (pass == 1 ? TC_XPART_BYTES : 0);
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
            // The accumulator adds with truncation (up to an ulp of the RUNNING sum per MMA, towards zero): 1044 adds
            // into one G biased the gradient by 1e-4 of its size near the posterior mode (round 2 parity test at C2
            // size), so G is drained into the slab's fp32 partial (round-to-nearest adds) every F tiles.
            const int F = ws.flush_tiles;
            for (int u = 0; u < T; ++u) {
                const int s = u % STAGES, b = u & 1;
                const bool chunk_first = (u % F) == 0, chunk_last = ((u + 1) % F) == 0 || u == T - 1;
                if (lane == 0) TC_STAMP(2, u);
                mbar_wait(p_full + b, (u >> 1) & 1, ws.err, 5);
                if (chunk_first && u > 0) mbar_wait(g_empty, ((u / F) - 1) & 1, ws.err, 10);   // the previous chunk's G has been read out
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
                            mma_ts(d, pa + j * 8, bd, idesc_g2, (!chunk_first || pass > 0 || j > 0) ? 1u : 0u);
                        }
                    }
                    tc_commit(x_empty + s);
                    tc_commit(p_empty + b);
                    if (chunk_last) tc_commit(g_full);
                    TC_STAMP(3, u);
                }
                __syncwarp();
            }
        }
        tc_fence_before();
        tc_main_sync();
        if (warp == 2) {
            tc_fence_after();
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
            if (ws.role_clk && lane == 0) {
                const unsigned long long dt = (unsigned long long)(clock64() - tm0);
                atomicAdd((unsigned long long*)ws.role_clk + 3, dt);
                atomicAdd((unsigned long long*)ws.role_clk + 4, 1ull);
                atomicMax((unsigned long long*)ws.role_clk + 5, dt);
            }
        }
    } else {
        tc_fence_before();
        tc_main_sync();                                            // barriers initialised, TMEM allocated
        tc_fence_after();
        const uint32_t tmem = *tmem_slot;
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
                const float a = (live && k < ws.K1) ? q[k] - ws.q_ref[k] : 0.f;            // dq = q - q_ref
                const float b2 = (live && k + 1 < ws.K1) ? q[k + 1] - ws.q_ref[k + 1] : 0.f;
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
        // Chunked gradient accumulation (see the GEMM2 issuer): after GEMM2 of a chunk's last tile every epilogue
        // warp adds its 32 columns of G into the slab's fp32 partial in global memory -- the same thread owns the
        // same elements every time, so plain load-add-store -- and releases the accumulator.
        const int F = ws.flush_tiles;
        const int n_chunks = (T + F - 1) / F;
        int k_drain = 0;                                           // next chunk this warp has to drain
        float* gout = ws.gpart + ((size_t)split * gm.stride + ctile * TC_CHAINS + row) * TC_KP + 32 * cg;
        auto drain = [&](int k) {
            mbar_wait(g_full, k & 1, ws.err, 8);
            tc_fence_after();
            uint32_t g32[32];
            TC_LD32(tmem + lane_addr + TC_COL_G + 32 * cg, g32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(g_empty);                   // G is in registers: the next chunk may start
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
        for (int t = pair; t < T; t += 2) {
            // before touching a tile of a new chunk, drain the chunks that end before it (draining after the tile
            // instead can deadlock: GEMM2 of the next chunk waits for all sixteen warps, see b2_glm_tcw.cu)
            while (k_drain < n_chunks - 1 && (k_drain + 1) * F - 1 < t) { drain(k_drain); ++k_drain; }
            const int s = t % STAGES, b = pair;
            // this warp's 32 observations of the stage's y block: y (EPI 0) or y - 1/2 (EPI 1)
            const uint32_t ys_addr = smem_u32(y_s + s * TC_YS_BYTES) + (EPI ? TC_OBS * 4 : 0) + 128 * sub;
            const uint32_t ref_addr = smem_u32(y_s + s * TC_YS_BYTES) + TC_Y_BYTES + 128 * sub;
            const bool stamp = (lane == 0) && (sub == 0) && (wq == 0);
            if (stamp) TC_STAMP(4, t);
            mbar_wait(x_full + s, (t / STAGES) & 1, ws.err, 9);    // y values of this stage (async-proxy writes)
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
            for (int hh = 0; hh < 2; ++hh)
                lsum += tc_epilogue16<EPI>(v[hh], ys_addr + 64 * hh, ref_addr + 64 * hh, hi[hh], lo[hh]);
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
        // what is left: at least the last chunk, G[chain row][128 features] -> global partials (32 columns per group)
        while (k_drain < n_chunks) { drain(k_drain); ++k_drain; }
        const int slot = ctile * TC_CHAINS + row;
        // logp partial of this column group; the finalize kernel adds the TC_EPI_GROUPS partials
        ws.lpart[((size_t)split * TC_EPI_GROUPS + cg) * gm.stride + slot] = logp;
        tc_fence_before();
        tc_main_sync();                                            // every TMEM read is done: warp 2 may free it
    }
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

// Stand-alone state-machine kernel of the two-kernel step (B2_TC_FUSED=0, round 1's schedule; A/B measurements)
__global__ void __launch_bounds__(128) k_glm_tc_post(TcWorkspace ws, B2View<float> w, double prior_tau) {
    extern __shared__ __align__(16) unsigned char post_dyn[];
    w.dbg = ws.dbg ? ws.dbg + 48 * TC_DBG_TILES : nullptr;          // post stamps live behind the likelihood's
    tc_post_role(ws, w, ws.K1, prior_tau, post_dyn, threadIdx.x >> 5, 4, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------------------------- host
#define TC_SMEM_LIMIT 232448                     // 227 KB of dynamic shared memory per CTA on sm_100

struct TcHostState {
    TcWorkspace full;          // every chain in one workspace: b2_logp_dlogp, the stepwise API, the two-kernel step
    TcWorkspace half[2];       // fused lock-step: chains [0, h) | [h, C), one likelihood launch each per leapfrog
    bool pending[2];           // half's likelihood has run, its state machine has not consumed the partials yet
    bool fused;                // B2_TC_FUSED (default 0)
    int epi;                   // B2_TC_EPI   (default 0)
    int stages_unfused;        // X ring depth of the two-kernel build: 4 (default), 5 or 6 (B2_TC_STAGES_UNFUSED)
    int stages_fused;          // X ring depth of the fused kernel (5, or 4 when the state-machine warps need the room)
    int post_levels;
};

bool b2_glm_tc_supported(const b2_engine* e) {
    return e->md.family == B2_FAMILY_GLM_LOGIT && e->dtype == B2_F32 && e->md.G + 1 <= TC_KP && e->md.N >= 1;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// geometry + private buffers of one workspace covering chains [first, first + count)
static int tc_ws_init(b2_engine* e, TcWorkspace& w, const TcWorkspace& shared, int first, int count, cudaStream_t stream) {
    w = shared;                                   // xt, err, dbg, n_tiles, n_pad_rows
    w.first = first; w.count = count;
    w.chain_tiles = (count + TC_CHAINS - 1) / TC_CHAINS;
    w.c_pad = ((e->C + TC_CHAINS - 1) / TC_CHAINS) * TC_CHAINS;        // maps are indexed by global chain id
    // fixed grid: a whole number of slabs per chain tile when every tile is live, one CTA per SM at most
    int splits = w.chain_tiles > 0 ? e->sm_count / w.chain_tiles : 1;
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
    const size_t rows = (size_t)(w.grid_ctas + w.chain_tiles + 1) * TC_CHAINS;
    B2_CUDA_OK(cudaMalloc(&w.gpart, rows * TC_KP * sizeof(float)));
    B2_CUDA_OK(cudaMalloc(&w.lpart, rows * TC_EPI_GROUPS * sizeof(double)));
    return 0;
}

static void tc_ws_free(TcWorkspace& w) {
    cudaFree(w.counters); cudaFree(w.chain_of_slot); cudaFree(w.slot_of_chain); cudaFree(w.gpart); cudaFree(w.lpart);
}

template <int STAGES, int EPI, bool FUSED>
static int tc_main_attr(size_t smem) {
    B2_CUDA_OK(cudaFuncSetAttribute(k_glm_tc_main<STAGES, EPI, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}

static size_t tc_fused_smem(const TcHostState* hs, int Dp) {
    return TC_MAIN_SMEM(hs->stages_fused) + TC_POST_WARPS * tc_post_warp_smem(Dp, hs->post_levels);
}

static int tc_setup(b2_engine* e, cudaStream_t stream) {
    TcHostState* hs = new TcHostState();
    memset(hs, 0, sizeof(*hs));
    TcWorkspace shared;
    memset(&shared, 0, sizeof(shared));
    const int N = e->md.N;
    shared.n_tiles = (N + TC_OBS - 1) / TC_OBS;
    shared.n_pad_rows = shared.n_tiles * TC_OBS - N;
    // drains of the gradient accumulator: off by default here (<= 87 tiles = 1044 truncating adds per slab at C2,
    // a bias of ~3e-5 of a slab partial; each drain costs the pipeline 1-2 tile periods).  B2_TC_FLUSH=32 enables.
    shared.flush_tiles = env_int("B2_TC_FLUSH", 1 << 20);
    // X tiles are streamed once per launch (52 MB at C2, more than a launch leaves of L2 anyway): marking their lines
    // evict-first keeps the chains' state and the slab partials in L2 for the state-machine kernel (21.1 vs 22.4 us per
    // step, likelihood +1 us: +1 % on the job, profiles/README.md).  B2_TC_XPOLICY=0 none, 2 evict_last.
    shared.x_policy = env_int("B2_TC_XPOLICY", 1);
    if (shared.flush_tiles < 4) shared.flush_tiles = 4;
    B2_CUDA_OK(cudaMalloc(&shared.xt, (size_t)shared.n_tiles * TC_STAGE_DATA));
    B2_CUDA_OK(cudaMalloc(&shared.err, TC_ERR_INTS * sizeof(int)));
    B2_CUDA_OK(cudaMemsetAsync(shared.err, 0, TC_ERR_INTS * sizeof(int), stream));
    float *q_ref = nullptr, *eta_ref = nullptr;
    B2_CUDA_OK(cudaMalloc(&q_ref, TC_KP * sizeof(float)));
    B2_CUDA_OK(cudaMalloc(&eta_ref, (size_t)shared.n_tiles * TC_OBS * sizeof(float)));
    B2_CUDA_OK(cudaMemsetAsync(q_ref, 0, TC_KP * sizeof(float), stream));
    B2_CUDA_OK(cudaMemsetAsync(eta_ref, 0, (size_t)shared.n_tiles * TC_OBS * sizeof(float), stream));
    shared.q_ref = q_ref; shared.eta_ref = eta_ref;
    if (getenv("B2_TC_TIMELINE")) {
        B2_CUDA_OK(cudaMalloc(&shared.dbg, (48 * TC_DBG_TILES + 4096 * 16) * sizeof(long long)));
        B2_CUDA_OK(cudaMemsetAsync(shared.dbg, 0, (48 * TC_DBG_TILES + 4096 * 16) * sizeof(long long), stream));
    }
    if (getenv("B2_TC_ROLE_CLOCKS")) {
        B2_CUDA_OK(cudaMalloc(&shared.role_clk, 8 * sizeof(long long)));
        B2_CUDA_OK(cudaMemsetAsync(shared.role_clk, 0, 8 * sizeof(long long), stream));
    }
    // Default: the two-kernel step.  The fused launch measured SLOWER on B200 (round 2, profiles/README.md): the
    // state-machine warps' shuffles and shared-memory accesses queue behind the epilogue's MUFU / TMEM traffic in the
    // SM's MIO pipe and a transition end takes 150-185 k cycles instead of ~50 k stand-alone.
    hs->fused = env_int("B2_TC_FUSED", 0) != 0 && !shared.dbg;      // the timeline tools read the two-kernel step
    // Defaults from round 2's A/B runs on one box (profiles/README.md): (1) the packed-fp32 / polynomial epilogue
    // (EPI 1) takes a third of the MUFU work away but is 1 % SLOWER in the lock-step job (FMA pipe 28 -> 41 %);
    // (2) a 6-stage ring plus the y / y - 1/2 / eta_ref rows needs 197.75 KB, which tips the launch from the 196 KB
    // into the 228 KB shared-memory carve-out (L1 60 -> 28 KB) and costs 5 % (122.5 -> 128.7 us per launch, the same
    // with a 5-stage ring padded to that size); 5 stages fit the 196 KB carve-out (123.1 us), 4 stages the 164 KB
    // one (121.1 us) -- the ring is deep enough for L2-resident tiles either way, so 4 it is.
    hs->epi = env_int("B2_TC_EPI", 0) != 0 ? 1 : 0;
    const int su = env_int("B2_TC_STAGES_UNFUSED", 4);
    hs->stages_unfused = (su >= 3 && su <= 6) ? su : 4;
    // shared memory of the fused launch: X ring + the four state-machine warps (hot slots + staged merge levels);
    // prefer the deeper ring, stage as many merge levels as still fit
    hs->stages_fused = env_int("B2_TC_STAGES", 5) <= 4 ? 4 : 5;
    hs->post_levels = env_int("B2_TC_POST_LEVELS", TC_POST_STAGE_MAX);
    if (hs->post_levels > TC_POST_STAGE_MAX) hs->post_levels = TC_POST_STAGE_MAX;
    const size_t smem_cap = (size_t)env_int("B2_TC_SMEM_CAP", TC_SMEM_LIMIT);
    for (;;) {
        if (tc_fused_smem(hs, e->Dp) <= smem_cap) break;
        if (hs->post_levels > 3) { hs->post_levels -= 1; continue; }
        if (hs->stages_fused == 5) { hs->stages_fused = 4; hs->post_levels = TC_POST_STAGE_MAX; continue; }
        hs->post_levels -= 1;
        if (hs->post_levels < 0) { b2_set_error("tcgen05 GLM path: state-machine warps do not fit shared memory"); return -6; }
    }
    int rc = tc_ws_init(e, hs->full, shared, 0, e->C, stream);
    if (rc) return rc;
    // halves: whole chain tiles, A gets the extra one
    const int tiles = (e->C + TC_CHAINS - 1) / TC_CHAINS;
    const int hA = ((tiles + 1) / 2) * TC_CHAINS < e->C ? ((tiles + 1) / 2) * TC_CHAINS : e->C;
    if (hs->fused) {                                   // the half workspaces exist only for the fused schedule
        rc = tc_ws_init(e, hs->half[0], shared, 0, hA, stream);
        if (rc) return rc;
        rc = tc_ws_init(e, hs->half[1], shared, hA, e->C - hA, stream);
        if (rc) return rc;
        hs->half[0].post_levels = hs->half[1].post_levels = hs->post_levels;
    }
    hs->full.post_levels = TC_POST_STAGE_MAX;
    k_glm_tc_prep_x<<<shared.n_tiles, 256, 0, stream>>>(e->md.X, e->md.yf, N, e->md.G, shared.xt, shared.n_tiles);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 1;
    const size_t fs = tc_fused_smem(hs, e->Dp);
    // only the instantiation this engine launches (each first touch of a kernel loads its code: module loading is lazy)
    const int su_ = hs->stages_unfused, ep_ = hs->epi;
#define TC_ATTR(S, E, F, BYTES) do { if ((rc = tc_main_attr<S, E, F>(BYTES))) return rc; } while (0)
    if (!hs->fused || env_int("B2_TC_HOOK_FUSED", 0) == 0) {
        if (su_ == 3) { if (ep_) TC_ATTR(3, 1, false, TC_MAIN_SMEM(3)); else TC_ATTR(3, 0, false, TC_MAIN_SMEM(3)); }
        else if (su_ == 4) { if (ep_) TC_ATTR(4, 1, false, TC_MAIN_SMEM(4)); else TC_ATTR(4, 0, false, TC_MAIN_SMEM(4)); }
        else if (su_ == 5) { if (ep_) TC_ATTR(5, 1, false, TC_MAIN_SMEM(6)); else TC_ATTR(5, 0, false, TC_MAIN_SMEM(6)); }
        else { if (ep_) TC_ATTR(6, 1, false, TC_MAIN_SMEM(6)); else TC_ATTR(6, 0, false, TC_MAIN_SMEM(6)); }
    }
    if (hs->fused || env_int("B2_TC_HOOK_FUSED", 0)) {
        if (hs->stages_fused == 5) { if (ep_) TC_ATTR(5, 1, true, fs); else TC_ATTR(5, 0, true, fs); }
        else { if (ep_) TC_ATTR(4, 1, true, fs); else TC_ATTR(4, 0, true, fs); }
    }
#undef TC_ATTR
    B2_CUDA_OK(cudaFuncSetAttribute(k_glm_tc_post, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(4 * tc_post_warp_smem(e->Dp, TC_POST_STAGE_MAX))));
    e->glm_tc = hs;            // owned by the engine; released in b2_glm_tc_release
    return 0;
}

void b2_glm_tc_release(b2_engine* e) {
    if (!e->glm_tc) return;
    TcHostState* hs = (TcHostState*)e->glm_tc;
    cudaFree(hs->full.dbg); cudaFree(hs->full.xt); cudaFree(hs->full.err); cudaFree(hs->full.role_clk);
    cudaFree((void*)hs->full.q_ref); cudaFree((void*)hs->full.eta_ref);
    tc_ws_free(hs->full); tc_ws_free(hs->half[0]); tc_ws_free(hs->half[1]);
    delete hs;
    e->glm_tc = nullptr;
}

static int tc_ensure(b2_engine* e, cudaStream_t stream) {
    if (!e->glm_tc) return tc_setup(e, stream);
    return 0;
}

// q_ref = mean pending position of the chains about to be evaluated, eta_ref = Xa . q_ref (see TcWorkspace::q_ref);
// shared by every workspace of the engine, refreshed at the start of a run chunk / before a likelihood-only launch
static int tc_refresh_reference(b2_engine* e, const TcWorkspace& w, const float* qA, const float* qB, int ld,
                                const B2ChainState* st, int n, cudaStream_t stream) {
    if (env_int("B2_TC_NOREF", 0)) return 0;                  // A/B switch: positions relative to zero (round 1)
    const int K1 = e->md.G + 1;
    k_glm_ref_mean<<<TC_KP, 256, 0, stream>>>(qA, qB, ld, st, 0, n, K1, const_cast<float*>(w.q_ref), TC_KP);
    const int n_pad = w.n_tiles * TC_OBS;
    k_glm_ref_eta<<<(n_pad + 7) / 8, 256, 0, stream>>>(e->md.X, e->md.N, e->md.G, w.q_ref, 1, const_cast<float*>(w.eta_ref), n_pad);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 2;
    return 0;
}

// dense slot <-> chain maps of one workspace from the chain states (first launch of a run / parity hook)
static int tc_compact(b2_engine* e, TcWorkspace& w, const float* qA, const float* qB, int ld, const B2ChainState* st,
                      int count, cudaStream_t stream) {
    w.qA = qA; w.qB = qB; w.ld = ld; w.st = st; w.n_chains = count; w.K1 = e->md.G + 1;
    w.parity = 0;
    B2_CUDA_OK(cudaMemsetAsync(w.counters, 0, 2 * sizeof(int), stream));
    if (count > 0) {
        TcWorkspace tmp = w;
        tmp.count = count;
        k_glm_tc_compact<<<(count + 255) / 256, 256, 0, stream>>>(tmp, st);
        B2_CUDA_OK(cudaGetLastError());
        e->launches += 1;
    }
    return 0;
}

// start of a lock-step run: positions are split to bf16 hi/lo inside the likelihood warps (written straight into
// TMEM), so only the one-time X tiling and the slot maps have to exist before the first launch
int b2_glm_tc_pack(b2_engine* e, const float* qA, const float* qB, int ld, const B2ChainState* st, int n, cudaStream_t stream) {
    int rc = tc_ensure(e, stream);
    if (rc) return rc;
    TcHostState* hs = (TcHostState*)e->glm_tc;
    if ((rc = tc_refresh_reference(e, hs->full, qA, qB, ld, st, n, stream))) return rc;
    if (!hs->fused) return tc_compact(e, hs->full, qA, qB, ld, st, n, stream);
    for (int h = 0; h < 2; ++h) {
        if ((rc = tc_compact(e, hs->half[h], qA, qB, ld, st, hs->half[h].count, stream))) return rc;
        hs->pending[h] = false;
    }
    return 0;
}

template <bool FUSED>
static int tc_launch(b2_engine* e, const TcHostState* hs, const TcWorkspace& ws, const TcWorkspace& wsp, const B2View<float>& v,
                     cudaStream_t stream) {
    int grid = ws.count > 0 ? ws.grid_ctas : 0;
    size_t smem = TC_MAIN_SMEM(6);
    if (FUSED) {
        int post_blocks = (wsp.count + TC_POST_WARPS - 1) / TC_POST_WARPS;
        if (post_blocks > e->sm_count) post_blocks = e->sm_count;
        if (grid < post_blocks) grid = post_blocks;
        smem = tc_fused_smem(hs, e->Dp);
    }
    if (grid <= 0) return 0;
    const double tau = e->md.hp[0];
    if (!FUSED) {
        if (hs->stages_unfused == 5) {
            smem = TC_MAIN_SMEM(5) + (size_t)env_int("B2_TC_SMEM_PAD", 0);      // experiment: push the launch into the 228 KB carve-out
            if (hs->epi) k_glm_tc_main<5, 1, false><<<grid, TC_THREADS, smem, stream>>>(ws, wsp, v, tau);
            else k_glm_tc_main<5, 0, false><<<grid, TC_THREADS, smem, stream>>>(ws, wsp, v, tau);
        } else if (hs->stages_unfused == 3) {
            smem = TC_MAIN_SMEM(3);
            if (hs->epi) k_glm_tc_main<3, 1, false><<<grid, TC_THREADS, smem, stream>>>(ws, wsp, v, tau);
            else k_glm_tc_main<3, 0, false><<<grid, TC_THREADS, smem, stream>>>(ws, wsp, v, tau);
        } else if (hs->stages_unfused == 4) {
            smem = TC_MAIN_SMEM(4);
            if (hs->epi) k_glm_tc_main<4, 1, false><<<grid, TC_THREADS, smem, stream>>>(ws, wsp, v, tau);
            else k_glm_tc_main<4, 0, false><<<grid, TC_THREADS, smem, stream>>>(ws, wsp, v, tau);
        } else if (hs->epi) k_glm_tc_main<6, 1, false><<<grid, TC_THREADS, smem, stream>>>(ws, wsp, v, tau);
        else k_glm_tc_main<6, 0, false><<<grid, TC_THREADS, smem, stream>>>(ws, wsp, v, tau);
    } else if (hs->stages_fused == 5) {
        if (hs->epi) k_glm_tc_main<5, 1, true><<<grid, TC_THREADS_FUSED, smem, stream>>>(ws, wsp, v, tau);
        else k_glm_tc_main<5, 0, true><<<grid, TC_THREADS_FUSED, smem, stream>>>(ws, wsp, v, tau);
    } else {
        if (hs->epi) k_glm_tc_main<4, 1, true><<<grid, TC_THREADS_FUSED, smem, stream>>>(ws, wsp, v, tau);
        else k_glm_tc_main<4, 0, true><<<grid, TC_THREADS_FUSED, smem, stream>>>(ws, wsp, v, tau);
    }
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 1;
    return 0;
}

// One leapfrog of every live chain (lock-step run).  Fused schedule: two launches,
//   likelihood(A) + state machine(B, on the previous launch's partials), then likelihood(B) + state machine(A);
// two-kernel schedule (B2_TC_FUSED=0): likelihood(all), then the state-machine kernel.
int b2_glm_tc_step(b2_engine* e, const void* view_f32, cudaStream_t stream, cudaEvent_t mid) {
    TcHostState* hs = (TcHostState*)e->glm_tc;
    const B2View<float>& v = *reinterpret_cast<const B2View<float>*>(view_f32);
    int rc;
    if (!hs->fused) {
        TcWorkspace none;
        memset(&none, 0, sizeof(none));
        if ((rc = tc_launch<false>(e, hs, hs->full, none, v, stream))) return rc;
        if (mid) cudaEventRecord(mid, stream);
        const size_t dyn = 4 * tc_post_warp_smem(e->Dp, TC_POST_STAGE_MAX);
        k_glm_tc_post<<<(e->C + 3) / 4, 128, dyn, stream>>>(hs->full, v, e->md.hp[0]);
        B2_CUDA_OK(cudaGetLastError());
        e->launches += 1;
        hs->full.parity ^= 1;                          // the next step reads the counter this launch filled
        return 0;
    }
    TcWorkspace none;
    memset(&none, 0, sizeof(none));
    for (int h = 0; h < 2; ++h) {
        const int o = 1 - h;
        const bool post = hs->pending[o];
        if ((rc = tc_launch<true>(e, hs, hs->half[h], post ? hs->half[o] : none, v, stream))) return rc;
        if (post) { hs->half[o].parity ^= 1; hs->pending[o] = false; }
        hs->pending[h] = hs->half[h].count > 0;
    }
    if (mid) cudaEventRecord(mid, stream);
    return 0;
}

bool b2_glm_tc_is_fused(const b2_engine* e) { return e->glm_tc && ((const TcHostState*)e->glm_tc)->fused; }

// debugging aid: copies the clock64 timeline of CTA (0,0) to the host (9 events x TC_DBG_TILES)
extern "C" int b2_debug_tc_timeline(b2_engine* e, long long* host_out) {
    if (!e || !e->glm_tc) return -1;
    TcWorkspace& w = ((TcHostState*)e->glm_tc)->full;
    if (!w.dbg) return -2;
    B2_CUDA_OK(cudaDeviceSynchronize());
    B2_CUDA_OK(cudaMemcpy(host_out, w.dbg, 48 * TC_DBG_TILES * sizeof(long long), cudaMemcpyDeviceToHost));
    return 0;
}

// debugging aid (B2_TC_ROLE_CLOCKS=1): cycles the fused launches spent in their two roles since the engine was built:
// host_out[0..2] = {sum, count, max} over state-machine warps that had a chain, [3..5] the same over likelihood CTAs
extern "C" int b2_debug_tc_role_clocks(b2_engine* e, long long* host_out) {
    if (!e || !e->glm_tc) return -1;
    TcHostState* hs = (TcHostState*)e->glm_tc;
    if (!hs->full.role_clk) return -2;
    B2_CUDA_OK(cudaDeviceSynchronize());
    B2_CUDA_OK(cudaMemcpy(host_out, hs->full.role_clk, 6 * sizeof(long long), cudaMemcpyDeviceToHost));
    return 0;
}

// clock64 stamps of chain 0 inside k_glm_tc_post, one row of 16 per leapfrog (ring of 4096)
extern "C" int b2_debug_post_timeline(b2_engine* e, long long* host_out) {
    if (!e || !e->glm_tc) return -1;
    TcWorkspace& w = ((TcHostState*)e->glm_tc)->full;
    if (!w.dbg) return -2;
    B2_CUDA_OK(cudaDeviceSynchronize());
    B2_CUDA_OK(cudaMemcpy(host_out, w.dbg + 48 * TC_DBG_TILES, 4096 * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    return 0;
}

// likelihood only (b2_logp_dlogp, the stepwise API): every chain in one workspace, then the finalize kernel
int b2_glm_tc_launch(b2_engine* e, const float* qA, const float* qB, float* gA, float* gB, int ld,
                     const B2ChainState* st, int n, double* logp, cudaStream_t stream) {
    int rc = tc_ensure(e, stream);
    if (rc) return rc;
    TcHostState* hs = (TcHostState*)e->glm_tc;
    TcWorkspace w = hs->full;                          // geometry of all C chains; n <= C points use its first tiles
    if ((rc = tc_refresh_reference(e, w, qA, qB, ld, st, n, stream))) return rc;
    if ((rc = tc_compact(e, w, qA, qB, ld, st, n, stream))) return rc;
    w.count = n;
    TcWorkspace none;
    memset(&none, 0, sizeof(none));
    B2View<float> v;
    memset(&v, 0, sizeof(v));
    if (env_int("B2_TC_HOOK_FUSED", 0)) rc = tc_launch<true>(e, hs, w, none, v, stream);     // timing aid: the fused build, no state-machine work
    else rc = tc_launch<false>(e, hs, w, none, v, stream);
    if (rc) return rc;
    k_glm_tc_finalize<<<(n + 3) / 4, 128, 0, stream>>>(w, n, e->md.G + 1, e->md.hp[0], qA, qB, gA, gB, ld, st, logp);
    B2_CUDA_OK(cudaGetLastError());
    e->launches += 1;
    return 0;
}
