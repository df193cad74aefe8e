// libb200nuts.so -- engine lifecycle, kernels that wrap the shared state machine, C ABI.
//
// Execution modes
//   persistent : one launch runs every chain's whole run; a warp (D <= B2_WARP_MAX_D) or a
//                block (larger D) owns a chain and loops {model logp+grad, advance}.  Used
//                when a chain's likelihood is cheap enough for one group (eight schools,
//                stochastic volatility, small-N GLM / hierarchical).
//   lock-step  : all chains advance one leapfrog per step: {chain-batched likelihood kernel,
//                advance kernel}; chains are asynchronous across *transitions* (a chain that
//                ends its tree starts the next one in the same step), so the batch stays full.
//
// Reference seams replaced: see include/b200nuts.h.
#include <cstdio>
#include <cstring>
#include <vector>
#include "b2_engine.cuh"

static thread_local std::string g_last_error;
void b2_set_error(const std::string& msg) { g_last_error = msg; }

#define B2_WARP_MAX_D 1024
#define B2_BLOCK_NT 256
#define B2_WARPS_PER_BLOCK 4

// ------------------------------------------------------------------------------- kernels
template <typename T>
__global__ void k_init_chains(B2View<T> w, const T* q0, const unsigned long long* seeds, double step0,
                              const double* mass_mean, const double* mass_var, double mass_weight,
                              int window, int iter0) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= w.C) return;
    B2WarpGroup g;
    B2ChainState s;
    b2_init_chain<T, B2WarpGroup>(g, w, c, s, q0 + (size_t)c * w.D, seeds[c], step0, mass_mean, mass_var,
                                  mass_weight, window, iter0);
    if (g.lane() == 0) w.st[c] = s;
}

// group-per-chain likelihood (parity hook, small models in lock-step mode, cross-check)
template <typename T, typename G>
__device__ __forceinline__ void logp_group_body(const G& g, const B2ModelData& m, const T* qA, const T* qB,
                                                T* gA, T* gB, int ld, const B2ChainState* st, int c,
                                                double* logp) {
    int sel = 0;
    if (st) {
        if (!b2_needs_grad(st[c].phase)) return;
        sel = st[c].sel;
    }
    const T* q = (sel ? qB : qA) + (size_t)c * ld;
    T* gr = (sel ? gB : gA) + (size_t)c * ld;
    const double lp = b2_eval_model<T, G>(g, m, q, gr, c);
    if (g.lane() == 0) logp[c] = lp;
}

template <typename T>
__global__ void k_logp_warp(B2ModelData m, const T* qA, const T* qB, T* gA, T* gB, int ld,
                            const B2ChainState* st, int n, double* logp) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    B2WarpGroup g;
    logp_group_body<T, B2WarpGroup>(g, m, qA, qB, gA, gB, ld, st, c, logp);
}

template <typename T>
__global__ void __launch_bounds__(B2_BLOCK_NT) k_logp_block(B2ModelData m, const T* qA, const T* qB, T* gA, T* gB,
                                                            int ld, const B2ChainState* st, int n, double* logp) {
    __shared__ double red[2 * 8 * (B2_BLOCK_NT / 32)];
    B2BlockGroup<B2_BLOCK_NT> g;
    g.red = red; g.flip = 0;
    logp_group_body<T, B2BlockGroup<B2_BLOCK_NT>>(g, m, qA, qB, gA, gB, ld, st, blockIdx.x, logp);
}

// resume_only: the launch that re-activates finished chains at the start of a continuation run
template <typename T>
__global__ void k_advance_warp(B2View<T> w, int resume_only) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= w.C) return;
    B2WarpGroup g;
    B2ChainState s = w.st[c];
    if (resume_only) {
        if (s.phase != B2_PHASE_DONE) return;
        s.phase = B2_PHASE_RESUME;
    } else if (!b2_needs_grad(s.phase)) return;
    b2_advance<T, B2WarpGroup>(g, w, c, s, w.logp_eval[c]);
    if (g.lane() == 0) w.st[c] = s;
}

template <typename T>
__global__ void __launch_bounds__(B2_BLOCK_NT) k_advance_block(B2View<T> w, int resume_only) {
    __shared__ double red[2 * 8 * (B2_BLOCK_NT / 32)];
    B2BlockGroup<B2_BLOCK_NT> g;
    g.red = red; g.flip = 0;
    const int c = blockIdx.x;
    B2ChainState s = w.st[c];
    if (resume_only) {
        if (s.phase != B2_PHASE_DONE) return;
        s.phase = B2_PHASE_RESUME;
    } else if (!b2_needs_grad(s.phase)) return;
    b2_advance<T, B2BlockGroup<B2_BLOCK_NT>>(g, w, c, s, w.logp_eval[c]);
    if (g.lane() == 0) w.st[c] = s;
}

// hot != null: shared memory for this chain's first w.hot_slots vector slots (B2View::hot), loaded once,
// written back when the chain's run ends -- the persistent kernel then touches HBM/L2 only for the stack
// buffers, the Welford windows, the trace and (hot_slots == B2_V_EVERY_LEAPFROG) the once-per-doubling slots.
// lvh: shared memory for the chain's per-level scalars (log weights, energy, logp of the stack buffers).
template <typename T, typename G>
__device__ __forceinline__ void persistent_body(const G& g, B2View<T>& w, const B2ModelData& m, int c, T* hot, double* lvh,
                                                T* data_chip) {
    B2ChainState s = w.st[c];
    if (s.phase == B2_PHASE_FAILED || (s.phase == B2_PHASE_DONE && s.iter >= w.iter_cap)) return;
    {
        const double* src = w.lv + (size_t)c * 4 * B2_MAX_LEVELS;
        for (int i = g.lane(); i < 4 * B2_MAX_LEVELS; i += G::NT) lvh[i] = src[i];
        g.sync();
        w.lv_hot = lvh;
    }
    if (hot) {
        for (int slot = 0; slot < w.hot_slots; ++slot) {
            const T* src = w.Vglobal(slot, c);
            T* dst = hot + (size_t)slot * w.Dp;
            for (int i = g.lane(); i < w.Dp; i += G::NT) dst[i] = src[i];
        }
        g.sync();
        w.hot = hot;
    }
    if (s.phase == B2_PHASE_DONE && s.iter < w.iter_cap) {   // continuation run
        s.phase = B2_PHASE_RESUME;
        b2_advance<T, G>(g, w, c, s, 0.0);
    }
    B2ModelData mc = m;
    if (data_chip) {                                        // the data vector in the vector dtype, next to the hot slots
        for (int i = g.lane(); i < m.N; i += G::NT) data_chip[i] = (T)m.aux0[i];
        g.sync();
        mc.aux0_chip = data_chip;
    }
    bool active = b2_needs_grad(s.phase);
    while (active) {
        const T* q = w.V(B2_V_QE0 + s.sel, c);
        T* gr = w.V(B2_V_GE0 + s.sel, c);
        g.sync();
        const double lp = b2_eval_model<T, G>(g, mc, q, gr, c);
        g.sync();
        active = b2_advance<T, G>(g, w, c, s, lp);
    }
    if (hot) {
        g.sync();
        for (int slot = 0; slot < w.hot_slots; ++slot) {
            T* dst = w.Vglobal(slot, c);
            const T* src = hot + (size_t)slot * w.Dp;
            for (int i = g.lane(); i < w.Dp; i += G::NT) dst[i] = src[i];
        }
    }
    {
        g.sync();
        double* dst = w.lv + (size_t)c * 4 * B2_MAX_LEVELS;
        for (int i = g.lane(); i < 4 * B2_MAX_LEVELS; i += G::NT) dst[i] = lvh[i];
    }
    if (g.lane() == 0) w.st[c] = s;
}

template <typename T>
__global__ void k_persistent_warp(B2View<T> w, B2ModelData m, int hot_elems) {
    extern __shared__ __align__(16) unsigned char hot_raw[];
    __shared__ double lvh[B2_WARPS_PER_BLOCK][4 * B2_MAX_LEVELS];
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= w.C) return;
    B2WarpGroup g;
    persistent_body<T, B2WarpGroup>(g, w, m, c, hot_elems ? reinterpret_cast<T*>(hot_raw) + (size_t)(threadIdx.x >> 5) * hot_elems : (T*)0,
                                    lvh[threadIdx.x >> 5], (T*)0);
}

// NT threads own one chain.  At D ~ 3000 all 11 hot slots take 128 KB of shared memory (fp32): one block per SM,
// which then has to bring enough warps itself (NT = 512).  With only the 7 every-leapfrog slots on chip (82 KB) two
// blocks share an SM (CTAS = 2): the barriers and L2 round trips of one chain's step are filled by the other chain.
template <typename T, int NT, int CTAS>
__global__ void __launch_bounds__(NT, CTAS) k_persistent_block(B2View<T> w, B2ModelData m, int hot_elems, int data_chip) {
    extern __shared__ __align__(16) unsigned char hot_raw[];
    __shared__ double red[2 * 8 * (NT / 32)];
    __shared__ double lvh[4 * B2_MAX_LEVELS];
    B2BlockGroup<NT> g;
    g.red = red; g.flip = 0;
    persistent_body<T, B2BlockGroup<NT>>(g, w, m, blockIdx.x, hot_elems ? reinterpret_cast<T*>(hot_raw) : (T*)0, lvh,
                                         data_chip ? reinterpret_cast<T*>(hot_raw) + hot_elems : (T*)0);
}

template <typename T, int NT, int CTAS>
static int launch_persistent_block(b2_engine* e, const B2View<T>& w, int hot_elems, size_t hot_bytes, int data_chip, cudaStream_t s) {
    if (hot_bytes > 48 * 1024)
        B2_CUDA_OK(cudaFuncSetAttribute(k_persistent_block<T, NT, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes));
    k_persistent_block<T, NT, CTAS><<<e->C, NT, hot_bytes, s>>>(w, e->md, hot_elems, data_chip);
    return 0;
}

static int env_choice(const char* name, int dflt, int a, int b, int c) {
    const char* env = getenv(name);
    if (env) { const int v = atoi(env); if (v == a || v == b || v == c) return v; }
    return dflt;
}

// chains that still owe transitions of this call (with run-ahead, faster chains are already past iter_end)
__global__ void k_count_active(const B2ChainState* st, int C, int iter_end, int* out) {
    int n = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x)
        n += (b2_needs_grad(st[c].phase) && st[c].iter < iter_end) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, n);
}

// ------------------------------------------------------------------------------ helpers
static size_t dsize(int dtype) { return dtype == B2_F64 ? 8 : 4; }

template <typename T>
static B2View<T> make_view(b2_engine* e) {
    B2View<T> w;
    memset(&w, 0, sizeof(w));
    w.C = e->C; w.D = e->D; w.Dp = e->Dp;
    w.vec = (T*)e->vec; w.wv_mean = e->wv_mean; w.wv_m2 = e->wv_m2; w.st = e->st; w.lv = e->lv; w.logp_eval = e->logp_eval;
    w.hot_slots = B2_V_STACK0;
    return w;
}

static bool use_block_group(const b2_engine* e) { return e->D > B2_WARP_MAX_D; }

static int ensure_glm_scratch(b2_engine* e) {
    if (e->md.family != B2_FAMILY_GLM_LOGIT || e->glm_scratch) return 0;
    const size_t bytes = (size_t)e->C * e->md.N * dsize(e->dtype);
    if (bytes > ((size_t)8 << 30)) {
        b2_set_error("group GLM evaluator needs C*N scratch > 8 GiB; use the SIMT/tcgen05 path");
        return -5;
    }
    B2_CUDA_OK(cudaMalloc(&e->glm_scratch, bytes));
    e->md.scratch = e->glm_scratch;
    return 0;
}

// group evaluator for n points (planes A/B of row stride ld)
template <typename T>
static int launch_logp_group(b2_engine* e, const T* qA, const T* qB, T* gA, T* gB, int ld,
                             const B2ChainState* st, int n, double* logp, cudaStream_t s) {
    int rc = ensure_glm_scratch(e);
    if (rc) return rc;
    if (use_block_group(e)) {
        k_logp_block<T><<<n, B2_BLOCK_NT, 0, s>>>(e->md, qA, qB, gA, gB, ld, st, n, logp);
    } else {
        const int nb = (n + B2_WARPS_PER_BLOCK - 1) / B2_WARPS_PER_BLOCK;
        k_logp_warp<T><<<nb, 32 * B2_WARPS_PER_BLOCK, 0, s>>>(e->md, qA, qB, gA, gB, ld, st, n, logp);
    }
    e->launches += 1;
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

static int pick_glm_path(const b2_engine* e, int requested) {
    if (e->md.family != B2_FAMILY_GLM_LOGIT) return B2_GLM_GROUP;
    if (requested == B2_GLM_AUTO) {
        if (e->dtype == B2_F32 && (b2_glm_tc_supported(e) || b2_glm_tcw_supported(e))) return B2_GLM_TCGEN05;
        return ((size_t)e->md.N * e->C >= (size_t)1 << 16) ? B2_GLM_SIMT : B2_GLM_GROUP;
    }
    return requested;
}

// one likelihood evaluation for every chain, dispatching to the best kernel for the family
template <typename T>
static int launch_likelihood(b2_engine* e, const T* qA, const T* qB, T* gA, T* gB, int ld,
                             const B2ChainState* st, int n, double* logp, int glm_path, cudaStream_t s);

template <>
int launch_likelihood<float>(b2_engine* e, const float* qA, const float* qB, float* gA, float* gB, int ld,
                             const B2ChainState* st, int n, double* logp, int glm_path, cudaStream_t s) {
    if (e->md.family == B2_FAMILY_GLM_LOGIT) {
        const int path = pick_glm_path(e, glm_path);
        if (path == B2_GLM_TCGEN05) {
            if (b2_glm_tcw_supported(e)) return b2_glm_tcw_launch(e, qA, qB, gA, gB, ld, st, n, logp, s);
            if (!b2_glm_tc_supported(e)) { b2_set_error("tcgen05 GLM path does not support this shape"); return -6; }
            return b2_glm_tc_launch(e, qA, qB, gA, gB, ld, st, n, logp, s);
        }
        if (path == B2_GLM_SIMT) return b2_glm_simt_launch<float>(e, qA, qB, gA, gB, ld, st, n, logp, s);
    }
    if (e->md.family == B2_FAMILY_HIER_LINEAR_NCP && (size_t)e->md.N * n >= (size_t)1 << 22)
        return b2_hier_launch<float>(e, qA, qB, gA, gB, ld, st, n, logp, s);
    return launch_logp_group<float>(e, qA, qB, gA, gB, ld, st, n, logp, s);
}

template <>
int launch_likelihood<double>(b2_engine* e, const double* qA, const double* qB, double* gA, double* gB, int ld,
                              const B2ChainState* st, int n, double* logp, int glm_path, cudaStream_t s) {
    if (e->md.family == B2_FAMILY_GLM_LOGIT) {
        const int path = pick_glm_path(e, glm_path);
        if (path == B2_GLM_TCGEN05) { b2_set_error("tcgen05 GLM path is fp32-only (use B2_F32)"); return -6; }
        if (path == B2_GLM_SIMT) return b2_glm_simt_launch<double>(e, qA, qB, gA, gB, ld, st, n, logp, s);
    }
    if (e->md.family == B2_FAMILY_HIER_LINEAR_NCP && (size_t)e->md.N * n >= (size_t)1 << 22)
        return b2_hier_launch<double>(e, qA, qB, gA, gB, ld, st, n, logp, s);
    return launch_logp_group<double>(e, qA, qB, gA, gB, ld, st, n, logp, s);
}

// ------------------------------------------------------------------------------- C ABI
extern "C" int b2_abi_version(void) { return B2_ABI_VERSION; }
extern "C" const char* b2_last_error(void) { return g_last_error.c_str(); }

extern "C" int b2_engine_create(const b2_model_desc* desc, int32_t n_chains, int32_t dtype, int32_t device,
                                b2_engine** out) {
    if (!desc || !out || n_chains <= 0) { b2_set_error("b2_engine_create: bad arguments"); return -1; }
    if (dtype != B2_F32 && dtype != B2_F64) { b2_set_error("b2_engine_create: dtype must be B2_F32 or B2_F64"); return -2; }
    int expect_D = -1;
    switch (desc->family) {
    case B2_STD_NORMAL: expect_D = desc->D; break;
    case B2_EIGHT_SCHOOLS_NCP: expect_D = desc->N + 2; break;
    case B2_GLM_LOGIT: expect_D = desc->G + 1; break;
    case B2_HIER_LINEAR_NCP: expect_D = 2 * desc->G + 5; break;
    case B2_STOCH_VOL: expect_D = desc->N + 2; break;
    default: b2_set_error("b2_engine_create: unknown model family"); return -3;
    }
    if (desc->D != expect_D || desc->D <= 0) { b2_set_error("b2_engine_create: D inconsistent with family/shape"); return -4; }
    int count = 0;
    B2_CUDA_OK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) { b2_set_error("b2_engine_create: no such CUDA device"); return -7; }
    B2_CUDA_OK(cudaSetDevice(device));
    b2_engine* e = new b2_engine();
    memset(e, 0, sizeof(*e));
    e->desc = *desc;
    e->C = n_chains; e->D = desc->D; e->Dp = (desc->D + 3) & ~3; e->dtype = dtype; e->device = device;
    e->md.family = desc->family; e->md.D = desc->D; e->md.N = desc->N; e->md.G = desc->G;
    e->md.aux0 = desc->d_aux0; e->md.aux1 = desc->d_aux1; e->md.X = desc->d_X; e->md.yf = desc->d_y;
    e->md.floor_u8 = desc->d_floor; e->md.grp_off = desc->d_grp_off; e->md.scratch = nullptr; e->md.aux0_chip = nullptr;
    for (int i = 0; i < 4; ++i) e->md.hp[i] = desc->hp[i];
    cudaDeviceProp prop;
    B2_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    e->sm_count = prop.multiProcessorCount;
    const size_t nvec = (size_t)B2_NUM_VEC_SLOTS * e->C * e->Dp;
    B2_CUDA_OK(cudaMalloc(&e->vec, nvec * dsize(dtype)));
    B2_CUDA_OK(cudaMemset(e->vec, 0, nvec * dsize(dtype)));
    B2_CUDA_OK(cudaMalloc(&e->wv_mean, (size_t)2 * e->C * e->Dp * sizeof(double)));
    B2_CUDA_OK(cudaMalloc(&e->wv_m2, (size_t)2 * e->C * e->Dp * sizeof(double)));
    B2_CUDA_OK(cudaMalloc(&e->st, (size_t)e->C * sizeof(B2ChainState)));
    B2_CUDA_OK(cudaMemset(e->st, 0, (size_t)e->C * sizeof(B2ChainState)));
    B2_CUDA_OK(cudaMalloc(&e->lv, (size_t)e->C * 4 * B2_MAX_LEVELS * sizeof(double)));
    B2_CUDA_OK(cudaMemset(e->lv, 0, (size_t)e->C * 4 * B2_MAX_LEVELS * sizeof(double)));
    B2_CUDA_OK(cudaMalloc(&e->logp_eval, (size_t)e->C * sizeof(double)));
    B2_CUDA_OK(cudaMalloc(&e->d_active, sizeof(int)));
    B2_CUDA_OK(cudaMallocHost(&e->h_active, sizeof(int)));
    *out = e;
    return 0;
}

extern "C" int b2_engine_destroy(b2_engine* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaFree(e->vec); cudaFree(e->wv_mean); cudaFree(e->wv_m2); cudaFree(e->st); cudaFree(e->lv); cudaFree(e->logp_eval);
    b2_glm_tc_release(e);
    b2_glm_tcw_release(e);
    cudaFree(e->glm_scratch); cudaFree(e->d_active); cudaFree(e->glm_ws); cudaFree(e->hier_ws);
    cudaFree(e->dense_L); cudaFree(e->dense_Lt); cudaFree(e->dense_q); cudaFree(e->dense_g);
    cudaFreeHost(e->h_active);
    if (e->own_stream) { cudaStreamDestroy(e->own_stream); cudaEventDestroy(e->own_event); }
    if (e->ev[0]) for (int i = 0; i < 96; ++i) cudaEventDestroy(e->ev[i]);
    delete e;
    return 0;
}

extern "C" int b2_logp_dlogp(b2_engine* e, const void* d_q, int32_t n_points, double* d_logp, void* d_grad,
                             int32_t glm_path, void* stream) {
    if (!e || !d_q || !d_logp || !d_grad) { b2_set_error("b2_logp_dlogp: null argument"); return -1; }
    if (n_points <= 0 || n_points > e->C) { b2_set_error("b2_logp_dlogp: n_points must be in [1, n_chains]"); return -2; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (e->dtype == B2_F64)
        return launch_likelihood<double>(e, (const double*)d_q, (const double*)d_q, (double*)d_grad, (double*)d_grad,
                                         e->D, nullptr, n_points, d_logp, glm_path, s);
    return launch_likelihood<float>(e, (const float*)d_q, (const float*)d_q, (float*)d_grad, (float*)d_grad,
                                    e->D, nullptr, n_points, d_logp, glm_path, s);
}

template <typename T>
static int set_state_t(b2_engine* e, const void* d_q0, const uint64_t* d_seeds, double step0,
                       const double* mm, const double* mv, double mw, int window, cudaStream_t s) {
    B2View<T> w = make_view<T>(e);
    const int nb = (e->C + B2_WARPS_PER_BLOCK - 1) / B2_WARPS_PER_BLOCK;
    k_init_chains<T><<<nb, 32 * B2_WARPS_PER_BLOCK, 0, s>>>(w, (const T*)d_q0, (const unsigned long long*)d_seeds,
                                                            step0, mm, mv, mw, window, 0);
    e->launches += 1;
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int b2_set_state(b2_engine* e, const void* d_q0, const uint64_t* d_seeds, double step_size0,
                            const double* d_mass_mean, const double* d_mass_var, double mass_weight,
                            int32_t adaptation_window, void* stream) {
    if (!e || !d_q0 || !d_seeds || !d_mass_mean || !d_mass_var) { b2_set_error("b2_set_state: null argument"); return -1; }
    if (!(step_size0 > 0) || adaptation_window <= 0) { b2_set_error("b2_set_state: step size and window must be > 0"); return -2; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    int rc = e->dtype == B2_F64
                 ? set_state_t<double>(e, d_q0, d_seeds, step_size0, d_mass_mean, d_mass_var, mass_weight, adaptation_window, (cudaStream_t)stream)
                 : set_state_t<float>(e, d_q0, d_seeds, step_size0, d_mass_mean, d_mass_var, mass_weight, adaptation_window, (cudaStream_t)stream);
    if (rc) return rc;
    e->iter_done = 0;
    e->state_set = true;
    return 0;
}

template <typename T>
__global__ void k_set_position(B2View<T> w, const T* q) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)w.C * w.D) return;
    const int c = (int)(idx / w.D), i = (int)(idx - (size_t)c * w.D);
    w.V(B2_V_PROPQ, c)[i] = q[idx];
    w.V(B2_V_QE1, c)[i] = q[idx];
    if (i == 0 && w.st[c].phase != B2_PHASE_FAILED) { w.st[c].phase = B2_PHASE_INIT; w.st[c].sel = 1; }
}

extern "C" int b2_set_position(b2_engine* e, const void* d_q, void* stream) {
    if (!e || !d_q) { b2_set_error("b2_set_position: null argument"); return -1; }
    if (!e->state_set) { b2_set_error("b2_set_position: call b2_set_state first"); return -2; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    const size_t n = (size_t)e->C * e->D;
    const int nb = (int)((n + 255) / 256);
    if (e->dtype == B2_F64) k_set_position<double><<<nb, 256, 0, (cudaStream_t)stream>>>(make_view<double>(e), (const double*)d_q);
    else k_set_position<float><<<nb, 256, 0, (cudaStream_t)stream>>>(make_view<float>(e), (const float*)d_q);
    e->launches += 1;
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ dense mass matrix (reparameterised run)
// One warp per chain.  fwd: q = L z for the chain's pending position (the edge st[c].sel points at);
// bwd: grad_z = L^T grad_q written to that edge's gradient slot.  Lt / L are read with consecutive lanes on
// consecutive addresses; the chain's vector is staged in shared memory.  D x D multiply-adds per chain and pass:
// noise next to a likelihood over N observations.
template <typename T>
__global__ void k_dense_fwd(const T* __restrict__ Lt, int D, int Dp, const T* zA, const T* zB, const B2ChainState* st, int C, T* q_out) {
    extern __shared__ __align__(16) unsigned char dense_raw[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + wib;
    if (c >= C || !b2_needs_grad(st[c].phase)) return;
    T* z = reinterpret_cast<T*>(dense_raw) + (size_t)wib * Dp;
    const T* src = (st[c].sel ? zB : zA) + (size_t)c * Dp;
    for (int i = lane; i < D; i += 32) z[i] = src[i];
    __syncwarp();
    for (int i = lane; i < D; i += 32) {
        T acc = (T)0;
        for (int j = 0; j <= i; ++j) acc += Lt[(size_t)j * D + i] * z[j];      // L[i][j], j <= i
        q_out[(size_t)c * Dp + i] = acc;
    }
}

template <typename T>
__global__ void k_dense_bwd(const T* __restrict__ L, int D, int Dp, const T* g_in, T* gA, T* gB, const B2ChainState* st, int C) {
    extern __shared__ __align__(16) unsigned char dense_raw[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + wib;
    if (c >= C || !b2_needs_grad(st[c].phase)) return;
    T* g = reinterpret_cast<T*>(dense_raw) + (size_t)wib * Dp;
    for (int i = lane; i < D; i += 32) g[i] = g_in[(size_t)c * Dp + i];
    __syncwarp();
    T* dst = (st[c].sel ? gB : gA) + (size_t)c * Dp;
    for (int j = lane; j < D; j += 32) {
        T acc = (T)0;
        for (int i = j; i < D; ++i) acc += L[(size_t)i * D + j] * g[i];        // (L^T g)_j = sum_{i >= j} L[i][j] g[i]
        dst[j] = acc;
    }
}

template <typename T>
__global__ void k_dense_prepare(const T* chol, int D, T* L, T* Lt) {       // keeps the lower triangle only
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D * D) return;
    const int i = idx / D, j = idx % D;
    const T v = j <= i ? chol[idx] : (T)0;
    L[idx] = v;
    Lt[(size_t)j * D + i] = v;
}

extern "C" int b2_set_dense_mass(b2_engine* e, const void* d_chol, void* stream) {
    if (!e) { b2_set_error("b2_set_dense_mass: null engine"); return -1; }
    if (e->stepping) { b2_set_error("b2_set_dense_mass: a stepwise run is in progress"); return -6; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    if (!d_chol) {
        cudaFree(e->dense_L); cudaFree(e->dense_Lt); cudaFree(e->dense_q); cudaFree(e->dense_g);
        e->dense_L = e->dense_Lt = e->dense_q = e->dense_g = nullptr;
        return 0;
    }
    if (e->D > 1024) { b2_set_error("b2_set_dense_mass: D > 1024 (a D x D factor per leapfrog and chain) is not supported"); return -5; }
    const size_t el = e->dtype == B2_F64 ? 8 : 4;
    if (!e->dense_L) {
        B2_CUDA_OK(cudaMalloc(&e->dense_L, (size_t)e->D * e->D * el));
        B2_CUDA_OK(cudaMalloc(&e->dense_Lt, (size_t)e->D * e->D * el));
        B2_CUDA_OK(cudaMalloc(&e->dense_q, (size_t)e->C * e->Dp * el));
        B2_CUDA_OK(cudaMalloc(&e->dense_g, (size_t)e->C * e->Dp * el));
        B2_CUDA_OK(cudaMemsetAsync(e->dense_q, 0, (size_t)e->C * e->Dp * el, (cudaStream_t)stream));
        B2_CUDA_OK(cudaMemsetAsync(e->dense_g, 0, (size_t)e->C * e->Dp * el, (cudaStream_t)stream));
    }
    const int n = e->D * e->D;
    if (e->dtype == B2_F64) k_dense_prepare<double><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const double*)d_chol, e->D, (double*)e->dense_L, (double*)e->dense_Lt);
    else k_dense_prepare<float><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)d_chol, e->D, (float*)e->dense_L, (float*)e->dense_Lt);
    e->launches += 1;
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

static bool persistent_ok(const b2_engine* e) {
    switch (e->md.family) {
    case B2_FAMILY_STD_NORMAL:
    case B2_FAMILY_EIGHT_SCHOOLS_NCP:
    case B2_FAMILY_STOCH_VOL:
        return true;
    case B2_FAMILY_GLM_LOGIT:
        return (size_t)e->md.N * (e->md.G + 1) <= (size_t)1 << 16;     // one warp can afford the whole likelihood
    case B2_FAMILY_HIER_LINEAR_NCP:
        return e->md.N <= 1 << 14;
    }
    return false;
}

template <typename T>
static B2View<T> build_view(b2_engine* e, const b2_sampler_opts* o, const b2_trace_out* tr) {
    B2View<T> w = make_view<T>(e);
    w.kind = o->kind;
    w.iter_base = e->iter_done; w.iter_end = e->iter_done + o->n_iters; w.tune_until = o->tune_until;
    w.iter_cap = w.iter_end;                      // run_t raises it for lock-step runs with run_ahead
    w.max_treedepth = o->max_treedepth; w.early_max_treedepth = o->early_max_treedepth;
    w.emax = o->Emax; w.target = o->target_accept; w.gamma = o->gamma; w.k = o->k; w.t0 = o->t0;
    w.adapt_step = o->adapt_step_size; w.adapt_mass = o->adapt_mass;
    w.path_length = o->path_length; w.max_steps = o->max_steps; w.hmc_jitter = o->hmc_jitter;
    if (tr) {
        w.tr_q = (T*)tr->d_q; w.tr_energy = tr->d_energy; w.tr_energy_error = tr->d_energy_error;
        w.tr_max_energy_error = tr->d_max_energy_error; w.tr_mean_tree_accept = tr->d_mean_tree_accept;
        w.tr_step_size = tr->d_step_size; w.tr_step_size_bar = tr->d_step_size_bar; w.tr_model_logp = tr->d_model_logp;
        w.tr_accept = tr->d_accept; w.tr_depth = tr->d_depth; w.tr_tree_size = tr->d_tree_size; w.tr_n_steps = tr->d_n_steps;
        w.tr_diverging = tr->d_diverging; w.tr_tune = tr->d_tune; w.tr_accepted = tr->d_accepted;
    }
    return w;
}

template <typename T>
static int run_t(b2_engine* e, const b2_sampler_opts* o, const b2_trace_out* tr, cudaStream_t s_in) {
    // The run is host-synchronous (it returns when every chain has done its transitions), so it may run on a
    // stream of its own: the caller's stream is usually the legacy default stream, which cannot be captured into a
    // CUDA graph.  Work already queued on the caller's stream is ordered before the run through an event.
    cudaStream_t s = s_in;
    if (s_in == (cudaStream_t)0 || s_in == cudaStreamLegacy || s_in == cudaStreamPerThread) {
        if (!e->own_stream) {
            B2_CUDA_OK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
            B2_CUDA_OK(cudaEventCreateWithFlags(&e->own_event, cudaEventDisableTiming));
        }
        B2_CUDA_OK(cudaEventRecord(e->own_event, s_in));
        B2_CUDA_OK(cudaStreamWaitEvent(e->own_stream, e->own_event, 0));
        s = e->own_stream;
    }
    B2View<T> w = build_view<T>(e, o, tr);
    int mode = o->exec_mode;
    if (mode == B2_EXEC_AUTO) mode = persistent_ok(e) ? B2_EXEC_PERSISTENT : B2_EXEC_LOCKSTEP;
    const bool dense = e->dense_L != nullptr;
    if (dense) {
        if (o->adapt_mass) { b2_set_error("b2_sample_run: a dense mass matrix is static (adapt_mass must be 0)"); return -5; }
        mode = B2_EXEC_LOCKSTEP;                      // the reparameterisation wraps the chain-batched likelihood launch
    }
    if (mode == B2_EXEC_PERSISTENT && e->md.family == B2_FAMILY_GLM_LOGIT) {
        int rc = ensure_glm_scratch(e);
        if (rc) return rc;
    }
    const bool blk = use_block_group(e);
    const int nb_warp = (e->C + B2_WARPS_PER_BLOCK - 1) / B2_WARPS_PER_BLOCK;
    if (mode == B2_EXEC_PERSISTENT) {
        // shared-memory residency of each chain's hot vector slots when it fits (11 * Dp elements per chain)
        // Block per chain: two blocks per SM when two sets of the 7 every-leapfrog slots fit next to each other
        // (B2_PBLOCK_CTAS / _NT / _HOT override the choice for A/B runs).
        const size_t slot_bytes = (size_t)e->Dp * sizeof(T);
        const size_t smem_sm = (size_t)227 * 1024;
        int ctas = 1, nt = e->D >= 2048 ? 512 : 256;
        if (blk) {
            // the lean layout (7 slots) pays when all 11 would not leave room for a second block; with no more chains
            // than SMs a block has its SM to itself and takes 512 threads instead (+6 % at 64 / 148 chains)
            const bool lean = 2 * (B2_V_EVERY_LEAPFROG * slot_bytes + 4096) <= smem_sm && 2 * (B2_V_STACK0 * slot_bytes + 4096) > smem_sm;
            if (lean && e->C > e->sm_count) { ctas = 2; nt = 256; }
            ctas = env_choice("B2_PBLOCK_CTAS", ctas, 1, 2, 2);
            nt = env_choice("B2_PBLOCK_NT", nt, 256, 512, 1024);
            if (ctas == 2 && nt == 1024) nt = 512;
            w.hot_slots = env_choice("B2_PBLOCK_HOT", (lean || ctas == 2) ? B2_V_EVERY_LEAPFROG : B2_V_STACK0, B2_V_EVERY_LEAPFROG, B2_V_STACK0, B2_V_STACK0);
        }
        int hot_elems = w.hot_slots * e->Dp;
        size_t hot_bytes = (size_t)hot_elems * sizeof(T) * (blk ? 1 : B2_WARPS_PER_BLOCK);
        if (hot_bytes > (size_t)200 * 1024 && blk && w.hot_slots == B2_V_STACK0) {       // fp64 at D ~ 3000: the 7 still fit
            w.hot_slots = B2_V_EVERY_LEAPFROG; hot_elems = w.hot_slots * e->Dp; hot_bytes = (size_t)hot_elems * sizeof(T);
        }
        if (hot_bytes > (size_t)200 * 1024) { hot_elems = 0; hot_bytes = 0; }
        // stochastic volatility: the returns are read once per latent per leapfrog and the tree-stack traffic keeps
        // evicting them from L1 -- a copy in the vector dtype goes next to the hot slots when it fits
        int data_chip = 0;
        if (blk && hot_elems && e->md.family == B2_FAMILY_STOCH_VOL && !(getenv("B2_PBLOCK_DATA") && atoi(getenv("B2_PBLOCK_DATA")) == 0)) {
            const size_t with = hot_bytes + (size_t)e->md.N * sizeof(T);
            if (with <= (size_t)200 * 1024 && (ctas == 1 || 2 * (with + 4096) <= smem_sm)) { data_chip = 1; hot_bytes = with; }
        }
        if (hot_bytes > 48 * 1024 && !blk)
            B2_CUDA_OK(cudaFuncSetAttribute(k_persistent_warp<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes));
        if (blk) {
            int rc = ctas == 2 ? (nt == 512 ? launch_persistent_block<T, 512, 2>(e, w, hot_elems, hot_bytes, data_chip, s)
                                            : launch_persistent_block<T, 256, 2>(e, w, hot_elems, hot_bytes, data_chip, s))
                   : nt == 1024 ? launch_persistent_block<T, 1024, 1>(e, w, hot_elems, hot_bytes, data_chip, s)
                   : nt == 512  ? launch_persistent_block<T, 512, 1>(e, w, hot_elems, hot_bytes, data_chip, s)
                                : launch_persistent_block<T, 256, 1>(e, w, hot_elems, hot_bytes, data_chip, s);
            if (rc) return rc;
        }
        else k_persistent_warp<T><<<nb_warp, 32 * B2_WARPS_PER_BLOCK, hot_bytes, s>>>(w, e->md, hot_elems);
        e->launches += 1;
        B2_CUDA_OK(cudaGetLastError());
        B2_CUDA_OK(cudaStreamSynchronize(s));
    } else {
        const T* qA = w.V(B2_V_QE0, 0); const T* qB = w.V(B2_V_QE1, 0);
        T* gA = w.V(B2_V_GE0, 0); T* gB = w.V(B2_V_GE1, 0);
        const int batch = 32;
        // fast chains may run ahead of this call's last iteration (rows exist in the caller's trace buffers)
        if (o->run_ahead > 0) w.iter_cap = w.iter_end + o->run_ahead;
        // GLM on the tensor-core path: {tcgen05 likelihood, fused finalize+advance+repack} per step
        const bool fused_tc = sizeof(T) == 4 && !blk && !dense && e->md.family == B2_FAMILY_GLM_LOGIT &&
                              pick_glm_path(e, o->glm_path) == B2_GLM_TCGEN05 && !b2_glm_tcw_supported(e);
        if (fused_tc && !b2_glm_tc_supported(e)) { b2_set_error("tcgen05 GLM path does not support this shape"); return -6; }
        if (e->iter_done > 0) {                       // re-activate chains that finished the previous call
            if (blk) k_advance_block<T><<<e->C, B2_BLOCK_NT, 0, s>>>(w, 1);
            else k_advance_warp<T><<<nb_warp, 32 * B2_WARPS_PER_BLOCK, 0, s>>>(w, 1);
            e->launches += 1;
        }
        if (fused_tc) {
            int rc = b2_glm_tc_pack(e, (const float*)qA, (const float*)qB, e->Dp, e->st, e->C, s);
            if (rc) return rc;
        } else if (sizeof(T) == 4 && b2_glm_tcw_supported(e) && pick_glm_path(e, o->glm_path) == B2_GLM_TCGEN05) {
            const float *rA = (const float*)qA, *rB = (const float*)qB;
            if (dense) {                                  // the likelihood sees q = L z: centre the reference there
                k_dense_fwd<T><<<nb_warp, 32 * B2_WARPS_PER_BLOCK, (size_t)B2_WARPS_PER_BLOCK * e->Dp * sizeof(T), s>>>(
                    (const T*)e->dense_Lt, e->D, e->Dp, qA, qB, e->st, e->C, (T*)e->dense_q);
                e->launches += 1;
                rA = rB = (const float*)e->dense_q;
            }
            int rc = b2_glm_tcw_refresh(e, rA, rB, e->Dp, e->st, e->C, s);   // reference position of this run
            if (rc) return rc;
        }
        // one batch = `batch` leapfrogs of every live chain + the count of chains that still owe transitions
        auto one_batch = [&]() -> int {
            for (int b = 0; b < batch; ++b) {
                if (e->profile) cudaEventRecord(e->ev[3 * b], s);
                if (fused_tc) {
                    // likelihood + state machine (b2_glm_tc.cu); the events split the two in the two-kernel schedule
                    int rc = b2_glm_tc_step(e, &w, s, e->profile ? e->ev[3 * b + 1] : (cudaEvent_t)0);
                    if (rc) return rc;
                } else {
                    int rc;
                    if (dense) {                          // q = L z  ->  logp, grad_q  ->  grad_z = L^T grad_q
                        T *dq = (T*)e->dense_q, *dg = (T*)e->dense_g;
                        const size_t sm = (size_t)B2_WARPS_PER_BLOCK * e->Dp * sizeof(T);
                        k_dense_fwd<T><<<nb_warp, 32 * B2_WARPS_PER_BLOCK, sm, s>>>((const T*)e->dense_Lt, e->D, e->Dp, qA, qB, e->st, e->C, dq);
                        rc = launch_likelihood<T>(e, dq, dq, dg, dg, e->Dp, e->st, e->C, e->logp_eval, o->glm_path, s);
                        if (rc) return rc;
                        k_dense_bwd<T><<<nb_warp, 32 * B2_WARPS_PER_BLOCK, sm, s>>>((const T*)e->dense_L, e->D, e->Dp, dg, gA, gB, e->st, e->C);
                        e->launches += 2;
                    } else {
                        rc = launch_likelihood<T>(e, qA, qB, gA, gB, e->Dp, e->st, e->C, e->logp_eval, o->glm_path, s);
                        if (rc) return rc;
                    }
                    if (e->profile) cudaEventRecord(e->ev[3 * b + 1], s);
                    if (blk) k_advance_block<T><<<e->C, B2_BLOCK_NT, 0, s>>>(w, 0);
                    else k_advance_warp<T><<<nb_warp, 32 * B2_WARPS_PER_BLOCK, 0, s>>>(w, 0);
                    e->launches += 1;
                }
                if (e->profile) cudaEventRecord(e->ev[3 * b + 2], s);
            }
            B2_CUDA_OK(cudaMemsetAsync(e->d_active, 0, sizeof(int), s));
            k_count_active<<<(e->C + 255) / 256 < 64 ? (e->C + 255) / 256 : 64, 256, 0, s>>>(e->st, e->C, w.iter_end, e->d_active);
            e->launches += 1;
            B2_CUDA_OK(cudaMemcpyAsync(e->h_active, e->d_active, sizeof(int), cudaMemcpyDeviceToHost, s));
            return 0;
        };
        // The batch is launch-bound at its seams (2-5 launches per leapfrog, a few us of launch latency each against
        // 20-140 us kernels), so after a first direct batch (lazy workspace allocations happen there) it is captured
        // into a CUDA graph and replayed.  Not on the legacy default stream (not capturable), not while per-launch
        // events are recorded, not for the fused tensor-core schedule (its first launch differs from the rest).
        bool use_graph = !e->profile && (cudaStream_t)s != (cudaStream_t)0 && s != cudaStreamLegacy && s != cudaStreamPerThread &&
                         !(getenv("B2_GRAPH") && atoi(getenv("B2_GRAPH")) == 0) && !(fused_tc && b2_glm_tc_is_fused(e));
        cudaGraphExec_t exec = nullptr;
        long long per_batch = 0;
        for (int it = 0;; ++it) {
            if (use_graph && it >= 1) {
                if (!exec) {
                    const long long l0 = e->launches;
                    cudaGraph_t graph = nullptr;
                    bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                    if (ok) {
                        const int rc = one_batch();
                        ok = (cudaStreamEndCapture(s, &graph) == cudaSuccess) && rc == 0 && graph != nullptr;
                    }
                    if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
                    if (graph) cudaGraphDestroy(graph);
                    per_batch = e->launches - l0;
                    e->launches = l0;
                    if (!ok) { cudaGetLastError(); exec = nullptr; use_graph = false; }
                }
                if (exec) {
                    B2_CUDA_OK(cudaGraphLaunch(exec, s));
                    e->launches += per_batch;
                } else {
                    int rc = one_batch();
                    if (rc) return rc;
                }
            } else {
                int rc = one_batch();
                if (rc) return rc;
            }
            B2_CUDA_OK(cudaStreamSynchronize(s));
            if (e->profile) {
                for (int b = 0; b < batch; ++b) {
                    float ms = 0.f;
                    if (cudaEventElapsedTime(&ms, e->ev[3 * b], e->ev[3 * b + 1]) == cudaSuccess) { e->like_ms += ms; e->like_n += 1; }
                    if (cudaEventElapsedTime(&ms, e->ev[3 * b + 1], e->ev[3 * b + 2]) == cudaSuccess) e->adv_ms += ms;
                }
            }
            if (*e->h_active == 0) break;
        }
        if (exec) cudaGraphExecDestroy(exec);
        B2_CUDA_OK(cudaGetLastError());
    }
    e->iter_done += o->n_iters;
    return 0;
}

// argument checks shared by b2_sample_run and b2_step_begin (the stepwise path runs the same state machine)
static int check_run_args(const char* who, const b2_engine* e, const b2_sampler_opts* o) {
    const std::string w(who);
    if (!e || !o) { b2_set_error(w + ": null argument"); return -1; }
    if (!e->state_set) { b2_set_error(w + ": call b2_set_state first"); return -2; }
    if (o->n_iters <= 0) { b2_set_error(w + ": n_iters must be > 0"); return -3; }
    if (o->kind != B2_NUTS && o->kind != B2_HMC) { b2_set_error(w + ": unknown sampler kind"); return -4; }
    if (o->kind == B2_NUTS && (o->max_treedepth < 1 || o->max_treedepth > B2_MAX_LEVELS ||
                               o->early_max_treedepth < 1 || o->early_max_treedepth > B2_MAX_LEVELS)) {
        b2_set_error(w + ": tree depths must be in [1, 12]");     // slot_map holds 12 nibbles, B2_MAX_LEVELS stack buffers
        return -5;
    }
    if (o->kind == B2_HMC && o->max_steps < 1) { b2_set_error(w + ": max_steps must be >= 1"); return -5; }
    return 0;
}

extern "C" int b2_sample_run(b2_engine* e, const b2_sampler_opts* o, const b2_trace_out* trace, void* stream) {
    const int rc = check_run_args("b2_sample_run", e, o);
    if (rc) return rc;
    if (e->stepping) { b2_set_error("b2_sample_run: a stepwise run is in progress (b2_step_end first)"); return -6; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    return e->dtype == B2_F64 ? run_t<double>(e, o, trace, (cudaStream_t)stream)
                              : run_t<float>(e, o, trace, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ leapfrog integrator hook
// CpuLeapfrogIntegrator.compute_state + n x .step(epsilon, state) (integration.py:39-47, 49-109) for every
// chain at once, with a static diagonal potential `var` (QuadPotentialDiag, quadpotential.py:356-397).  It runs
// the very device functions the samplers use (b2_prepare_leapfrog / b2_finish_leapfrog + the family's
// likelihood kernel), so the reference's reversibility test (tests/test_hmc.py:27-46) and the oracle's
// Leapfrog.step can be held against the product's integrator directly.
enum { B2_LF_LOAD = 0, B2_LF_PREPARE = 1, B2_LF_FINISH = 2, B2_LF_STORE = 3 };

template <typename T, typename G>
__device__ __forceinline__ void lf_body(const G& g, const B2View<T>& w, int c, int op, double eps, const T* q_in,
                                        const T* p_in, const double* var, T* q_out, T* p_out, double* energy) {
    if (op == B2_LF_LOAD) {
        T *q = w.V(B2_V_QE0, c), *p = w.V(B2_V_PE0, c), *vr = w.V(B2_V_VAR, c);
        for (int i = g.lane(); i < w.D; i += G::NT) {
            q[i] = q_in[(size_t)c * w.D + i]; p[i] = p_in[(size_t)c * w.D + i]; vr[i] = (T)var[i];
        }
    } else if (op == B2_LF_PREPARE) {
        b2_prepare_leapfrog<T, G>(g, w, c, 0, eps);
    } else if (op == B2_LF_FINISH) {
        const double en = b2_finish_leapfrog<T, G>(g, w, c, 0, eps, w.logp_eval[c]);
        if (g.lane() == 0 && energy) energy[c] = en;
    } else {
        const T *q = w.V(B2_V_QE0, c), *p = w.V(B2_V_PE0, c);
        for (int i = g.lane(); i < w.D; i += G::NT) { q_out[(size_t)c * w.D + i] = q[i]; p_out[(size_t)c * w.D + i] = p[i]; }
    }
}

template <typename T>
__global__ void k_lf_warp(B2View<T> w, int op, double eps, const T* q_in, const T* p_in, const double* var, T* q_out,
                          T* p_out, double* energy) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= w.C) return;
    B2WarpGroup g;
    lf_body<T, B2WarpGroup>(g, w, c, op, eps, q_in, p_in, var, q_out, p_out, energy);
}

template <typename T>
__global__ void __launch_bounds__(B2_BLOCK_NT) k_lf_block(B2View<T> w, int op, double eps, const T* q_in, const T* p_in,
                                                          const double* var, T* q_out, T* p_out, double* energy) {
    __shared__ double red[2 * 8 * (B2_BLOCK_NT / 32)];
    B2BlockGroup<B2_BLOCK_NT> g;
    g.red = red; g.flip = 0;
    lf_body<T, B2BlockGroup<B2_BLOCK_NT>>(g, w, blockIdx.x, op, eps, q_in, p_in, var, q_out, p_out, energy);
}

template <typename T>
static int leapfrog_t(b2_engine* e, const void* d_q, const void* d_p, const double* d_var, double eps, int n_steps,
                      void* d_q_out, void* d_p_out, double* d_energy, int glm_path, cudaStream_t s) {
    B2View<T> w = make_view<T>(e);
    const bool blk = use_block_group(e);
    const int nb = (e->C + B2_WARPS_PER_BLOCK - 1) / B2_WARPS_PER_BLOCK;
    auto phase = [&](int op, double ee) {
        if (blk) k_lf_block<T><<<e->C, B2_BLOCK_NT, 0, s>>>(w, op, ee, (const T*)d_q, (const T*)d_p, d_var, (T*)d_q_out, (T*)d_p_out, d_energy);
        else k_lf_warp<T><<<nb, 32 * B2_WARPS_PER_BLOCK, 0, s>>>(w, op, ee, (const T*)d_q, (const T*)d_p, d_var, (T*)d_q_out, (T*)d_p_out, d_energy);
        e->launches += 1;
    };
    const T* qA = w.V(B2_V_QE0, 0);
    T* gA = w.V(B2_V_GE0, 0);
    phase(B2_LF_LOAD, 0.0);
    int rc = launch_likelihood<T>(e, qA, qA, gA, gA, e->Dp, nullptr, e->C, e->logp_eval, glm_path, s);   // compute_state
    if (rc) return rc;
    if (n_steps == 0) phase(B2_LF_FINISH, 0.0);           // energy of the start state (no kick)
    for (int i = 0; i < n_steps; ++i) {
        phase(B2_LF_PREPARE, eps);
        rc = launch_likelihood<T>(e, qA, qA, gA, gA, e->Dp, nullptr, e->C, e->logp_eval, glm_path, s);
        if (rc) return rc;
        phase(B2_LF_FINISH, eps);
    }
    phase(B2_LF_STORE, 0.0);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int b2_leapfrog(b2_engine* e, const void* d_q, const void* d_p, const double* d_var, double epsilon,
                           int32_t n_steps, void* d_q_out, void* d_p_out, double* d_energy_out, int32_t glm_path,
                           void* stream) {
    if (!e || !d_q || !d_p || !d_var || !d_q_out || !d_p_out) { b2_set_error("b2_leapfrog: null argument"); return -1; }
    if (n_steps < 0) { b2_set_error("b2_leapfrog: n_steps must be >= 0"); return -2; }
    if (e->stepping) { b2_set_error("b2_leapfrog: a stepwise run is in progress"); return -3; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    e->state_set = false;                  // the edge slots and the mass diagonal were overwritten: b2_set_state before sampling
    return e->dtype == B2_F64
               ? leapfrog_t<double>(e, d_q, d_p, d_var, epsilon, n_steps, d_q_out, d_p_out, d_energy_out, glm_path, (cudaStream_t)stream)
               : leapfrog_t<float>(e, d_q, d_p, d_var, epsilon, n_steps, d_q_out, d_p_out, d_energy_out, glm_path, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- stepwise lock-step API
// For the observation-sharded configuration: every rank runs every chain's state
// machine redundantly on its own rows; the host interleaves a collective between the two halves of
// a leapfrog:  b2_step_likelihood -> all-reduce(packed [C, D+1] fp64) -> b2_step_advance.
template <typename T>
__global__ void k_pack_eval(B2View<T> w, double* __restrict__ packed) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= w.C) return;
    double* row = packed + (size_t)c * (w.D + 1);
    const B2ChainState& s = w.st[c];
    if (!b2_needs_grad(s.phase)) {                 // keep the collective's payload finite for idle chains
        for (int i = lane; i <= w.D; i += 32) row[i] = 0.0;
        return;
    }
    const T* g = w.V(B2_V_GE0 + s.sel, c);
    if (lane == 0) row[0] = w.logp_eval[c];
    for (int i = lane; i < w.D; i += 32) row[1 + i] = (double)g[i];
}

// unpack the reduced values; for the GLM family remove the (world - 1) surplus copies of the prior that
// every rank added to its partial (prior_copies = world_size)
template <typename T>
__global__ void k_unpack_eval(B2View<T> w, const double* __restrict__ packed, int family, double prior_tau, int prior_copies) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= w.C) return;
    const B2ChainState& s = w.st[c];
    if (!b2_needs_grad(s.phase)) return;
    const double* row = packed + (size_t)c * (w.D + 1);
    const T* q = w.V(B2_V_QE0 + s.sel, c);
    T* g = w.V(B2_V_GE0 + s.sel, c);
    const double extra = (double)(prior_copies - 1);
    double prior = 0.0;
    for (int i = lane; i < w.D; i += 32) {
        double gi = row[1 + i];
        if (family == B2_FAMILY_GLM_LOGIT && i > 0 && extra > 0) {
            const double b = (double)q[i];
            gi += extra * prior_tau * b;
            prior += 0.5 * (-prior_tau * b * b + log(prior_tau) - B2_LOG_2PI);
        }
        g[i] = (T)gi;
    }
    for (int o = 16; o > 0; o >>= 1) prior += __shfl_xor_sync(0xffffffffu, prior, o);
    if (lane == 0) w.logp_eval[c] = row[0] - extra * prior;
}

template <typename T>
static int step_likelihood_t(b2_engine* e, double* d_packed, cudaStream_t s) {
    B2View<T> w = build_view<T>(e, &e->step_opts, &e->step_trace);
    const T* qA = w.V(B2_V_QE0, 0); const T* qB = w.V(B2_V_QE1, 0);
    T* gA = w.V(B2_V_GE0, 0); T* gB = w.V(B2_V_GE1, 0);
    int rc = launch_likelihood<T>(e, qA, qB, gA, gB, e->Dp, e->st, e->C, e->logp_eval, e->step_opts.glm_path, s);
    if (rc) return rc;
    k_pack_eval<T><<<(e->C + 3) / 4, 128, 0, s>>>(w, d_packed);
    e->launches += 1;
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
static int step_advance_t(b2_engine* e, const double* d_packed, int prior_copies, cudaStream_t s) {
    B2View<T> w = build_view<T>(e, &e->step_opts, &e->step_trace);
    const int nb = (e->C + B2_WARPS_PER_BLOCK - 1) / B2_WARPS_PER_BLOCK;
    k_unpack_eval<T><<<(e->C + 3) / 4, 128, 0, s>>>(w, d_packed, e->md.family, e->md.hp[0], prior_copies);
    if (use_block_group(e)) k_advance_block<T><<<e->C, B2_BLOCK_NT, 0, s>>>(w, 0);
    else k_advance_warp<T><<<nb, 32 * B2_WARPS_PER_BLOCK, 0, s>>>(w, 0);
    e->launches += 2;
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int b2_step_begin(b2_engine* e, const b2_sampler_opts* o, const b2_trace_out* trace, void* stream) {
    const int rc0 = check_run_args("b2_step_begin", e, o);
    if (rc0) return rc0;
    if (e->dense_L) { b2_set_error("b2_step_begin: the stepwise run does not support a dense mass matrix"); return -5; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    e->step_opts = *o;
    memset(&e->step_trace, 0, sizeof(e->step_trace));
    if (trace) e->step_trace = *trace;
    e->stepping = true;
    if (e->dtype == B2_F32 && b2_glm_tcw_supported(e) && pick_glm_path(e, o->glm_path) == B2_GLM_TCGEN05) {
        B2View<float> w0 = make_view<float>(e);
        int rc = b2_glm_tcw_refresh(e, w0.V(B2_V_QE0, 0), w0.V(B2_V_QE1, 0), e->Dp, e->st, e->C, (cudaStream_t)stream);
        if (rc) return rc;
    }
    if (e->iter_done > 0) {                           // re-activate chains that finished the previous run
        cudaStream_t s = (cudaStream_t)stream;
        const int nb = (e->C + B2_WARPS_PER_BLOCK - 1) / B2_WARPS_PER_BLOCK;
        if (e->dtype == B2_F64) {
            B2View<double> w = build_view<double>(e, o, trace);
            if (use_block_group(e)) k_advance_block<double><<<e->C, B2_BLOCK_NT, 0, s>>>(w, 1);
            else k_advance_warp<double><<<nb, 32 * B2_WARPS_PER_BLOCK, 0, s>>>(w, 1);
        } else {
            B2View<float> w = build_view<float>(e, o, trace);
            if (use_block_group(e)) k_advance_block<float><<<e->C, B2_BLOCK_NT, 0, s>>>(w, 1);
            else k_advance_warp<float><<<nb, 32 * B2_WARPS_PER_BLOCK, 0, s>>>(w, 1);
        }
        e->launches += 1;
        B2_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

extern "C" int b2_step_likelihood(b2_engine* e, double* d_packed, void* stream) {
    if (!e || !d_packed || !e->stepping) { b2_set_error("b2_step_likelihood: call b2_step_begin first"); return -1; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    return e->dtype == B2_F64 ? step_likelihood_t<double>(e, d_packed, (cudaStream_t)stream)
                              : step_likelihood_t<float>(e, d_packed, (cudaStream_t)stream);
}

extern "C" int b2_step_advance(b2_engine* e, const double* d_packed, int32_t prior_copies, void* stream) {
    if (!e || !d_packed || !e->stepping) { b2_set_error("b2_step_advance: call b2_step_begin first"); return -1; }
    if (prior_copies < 1) { b2_set_error("b2_step_advance: prior_copies must be >= 1"); return -2; }
    if (prior_copies > 1 && e->md.family != B2_FAMILY_GLM_LOGIT) {
        // only the GLM family knows how to take the surplus prior copies back out of a sum over row shards
        b2_set_error("b2_step_advance: observation sharding (prior_copies > 1) is implemented for B2_GLM_LOGIT only");
        return -3;
    }
    B2_CUDA_OK(cudaSetDevice(e->device));
    return e->dtype == B2_F64 ? step_advance_t<double>(e, d_packed, prior_copies, (cudaStream_t)stream)
                              : step_advance_t<float>(e, d_packed, prior_copies, (cudaStream_t)stream);
}

// number of chains that still need gradient evaluations (synchronises the stream)
extern "C" int b2_step_active(b2_engine* e, int32_t* host_count, void* stream) {
    if (!e || !host_count) { b2_set_error("b2_step_active: null argument"); return -1; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    B2_CUDA_OK(cudaMemsetAsync(e->d_active, 0, sizeof(int), s));
    k_count_active<<<(e->C + 255) / 256 < 64 ? (e->C + 255) / 256 : 64, 256, 0, s>>>(e->st, e->C, e->iter_done + e->step_opts.n_iters, e->d_active);
    e->launches += 1;
    B2_CUDA_OK(cudaMemcpyAsync(e->h_active, e->d_active, sizeof(int), cudaMemcpyDeviceToHost, s));
    B2_CUDA_OK(cudaStreamSynchronize(s));
    *host_count = *e->h_active;
    return 0;
}

extern "C" int b2_step_end(b2_engine* e) {
    if (!e || !e->stepping) { b2_set_error("b2_step_end: no stepwise run in progress"); return -1; }
    e->iter_done += e->step_opts.n_iters;
    e->stepping = false;
    return 0;
}

extern "C" int b2_get_chain_reports(b2_engine* e, b2_chain_report* out) {
    if (!e || !out) { b2_set_error("b2_get_chain_reports: null argument"); return -1; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    std::vector<B2ChainState> h(e->C);
    B2_CUDA_OK(cudaMemcpy(h.data(), e->st, (size_t)e->C * sizeof(B2ChainState), cudaMemcpyDeviceToHost));
    for (int c = 0; c < e->C; ++c) {
        out[c].phase = h[c].phase; out[c].fail_code = h[c].fail_code; out[c].iter = h[c].iter;
        out[c].n_div_post = h[c].n_div_post; out[c].n_maxdepth_post = h[c].n_maxdepth_post;
        out[c].n_post = h[c].n_post; out[c].n_grad = h[c].n_grad;
        out[c].step_size = exp(h[c].log_step); out[c].step_size_bar = exp(h[c].log_bar);
    }
    return 0;
}

template <typename T>
static int get_vec_t(b2_engine* e, int slot, double* out) {
    std::vector<T> h((size_t)e->C * e->Dp);
    B2_CUDA_OK(cudaMemcpy(h.data(), (T*)e->vec + (size_t)slot * e->C * e->Dp, h.size() * sizeof(T), cudaMemcpyDeviceToHost));
    for (int c = 0; c < e->C; ++c)
        for (int i = 0; i < e->D; ++i) out[(size_t)c * e->D + i] = (double)h[(size_t)c * e->Dp + i];
    return 0;
}

extern "C" int b2_get_mass_var(b2_engine* e, double* out) {
    if (!e || !out) { b2_set_error("b2_get_mass_var: null argument"); return -1; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    return e->dtype == B2_F64 ? get_vec_t<double>(e, B2_V_VAR, out) : get_vec_t<float>(e, B2_V_VAR, out);
}

extern "C" int b2_get_position(b2_engine* e, double* out) {
    if (!e || !out) { b2_set_error("b2_get_position: null argument"); return -1; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    return e->dtype == B2_F64 ? get_vec_t<double>(e, B2_V_PROPQ, out) : get_vec_t<float>(e, B2_V_PROPQ, out);
}

extern "C" int64_t b2_kernel_launches(b2_engine* e) { return e ? e->launches : 0; }

extern "C" int b2_set_profiling(b2_engine* e, int32_t on) {
    if (!e) { b2_set_error("b2_set_profiling: null engine"); return -1; }
    B2_CUDA_OK(cudaSetDevice(e->device));
    if (on && !e->ev[0])
        for (int i = 0; i < 96; ++i) B2_CUDA_OK(cudaEventCreate(&e->ev[i]));
    e->profile = on ? 1 : 0;
    e->like_ms = 0.0; e->like_n = 0; e->adv_ms = 0.0;
    return 0;
}

extern "C" int b2_get_profile(b2_engine* e, double* likelihood_ms, int64_t* likelihood_launches) {
    if (!e || !likelihood_ms || !likelihood_launches) { b2_set_error("b2_get_profile: null argument"); return -1; }
    *likelihood_ms = e->like_ms; *likelihood_launches = e->like_n;
    return 0;
}

extern "C" int b2_get_profile_advance(b2_engine* e, double* advance_ms) {
    if (!e || !advance_ms) { b2_set_error("b2_get_profile_advance: null argument"); return -1; }
    *advance_ms = e->adv_ms;
    return 0;
}
