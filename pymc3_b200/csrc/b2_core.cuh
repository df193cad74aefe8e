// Chain-batched, *iterative* NUTS / HamiltonianMC state machine.
//
// One "group" of threads (a warp, or a whole block for large D) owns one chain.  The chain
// is a small state machine whose unit of progress is ONE leapfrog: `b2_advance()` consumes
// the (logp, grad) that was just evaluated at the chain's pending position, performs all the
// tree bookkeeping that follows from it (second half-kick, energy, divergence test, sub-tree
// merges with U-turn checks and multinomial picks, the top-level merge, end-of-transition
// adaptation and trace write, the next transition's momentum draw ...) and leaves behind the
// next position to evaluate.  Lock-step kernels alternate {gradient kernel, advance kernel};
// the persistent kernel loops {model gradient, advance} inside one launch.
//
// Reference behaviour restated here (no recursion, no Python objects):
//   pymc3/step_methods/hmc/nuts.py:168-188   NUTS._hamiltonian_step (depth loop, early depth 8)
//   pymc3/step_methods/hmc/nuts.py:254-309   _Tree.extend        -> top_merge()
//   pymc3/step_methods/hmc/nuts.py:311-345   _Tree._single_step  -> finish_leaf()
//   pymc3/step_methods/hmc/nuts.py:347-389   _Tree._build_subtree-> binary-counter merges
//   pymc3/step_methods/hmc/nuts.py:391-406   _Tree.stats
//   pymc3/step_methods/hmc/integration.py:81-109  leapfrog
//   pymc3/step_methods/hmc/hmc.py:110-152    HamiltonianMC._hamiltonian_step
//   pymc3/step_methods/hmc/base_hmc.py:133-199  BaseHMC.astep
//   pymc3/step_methods/step_sizes.py:21-58   DualAverageAdaptation
//   pymc3/step_methods/hmc/quadpotential.py:140-225, 313-353  QuadPotentialDiagAdapt
//
// The header is host/device neutral: tests/hostsim builds it with a one-thread "group" so the
// exact same logic is checked against the CPU oracle without a GPU.  The product never runs
// that build (the shipped library contains device kernels only).
#pragma once
#include "b2_philox.cuh"

#define B2_MAX_LEVELS 12          // stack buffers; max_treedepth <= 12
#ifndef B2_WELFORD_WIDTH
#define B2_WELFORD_WIDTH 4        // components per pass of the mass-matrix update (b2_glm_tc.cu builds with 2: register budget)
#endif

enum { B2_PHASE_INIT = 0, B2_PHASE_TREE = 1, B2_PHASE_HMC = 2, B2_PHASE_DONE = 3, B2_PHASE_FAILED = 4,
       B2_PHASE_RESUME = 5 };   // RESUME: a finished chain asked to continue (next b2_sample_run call)
enum { B2_KIND_NUTS = 0, B2_KIND_HMC = 1 };
enum { B2_FAIL_NONE = 0, B2_FAIL_BAD_INITIAL_ENERGY = 1 };

// vector slots in the per-engine state buffer, each [n_chains][Dp]
enum {
    B2_V_QE0 = 0, B2_V_QE1, B2_V_PE0, B2_V_PE1, B2_V_GE0, B2_V_GE1,   // left / right edge (q, p, grad)
    B2_V_VAR,                                                          // the 7 slots above are touched by EVERY leapfrog,
    B2_V_POLD, B2_V_PSUM, B2_V_PROPQ, B2_V_PROPG,                      // these four once per tree doubling / transition
    B2_V_STACK0,                                                       // then 5 per stack buffer
};
#define B2_V_EVERY_LEAPFROG (B2_V_VAR + 1)
enum { B2_S_PFIRST = 0, B2_S_PLAST, B2_S_PSUM, B2_S_Q, B2_S_G, B2_S_NVEC };
#define B2_NUM_VEC_SLOTS (B2_V_STACK0 + B2_S_NVEC * B2_MAX_LEVELS)

struct B2ChainState {
    int phase, iter, fail_code, sel;            // sel: which edge holds the pending position
    uint32_t key0, key1;
    // dual averaging (step_sizes.py)
    double log_step, log_bar, hbar, mu;
    int da_count;
    // QuadPotentialDiagAdapt bookkeeping
    int n_seen, window, fg_sel;
    double wv_count[2];
    // current point
    double cur_logp;
    // transition in flight
    double eps, step_used, e0;
    int depth, max_depth, dir, leaf_n;
    double log_size, log_accept, max_de, prop_energy, prop_logp;
    int n_prop, diverged, turned;
    unsigned long long slot_map;
    // HMC
    int hmc_n_steps, hmc_step;
    // run counters
    int n_div_post, n_maxdepth_post, n_post;
    long long n_grad;
};

template <typename T>
struct B2View {
    // geometry
    int C, D, Dp;
    // state
    T* vec;                       // [B2_NUM_VEC_SLOTS][C][Dp]
    double* wv_mean;              // [2][C][Dp]  Welford means   (always fp64, quadpotential.py:321-327)
    double* wv_m2;                // [2][C][Dp]  Welford raw variances
    B2ChainState* st;             // [C]
    double* lv;                   // [C][4][B2_MAX_LEVELS] per stack buffer: log_size, log_accept_sum, energy, logp
                                  // (kept out of B2ChainState so the per-launch state copy stays ~230 B)
    double* logp_eval;            // [C]  written by the gradient kernel (lock-step mode)
    long long* dbg;               // optional clock stamps of chain 0 (profiling builds of the lock-step kernels)
    // sampler options
    int kind;                     // NUTS | HMC
    int iter_base, iter_end, tune_until;
    int iter_cap;                 // a chain stops here: iter_end, or later when the run lets fast chains run ahead
    int max_treedepth, early_max_treedepth;
    double emax, target, gamma, k, t0;
    int adapt_step, adapt_mass;
    double path_length;
    int max_steps, hmc_jitter;
    // trace outputs (row = iter - iter_base), any may be null
    T* tr_q;                      // [n][C][D]
    double *tr_energy, *tr_energy_error, *tr_max_energy_error, *tr_mean_tree_accept;
    double *tr_step_size, *tr_step_size_bar, *tr_model_logp, *tr_accept;
    int *tr_depth, *tr_tree_size, *tr_n_steps;
    unsigned char *tr_diverging, *tr_tune, *tr_accepted;

    // Optional on-chip copy of THIS chain's first `hot_slots` vector slots (all 11: edges, var, p_old, p_sum,
    // proposal; or only the B2_V_EVERY_LEAPFROG first): [hot_slots][Dp], staged in shared memory by the kernel
    // for the duration of one launch.
    T* hot;
    int hot_slots;
    B2_HD T* V(int slot, int c) const {
        if (hot && slot < hot_slots) return hot + (size_t)slot * Dp;
        return Vglobal(slot, c);
    }
    // (slot stride and the chain's offset are loop invariants of every kernel: one multiply-add per call is left)
    B2_HD T* Vglobal(int slot, int c) const { return vec + (size_t)c * Dp + (size_t)slot * ((size_t)C * Dp); }
    double* lv_hot;               // optional shared-memory copy of this chain's [4][B2_MAX_LEVELS] scalars
    B2_HD double& LV(int c, int which, int buf) const {
        if (lv_hot) return lv_hot[which * B2_MAX_LEVELS + buf];
        return lv[((size_t)c * 4 + which) * B2_MAX_LEVELS + buf];
    }
    // Optional on-chip copy of SOME of this chain's stack buffers (the ones the pending leaf will merge):
    // [n_staged][B2_S_NVEC][Dp]; bit `buf` of stk_mask says buffer `buf` is staged, its position is the
    // 4-bit field `buf` of stk_idx.  Filled and written back by the kernel that sets it.
    T* stk_hot;
    unsigned stk_mask;
    unsigned long long stk_idx;
    B2_HD T* S(int buf, int which, int c) const {
        if (stk_hot && ((stk_mask >> buf) & 1u))
            return stk_hot + ((size_t)((stk_idx >> (4 * buf)) & 0xFull) * B2_S_NVEC + which) * Dp;
        return V(B2_V_STACK0 + buf * B2_S_NVEC + which, c);
    }
};

#if defined(__CUDA_ARCH__)
// (stamps 3.. are taken after the leapfrog counter moved on: same row as stamps 0..2 of this launch)
#define B2_STAMP(w, c, s, k) do { if ((w).dbg && (c) == 0) (w).dbg[(((s).n_grad - ((k) >= 3 ? 1 : 0)) & 4095) * 16 + (k)] = clock64(); } while (0)
#else
#define B2_STAMP(w, c, s, k) do { } while (0)
#endif

B2_HD bool b2_needs_grad(int phase) {
    return phase == B2_PHASE_INIT || phase == B2_PHASE_TREE || phase == B2_PHASE_HMC;
}

// ----------------------------------------------------------------------------- thread groups
struct B2HostGroup {                       // tests only: one "lane"
    static constexpr int NT = 1;
    B2_HD int lane() const { return 0; }
    template <int K> B2_HD void allsum(double (&x)[K]) const {}
    B2_HD void sync() const {}
};

#if defined(__CUDACC__)
struct B2WarpGroup {                       // warp per chain
    static constexpr int NT = 32;
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    template <int K> __device__ __forceinline__ void allsum(double (&x)[K]) const {
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x[k] += __shfl_xor_sync(0xffffffffu, x[k], o);
        }
    }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

template <int NTHREADS>
struct B2BlockGroup {                      // block per chain (large D)
    static constexpr int NT = NTHREADS;
    double* red;                           // shared scratch, >= 2 * 8 * (NT/32) doubles: two halves used alternately
    mutable int flip;                      // which half the next reduction writes (same in every thread)
    __device__ __forceinline__ int lane() const { return threadIdx.x; }
    template <int K> __device__ __forceinline__ void allsum(double (&x)[K]) const {
        static_assert(K <= 8, "scratch sized for 8 simultaneous sums");
        constexpr int NW = NT / 32;
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x[k] += __shfl_xor_sync(0xffffffffu, x[k], o);
        }
        // One barrier per reduction: reduction n writes half n & 1.  A thread can only get to writing that half again
        // (reduction n + 2) after the barrier of reduction n + 1, which every thread passes after it has read the
        // results of reduction n.
        double* buf = red + flip * (8 * NW);
        flip ^= 1;
        if (l == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) buf[k * NW + w] = x[k];
        }
        __syncthreads();
        if (NW <= 8) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double s = 0.0;
                for (int j = 0; j < NW; ++j) s += buf[k * NW + j];   // same order in every thread
                x[k] = s;
            }
        } else {                           // 16 or 32 warps: one more butterfly, identical in every warp
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double s = l < NW ? buf[k * NW + l] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                x[k] = s;
            }
        }
    }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
#endif

// ----------------------------------------------------------------------------- small math
// Scalar transcendentals of the tree bookkeeping.  The fp64 check build evaluates them in double (to
// match the oracle draw for draw); the fp32 production build evaluates the *transcendental part* in
// float -- they only feed log-weights / acceptance probabilities, where 1e-7 relative is far below the
// Monte-Carlo noise, and a single warp executing double log/exp/pow chains was a visible share of the
// lock-step advance kernel (profiles/post_timeline.py).  Energies and log-densities stay fp64.
template <typename T> struct B2M;
template <> struct B2M<double> {
    B2_HD static double log_(double x) { return log(x); }
    B2_HD static double exp_(double x) { return exp(x); }
    B2_HD static double log1p_(double x) { return log1p(x); }
    B2_HD static double expm1_(double x) { return expm1(x); }
    B2_HD static double pow_(double x, double y) { return pow(x, y); }
};
template <> struct B2M<float> {
    B2_HD static double log_(double x) { return (double)logf((float)x); }
    B2_HD static double exp_(double x) { return (double)expf((float)x); }
    B2_HD static double log1p_(double x) { return (double)log1pf((float)x); }
    B2_HD static double expm1_(double x) { return (double)expm1f((float)x); }
    B2_HD static double pow_(double x, double y) { return (double)powf((float)x, (float)y); }
};

template <typename T>
B2_HD double b2_logaddexp(double a, double b) {          // numpy.logaddexp semantics
    if (a == b) return a + 0.693147180559945309417232121458;
    double d = a - b;
    if (d > 0) return a + B2M<T>::log1p_(B2M<T>::exp_(-d));
    if (d <= 0) return b + B2M<T>::log1p_(B2M<T>::exp_(d));
    return a + b;                                        // NaN
}
B2_HD bool b2_finite(double x) { return (x - x) == 0.0; }

B2_HD int b2_map_get(unsigned long long m, int level) { return (int)((m >> (4 * level)) & 0xFull); }
B2_HD unsigned long long b2_map_swap(unsigned long long m, int a, int b) {
    unsigned long long va = (m >> (4 * a)) & 0xFull, vb = (m >> (4 * b)) & 0xFull;
    m &= ~((0xFull << (4 * a)) | (0xFull << (4 * b)));
    return m | (vb << (4 * a)) | (va << (4 * b));
}
#define B2_MAP_IDENTITY 0xBA9876543210ull

// U-turn tests of nuts.py:361-370 / :298-307 for two adjacent sub-trajectories t1 | t2
// (t1 built earlier / on the left).  v = var (.) p  (quadpotential.py:185-187).
// Optionally writes the merged p_sum to `out_psum` and t2's last momentum to `out_plast`.
template <typename T, typename G>
B2_HD bool b2_uturn(const G& g, int D, const T* var,
                    const T* first1, const T* last1, const T* psum1,
                    const T* first2, const T* last2, const T* psum2,
                    bool extra, T* out_psum, T* out_plast) {
    // a lane's few components are summed in the vector dtype (only the signs of the totals matter), the
    // cross-lane reduction is fp64 like every other reduction
    T part[6] = {(T)0, (T)0, (T)0, (T)0, (T)0, (T)0};
    // Four components per pass, every load of the pass issued before the first store: the outputs alias the inputs
    // (p_sum is updated in place), so a plain loop would serialise one memory round trip per component -- the
    // operands of a block-per-chain group come from L2.
    constexpr int UW = 4;
    for (int i0 = g.lane(); i0 < D; i0 += UW * G::NT) {
        T vr[UW], f1[UW], l1[UW], s1[UW], f2[UW], l2[UW], s2[UW];
#pragma unroll
        for (int u = 0; u < UW; ++u) {
            const int i = i0 + u * G::NT;
            const bool ok = i < D;
            vr[u] = ok ? var[i] : (T)0;
            f1[u] = ok ? first1[i] : (T)0; s1[u] = ok ? psum1[i] : (T)0;
            l2[u] = ok ? last2[i] : (T)0; s2[u] = ok ? psum2[i] : (T)0;
            l1[u] = (ok && extra) ? last1[i] : (T)0; f2[u] = (ok && extra) ? first2[i] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < UW; ++u) {
            const int i = i0 + u * G::NT;
            const T tot = s1[u] + s2[u];
            part[0] += tot * (vr[u] * f1[u]);
            part[1] += tot * (vr[u] * l2[u]);
            if (extra) {
                const T a = s1[u] + f2[u];
                part[2] += a * (vr[u] * f1[u]);
                part[3] += a * (vr[u] * f2[u]);
                const T b = l1[u] + s2[u];
                part[4] += b * (vr[u] * l1[u]);
                part[5] += b * (vr[u] * l2[u]);
            }
            if (i < D) {
                if (out_psum) out_psum[i] = tot;
                if (out_plast) out_plast[i] = l2[u];
            }
        }
    }
    double d[6] = {(double)part[0], (double)part[1], (double)part[2], (double)part[3], (double)part[4], (double)part[5]};
    g.allsum(d);
    bool turning = (d[0] <= 0) || (d[1] <= 0);
    if (extra) turning = turning || (d[2] <= 0) || (d[3] <= 0) || (d[4] <= 0) || (d[5] <= 0);
    return turning;
}

template <typename T, typename G>
B2_HD void b2_copy(const G& g, int D, T* dst, const T* src) {
    for (int i = g.lane(); i < D; i += G::NT) dst[i] = src[i];
}

// two copies with all loads in flight together (proposal position + gradient)
template <typename T, typename G>
B2_HD void b2_copy2(const G& g, int D, T* dst_a, const T* src_a, T* dst_b, const T* src_b) {
    constexpr int UW = 4;                                  // all loads of a pass before its stores (see b2_uturn)
    for (int i0 = g.lane(); i0 < D; i0 += UW * G::NT) {
        T a[UW], b[UW];
#pragma unroll
        for (int u = 0; u < UW; ++u) {
            const int i = i0 + u * G::NT;
            a[u] = i < D ? src_a[i] : (T)0; b[u] = i < D ? src_b[i] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < UW; ++u) {
            const int i = i0 + u * G::NT;
            if (i < D) { dst_a[i] = a[u]; dst_b[i] = b[u]; }
        }
    }
}

// first half-kick + drift, in place on edge `e`:  integration.py:90-99
template <typename T, typename G>
B2_HD void b2_prepare_leapfrog(const G& g, const B2View<T>& w, int c, int e, double eps_signed) {
    T* q = w.V(B2_V_QE0 + e, c);
    T* p = w.V(B2_V_PE0 + e, c);
    const T* gr = w.V(B2_V_GE0 + e, c);
    const T* var = w.V(B2_V_VAR, c);
    const T eps = (T)eps_signed, half = (T)(0.5 * eps_signed);
    for (int i = g.lane(); i < w.D; i += G::NT) {
        const T pm = p[i] + half * gr[i];
        p[i] = pm;
        q[i] = q[i] + eps * (var[i] * pm);
    }
}

// second half-kick and energy:  integration.py:101-107
template <typename T, typename G>
B2_HD double b2_finish_leapfrog(const G& g, const B2View<T>& w, int c, int e, double eps_signed, double logp) {
    T* p = w.V(B2_V_PE0 + e, c);
    const T* gr = w.V(B2_V_GE0 + e, c);
    const T* var = w.V(B2_V_VAR, c);
    const T half = (T)(0.5 * eps_signed);
    T part = (T)0;                                        // a lane's few components in the vector dtype, lanes in fp64
    for (int i = g.lane(); i < w.D; i += G::NT) {
        const T pn = p[i] + half * gr[i];
        p[i] = pn;
        part += pn * (var[i] * pn);
    }
    double kin[1] = {(double)part};
    g.allsum(kin);
    return 0.5 * kin[0] - logp;
}

// ------------------------------------------------------------------------ chain initialisation
// Mirrors what the reference sets up per chain before the first draw:
//   base_hmc.py:93-96  (initial step size -> DualAverageAdaptation, step_sizes.py:22-32)
//   quadpotential.py:140-183 (QuadPotentialDiagAdapt: foreground window seeded with
//   `initial_weight` pseudo-samples of (mean, var); empty background window)
//   sampling.py:883-884 (per-chain seed) -> Philox key.
template <typename T, typename G>
B2_HD void b2_init_chain(const G& g, const B2View<T>& w, int c, B2ChainState& s, const T* q0,
                         unsigned long long seed, double step0, const double* mass_mean,
                         const double* mass_var, double mass_weight, int window, int iter0) {
    s.phase = B2_PHASE_INIT; s.iter = iter0; s.fail_code = B2_FAIL_NONE; s.sel = 1;
    s.key0 = (uint32_t)(seed & 0xFFFFFFFFull); s.key1 = (uint32_t)(seed >> 32);
    s.log_step = log(step0); s.log_bar = s.log_step; s.hbar = 0.0; s.mu = log(10.0 * step0); s.da_count = 1;
    s.n_seen = 0; s.window = window; s.fg_sel = 0; s.wv_count[0] = mass_weight; s.wv_count[1] = 0.0;
    s.cur_logp = 0.0; s.eps = step0; s.step_used = step0; s.e0 = 0.0;
    s.depth = 0; s.max_depth = 0; s.dir = 1; s.leaf_n = 0;
    s.log_size = 0.0; s.log_accept = 0.0; s.max_de = 0.0; s.prop_energy = 0.0; s.prop_logp = 0.0;
    s.n_prop = 0; s.diverged = 0; s.turned = 0; s.slot_map = B2_MAP_IDENTITY;
    s.hmc_n_steps = 0; s.hmc_step = 0;
    s.n_div_post = 0; s.n_maxdepth_post = 0; s.n_post = 0; s.n_grad = 0;
    T *pq = w.V(B2_V_PROPQ, c), *q1 = w.V(B2_V_QE1, c), *var = w.V(B2_V_VAR, c);
    for (int i = g.lane(); i < w.D; i += G::NT) {
        pq[i] = q0[i]; q1[i] = q0[i];
        var[i] = (T)mass_var[i];
        const size_t o0 = ((size_t)0 * w.C + c) * w.Dp + i, o1 = ((size_t)1 * w.C + c) * w.Dp + i;
        w.wv_mean[o0] = mass_mean[i]; w.wv_m2[o0] = mass_var[i] * mass_weight;
        w.wv_mean[o1] = 0.0; w.wv_m2[o1] = 0.0;
    }
}

// ------------------------------------------------------------------- transition start / end
// The pieces below return an "action" so that b2_advance() can sequence them with exactly one
// call site each (keeps the inlined device code small and the control flow group-uniform).
enum { B2_ACT_NONE = 0, B2_ACT_TOP_MERGE, B2_ACT_END_NUTS, B2_ACT_END_NUTS_MAXDEPTH, B2_ACT_END,
       B2_ACT_BEGIN_TRANSITION, B2_ACT_BEGIN_DOUBLING };

struct B2EndStats { double accept_stat, energy, energy_error, model_logp; bool accepted; };

template <typename T, typename G>
B2_HD void b2_begin_doubling(const G& g, const B2View<T>& w, int c, B2ChainState& s) {
    const double u = b2_uniform(s.key0, s.key1, (uint32_t)s.iter, B2_PURPOSE_DIRECTION, (uint32_t)s.depth, 0u);
    s.dir = (u < 0.5) ? 1 : 0;                              // log(u) < log(0.5)        // nuts.py:177
    b2_copy(g, w.D, w.V(B2_V_POLD, c), w.V(B2_V_PE0 + s.dir, c));
    s.leaf_n = 0;
    s.slot_map = B2_MAP_IDENTITY;
    b2_prepare_leapfrog(g, w, c, s.dir, s.dir ? s.eps : -s.eps);
    s.sel = s.dir;
    s.phase = B2_PHASE_TREE;
}

template <typename T, typename G>
B2_HD int b2_begin_transition(const G& g, const B2View<T>& w, int c, B2ChainState& s) {
    const uint32_t t = (uint32_t)s.iter;
    const bool tune = s.iter < w.tune_until;
    const T* var = w.V(B2_V_VAR, c);
    const T* pq = w.V(B2_V_PROPQ, c);
    const T* pg = w.V(B2_V_PROPG, c);
    T *q0 = w.V(B2_V_QE0, c), *q1 = w.V(B2_V_QE1, c), *p0 = w.V(B2_V_PE0, c), *p1 = w.V(B2_V_PE1, c);
    T *g0 = w.V(B2_V_GE0, c), *g1 = w.V(B2_V_GE1, c), *ps = w.V(B2_V_PSUM, c);
    const bool nuts = (w.kind == B2_KIND_NUTS);
    double kin[1] = {0.0};
    // a lane draws PAIRS of components: both Box-Muller outputs of a Philox block are used (drawing them one
    // component at a time computed every block twice; the momentum draw was 20 % of a transition end)
    for (int j = g.lane(); 2 * j < w.D; j += G::NT) {
        T n01[2];
        B2Normal<T>::draw2(s.key0, s.key1, t, (uint32_t)j, n01[0], n01[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = 2 * j + h;
            if (i < w.D) {
                // quadpotential.py:200-203: inv_stds * normal,  inv_stds = 1 / sqrt(var)
                const T inv_std = (T)1 / (T)sqrt(var[i]);
                const T p = inv_std * n01[h];
                const T qq = pq[i], gg = pg[i];
                q1[i] = qq; p1[i] = p; g1[i] = gg;
                if (nuts) { q0[i] = qq; p0[i] = p; g0[i] = gg; ps[i] = p; }
                kin[0] += (double)(p * (var[i] * p));
            }
        }
    }
    g.allsum(kin);
    s.e0 = 0.5 * kin[0] - s.cur_logp;                     // integration.py:45-46
    if (!b2_finite(s.e0)) {                               // base_hmc.py:138-158
        s.phase = B2_PHASE_FAILED;
        s.fail_code = B2_FAIL_BAD_INITIAL_ENERGY;
        return B2_ACT_NONE;
    }
    const bool adapt = tune && w.adapt_step;
    s.step_used = B2M<T>::exp_(adapt ? s.log_step : s.log_bar);   // step_sizes.py:34-38
    s.eps = s.step_used;
    s.diverged = 0; s.turned = 0;
    if (nuts) {
        s.depth = 0; s.log_size = 0.0; s.log_accept = -INFINITY; s.n_prop = 0; s.max_de = 0.0;
        s.prop_energy = s.e0; s.prop_logp = s.cur_logp;
        s.max_depth = (tune && s.iter < 200) ? w.early_max_treedepth : w.max_treedepth;   // nuts.py:169-172
        return B2_ACT_BEGIN_DOUBLING;
    }
    if (w.hmc_jitter)                                     // hmc.py:26-27 via base_hmc.py:164-165
        s.eps = (0.85 + 0.30 * b2_uniform(s.key0, s.key1, t, B2_PURPOSE_HMC_JITTER, 0u, 0u)) * s.step_used;
    int n = (int)(w.path_length / s.eps);                 // hmc.py:111-112
    n = n < 1 ? 1 : n;
    s.hmc_n_steps = n > w.max_steps ? w.max_steps : n;
    s.hmc_step = 0;
    b2_prepare_leapfrog(g, w, c, 1, s.eps);
    s.sel = 1;
    s.phase = B2_PHASE_HMC;
    return B2_ACT_NONE;
}

// base_hmc.py:169-199 tail + sampling.py:921-930 record
template <typename T, typename G>
B2_HD int b2_end_transition(const G& g, const B2View<T>& w, int c, B2ChainState& s, const B2EndStats& es) {
    const bool tune = s.iter < w.tune_until;
    const bool adapt = tune && w.adapt_step;
    if (adapt) {                                          // step_sizes.py:40-52
        const double cnt = (double)s.da_count;
        const double ww = 1.0 / (cnt + w.t0);
        s.hbar = (1.0 - ww) * s.hbar + ww * (w.target - es.accept_stat);
        s.log_step = s.mu - s.hbar * sqrt(cnt) / w.gamma;
        const double mk = B2M<T>::pow_(cnt, -w.k);
        s.log_bar = mk * s.log_step + (1.0 - mk) * s.log_bar;
        s.da_count += 1;
    }
    const T* pq = w.V(B2_V_PROPQ, c);
    if (tune && w.adapt_mass) {                           // quadpotential.py:211-225, 336-350
        T* var = w.V(B2_V_VAR, c);
        const int f = s.fg_sel, b = 1 - s.fg_sel;
        double* fm = w.wv_mean + ((size_t)f * w.C + c) * w.Dp;
        double* f2 = w.wv_m2 + ((size_t)f * w.C + c) * w.Dp;
        double* bm = w.wv_mean + ((size_t)b * w.C + c) * w.Dp;
        double* b2 = w.wv_m2 + ((size_t)b * w.C + c) * w.Dp;
        const double nf = s.wv_count[f] + 1.0, nb = s.wv_count[b] + 1.0;
        const bool swap = (s.n_seen > 0) && (s.n_seen % s.window == 0);
        // WW (four) components per pass, loads first and stores last: 16 loads in flight and 12 independent fp64
        // divisions instead of one load round trip and three dependent divisions per component (a chain that
        // ends its transition is the critical path of a lock-step launch)
        constexpr int WW = B2_WELFORD_WIDTH;
        for (int i0 = g.lane(); i0 < w.D; i0 += WW * G::NT) {
            double x[WW], mf[WW], rf[WW], mb[WW], rb[WW];
#pragma unroll
            for (int u = 0; u < WW; ++u) {
                const int i = i0 + u * G::NT;
                const bool ok = i < w.D;
                x[u] = ok ? (double)pq[i] : 0.0;
                mf[u] = ok ? fm[i] : 0.0; rf[u] = ok ? f2[i] : 0.0;
                mb[u] = ok ? bm[i] : 0.0; rb[u] = ok ? b2[i] : 0.0;
            }
            T vnew[WW];
#pragma unroll
            for (int u = 0; u < WW; ++u) {
                double od = x[u] - mf[u];
                mf[u] = mf[u] + od / nf;
                rf[u] = rf[u] + od * (x[u] - mf[u]);
                vnew[u] = (T)(rf[u] / nf);
                od = x[u] - mb[u];
                mb[u] = mb[u] + od / nb;
                rb[u] = rb[u] + od * (x[u] - mb[u]);
            }
#pragma unroll
            for (int u = 0; u < WW; ++u) {
                const int i = i0 + u * G::NT;
                if (i < w.D) {
                    var[i] = vnew[u];
                    if (swap) { fm[i] = 0.0; f2[i] = 0.0; } else { fm[i] = mf[u]; f2[i] = rf[u]; }
                    bm[i] = mb[u]; b2[i] = rb[u];
                }
            }
        }
        s.wv_count[f] = nf; s.wv_count[b] = nb;
        if (swap) { s.wv_count[f] = 0.0; s.fg_sel = b; }   // background becomes foreground
        s.n_seen += 1;
    }
    if (!tune) {
        s.n_post += 1;
        if (s.diverged) s.n_div_post += 1;
    }
    const int row = s.iter - w.iter_base;
    const size_t o = (size_t)row * w.C + c;
    if (w.tr_q) {
        T* dst = w.tr_q + o * w.D;
        for (int i = g.lane(); i < w.D; i += G::NT) dst[i] = pq[i];
    }
    if (g.lane() == 0) {
        if (w.tr_energy) w.tr_energy[o] = es.energy;
        if (w.tr_energy_error) w.tr_energy_error[o] = es.energy_error;
        if (w.tr_model_logp) w.tr_model_logp[o] = es.model_logp;
        if (w.tr_step_size) w.tr_step_size[o] = B2M<T>::exp_(s.log_step);           // step_sizes.py:54-58
        if (w.tr_step_size_bar) w.tr_step_size_bar[o] = B2M<T>::exp_(s.log_bar);
        if (w.tr_diverging) w.tr_diverging[o] = (unsigned char)(s.diverged != 0);
        if (w.tr_tune) w.tr_tune[o] = (unsigned char)tune;
        if (w.kind == B2_KIND_NUTS) {
            if (w.tr_max_energy_error) w.tr_max_energy_error[o] = s.max_de;
            if (w.tr_mean_tree_accept) w.tr_mean_tree_accept[o] = es.accept_stat;
            if (w.tr_depth) w.tr_depth[o] = s.depth;
            if (w.tr_tree_size) w.tr_tree_size[o] = s.n_prop;
        } else {
            if (w.tr_accept) w.tr_accept[o] = es.accept_stat;
            if (w.tr_accepted) w.tr_accepted[o] = (unsigned char)es.accepted;
            if (w.tr_n_steps) w.tr_n_steps[o] = s.hmc_n_steps;
        }
    }
    s.iter += 1;
    if (s.iter >= w.iter_cap) { s.phase = B2_PHASE_DONE; return B2_ACT_NONE; }
    return B2_ACT_BEGIN_TRANSITION;
}

// nuts.py:283-309: merge a completed sub-tree (stack level == old depth) into the main tree
template <typename T, typename G>
B2_HD int b2_top_merge(const G& g, const B2View<T>& w, int c, B2ChainState& s) {
    const int d_old = s.depth;
    const int buf = b2_map_get(s.slot_map, d_old);
    s.depth += 1;
    s.n_prop += (1 << d_old);
    const double sub_ls = w.LV(c, 0, buf);
    const double u = b2_uniform(s.key0, s.key1, (uint32_t)s.iter, B2_PURPOSE_TOP, (uint32_t)d_old, 0u);
    if (B2M<T>::log_(u) < sub_ls - s.log_size) {                   // nuts.py:289-291
        b2_copy2(g, w.D, w.V(B2_V_PROPQ, c), w.S(buf, B2_S_Q, c), w.V(B2_V_PROPG, c), w.S(buf, B2_S_G, c));
        s.prop_energy = w.LV(c, 2, buf);
        s.prop_logp = w.LV(c, 3, buf);
    }
    s.log_size = b2_logaddexp<T>(s.log_size, sub_ls);
    s.log_accept = b2_logaddexp<T>(s.log_accept, w.LV(c, 1, buf));
    const T* var = w.V(B2_V_VAR, c);
    T* psum = w.V(B2_V_PSUM, c);
    // dir=1: main tree | new sub-tree.   dir=0: new sub-tree (reversed in time) | main tree.
    const T* first1 = s.dir ? w.V(B2_V_PE0, c) : w.S(buf, B2_S_PLAST, c);
    const T* last1 = s.dir ? w.V(B2_V_POLD, c) : w.S(buf, B2_S_PFIRST, c);
    const T* psum1 = s.dir ? psum : w.S(buf, B2_S_PSUM, c);
    const T* first2 = s.dir ? w.S(buf, B2_S_PFIRST, c) : w.V(B2_V_POLD, c);
    const T* last2 = s.dir ? w.S(buf, B2_S_PLAST, c) : w.V(B2_V_PE1, c);
    const T* psum2 = s.dir ? w.S(buf, B2_S_PSUM, c) : psum;
    const bool turning = b2_uturn<T, G>(g, w.D, var, first1, last1, psum1, first2, last2, psum2, true, psum, (T*)0);
    if (turning) { s.turned = 1; return B2_ACT_END_NUTS; }
    if (s.depth >= s.max_depth) return B2_ACT_END_NUTS_MAXDEPTH;
    return B2_ACT_BEGIN_DOUBLING;
}

// nuts.py:311-345 then the binary-counter form of nuts.py:347-389
template <typename T, typename G>
B2_HD int b2_finish_leaf(const G& g, const B2View<T>& w, int c, B2ChainState& s, double logp_new) {
    const int e = s.dir;
    const double energy = b2_finish_leapfrog(g, w, c, e, e ? s.eps : -s.eps, logp_new);
    s.n_grad += 1;
    double de = energy - s.e0;
    if (de != de) de = INFINITY;                          // nuts.py:322-323
    if (fabs(de) > fabs(s.max_de)) s.max_de = de;
    if (!(fabs(de) < w.emax)) {                           // divergence, nuts.py:338-345
        s.diverged = 1;
        s.depth += 1;
        s.n_prop += s.leaf_n + 1;
        return B2_ACT_END_NUTS;
    }
    const double leaf_ls = -de;
    const double leaf_la = -de + (-de < 0.0 ? -de : 0.0);  // nuts.py:331 (sic)
    const T* qe = w.V(B2_V_QE0 + e, c);
    const T* pe = w.V(B2_V_PE0 + e, c);
    const T* ge = w.V(B2_V_GE0 + e, c);
    const T* var = w.V(B2_V_VAR, c);
    const int n = s.leaf_n;
    int j = 0;
    while ((n >> j) & 1) ++j;                             // trailing ones = number of merges
    if (j == 0) {
        const int buf = b2_map_get(s.slot_map, 0);
        T *sf = w.S(buf, B2_S_PFIRST, c), *sl = w.S(buf, B2_S_PLAST, c), *ss = w.S(buf, B2_S_PSUM, c);
        T *sq = w.S(buf, B2_S_Q, c), *sg = w.S(buf, B2_S_G, c);
        // A one-leaf sub-tree has p_first = p_last = p_sum.  Only the depth-0 doubling reads all three back (top merge);
        // otherwise the next leaf's merge takes p_first for all of them and writes p_last / p_sum itself.
        const bool solo = (s.depth == 0);
        for (int i = g.lane(); i < w.D; i += G::NT) {
            const T p = pe[i];
            sf[i] = p; sq[i] = qe[i]; sg[i] = ge[i];
            if (solo) { sl[i] = p; ss[i] = p; }
        }
        w.LV(c, 0, buf) = leaf_ls; w.LV(c, 1, buf) = leaf_la;
        w.LV(c, 2, buf) = energy; w.LV(c, 3, buf) = logp_new;
    } else {
        for (int k = 0; k < j; ++k) {
            const int b1 = b2_map_get(s.slot_map, k);
            const int b2i = b2_map_get(s.slot_map, k > 0 ? k - 1 : 0);
            const bool leaf2 = (k == 0);
            // t2 always ends in the leaf that was just finished: its p_last is the edge's momentum (on chip), and the
            // p_last of an intermediate result is never read -- only the chain's last merge stores it
            const T* f2 = leaf2 ? pe : w.S(b2i, B2_S_PFIRST, c);
            const T* l2 = pe;
            const T* s2 = leaf2 ? pe : w.S(b2i, B2_S_PSUM, c);
            const T* q2 = leaf2 ? qe : w.S(b2i, B2_S_Q, c);
            const T* g2 = leaf2 ? ge : w.S(b2i, B2_S_G, c);
            const double ls2 = leaf2 ? leaf_ls : w.LV(c, 0, b2i);
            const double la2 = leaf2 ? leaf_la : w.LV(c, 1, b2i);
            const double en2 = leaf2 ? energy : w.LV(c, 2, b2i);
            const double lp2 = leaf2 ? logp_new : w.LV(c, 3, b2i);
            T *f1 = w.S(b1, B2_S_PFIRST, c), *l1 = w.S(b1, B2_S_PLAST, c), *s1 = w.S(b1, B2_S_PSUM, c);
            // k == 0: t1 is the single leaf stored one step ago (p_first stands for its p_last and p_sum)
            const bool turning = b2_uturn<T, G>(g, w.D, var, f1, leaf2 ? f1 : l1, leaf2 ? f1 : s1, f2, l2, s2, k > 0, s1,
                                                k == j - 1 ? l1 : (T*)0);
            if (turning) { s.turned = 1; break; }
            const double ls1 = w.LV(c, 0, b1);
            const double la1 = w.LV(c, 1, b1);
            // Every lane of the group executes this read-modify-write redundantly; all of them must have
            // read the old level scalars before any lane stores the merged ones (warps of a block group
            // run out of step, and a late reader would fold the sub-tree in twice).
            g.sync();
            const double ls = b2_logaddexp<T>(ls1, ls2);
            const double u = b2_uniform(s.key0, s.key1, (uint32_t)s.iter, B2_PURPOSE_MERGE, (uint32_t)s.depth,
                                        ((uint32_t)(k + 1) << 16) | (uint32_t)n);
            if (B2M<T>::log_(u) < ls2 - ls) {                      // nuts.py:375-378
                b2_copy2(g, w.D, w.S(b1, B2_S_Q, c), q2, w.S(b1, B2_S_G, c), g2);
                w.LV(c, 2, b1) = en2; w.LV(c, 3, b1) = lp2;
            }
            w.LV(c, 0, b1) = ls;
            w.LV(c, 1, b1) = b2_logaddexp<T>(la1, la2);
        }
        if (s.turned) {
            s.depth += 1;
            s.n_prop += n + 1;
            return B2_ACT_END_NUTS;
        }
        s.slot_map = b2_map_swap(s.slot_map, j, j - 1);
    }
    s.leaf_n = n + 1;
    if (s.leaf_n == (1 << s.depth)) return B2_ACT_TOP_MERGE;
    b2_prepare_leapfrog(g, w, c, e, e ? s.eps : -s.eps);
    return B2_ACT_NONE;
}

// hmc.py:110-152
template <typename T, typename G>
B2_HD int b2_finish_hmc_step(const G& g, const B2View<T>& w, int c, B2ChainState& s, double logp_new, B2EndStats& es) {
    const double energy = b2_finish_leapfrog(g, w, c, 1, s.eps, logp_new);
    s.n_grad += 1;
    s.hmc_step += 1;
    if (s.hmc_step < s.hmc_n_steps) { b2_prepare_leapfrog(g, w, c, 1, s.eps); return B2_ACT_NONE; }
    bool div = !b2_finite(energy);                        // :123-125
    double de = s.e0 - energy;
    if (de != de) de = -INFINITY;
    if (fabs(de) > w.emax) div = true;                    // :129
    const double ex = B2M<T>::exp_(de);
    const double accept = ex < 1.0 ? ex : 1.0;
    bool accepted = false;
    if (!div) {
        const double u = b2_uniform(s.key0, s.key1, (uint32_t)s.iter, B2_PURPOSE_HMC_ACCEPT, 0u, 0u);
        accepted = !(u >= accept);                        // :136
    }
    s.diverged = div ? 1 : 0;
    if (accepted) {
        b2_copy2(g, w.D, w.V(B2_V_PROPQ, c), w.V(B2_V_QE1, c), w.V(B2_V_PROPG, c), w.V(B2_V_GE1, c));
        s.cur_logp = logp_new;
    }
    es.accept_stat = accept; es.energy = energy; es.energy_error = de; es.model_logp = logp_new;
    es.accepted = accepted;
    return B2_ACT_END;
}

// One unit of progress for chain c.  Returns true while the chain still needs gradients.
template <typename T, typename G>
B2_HD bool b2_advance(const G& g, const B2View<T>& w, int c, B2ChainState& s, double logp_new) {
    int act = B2_ACT_NONE;
    B2EndStats es = {0.0, 0.0, 0.0, 0.0, true};
    if (s.phase == B2_PHASE_INIT) {
        s.cur_logp = logp_new;
        s.n_grad += 1;
        b2_copy(g, w.D, w.V(B2_V_PROPG, c), w.V(B2_V_GE1, c));
        if (s.iter >= w.iter_cap) s.phase = B2_PHASE_DONE;
        else act = B2_ACT_BEGIN_TRANSITION;
    } else if (s.phase == B2_PHASE_RESUME) {              // continue from the last draw: no new gradient needed
        if (s.iter >= w.iter_cap) s.phase = B2_PHASE_DONE;
        else act = B2_ACT_BEGIN_TRANSITION;
    } else if (s.phase == B2_PHASE_TREE) {
        act = b2_finish_leaf(g, w, c, s, logp_new);
    } else if (s.phase == B2_PHASE_HMC) {
        act = b2_finish_hmc_step(g, w, c, s, logp_new, es);
    }
    B2_STAMP(w, c, s, 3);
    if (act == B2_ACT_TOP_MERGE) act = b2_top_merge(g, w, c, s);
    B2_STAMP(w, c, s, 4);
    if (act == B2_ACT_END_NUTS || act == B2_ACT_END_NUTS_MAXDEPTH) {
        es.accept_stat = 0.0;                             // nuts.py:391-397
        if (s.log_size > 0.0) {
            es.accept_stat = B2M<T>::exp_(s.log_accept) / B2M<T>::expm1_(s.log_size);
            // Deliberate deviation: for energy drops > ~709 (still below Emax = 1000) the reference's
            // exp()/expm1() is inf/inf = NaN, which poisons dual averaging for the rest of the chain.
            // Same quantity in log space, used only where the reference's expression is not finite.
            if (!b2_finite(es.accept_stat))
                es.accept_stat = B2M<T>::exp_(s.log_accept - (s.log_size + B2M<T>::log1p_(-B2M<T>::exp_(-s.log_size))));
        }
        if (act == B2_ACT_END_NUTS_MAXDEPTH && !(s.iter < w.tune_until)) s.n_maxdepth_post += 1;   // nuts.py:182-184
        s.cur_logp = s.prop_logp;
        es.energy = s.prop_energy; es.energy_error = s.prop_energy - s.e0; es.model_logp = s.prop_logp;
        act = B2_ACT_END;
    }
    if (act == B2_ACT_END) act = b2_end_transition(g, w, c, s, es);
    B2_STAMP(w, c, s, 5);
    if (act == B2_ACT_BEGIN_TRANSITION) act = b2_begin_transition(g, w, c, s);
    B2_STAMP(w, c, s, 6);
    if (act == B2_ACT_BEGIN_DOUBLING) b2_begin_doubling(g, w, c, s);
    B2_STAMP(w, c, s, 7);
    return s.phase == B2_PHASE_TREE || s.phase == B2_PHASE_HMC;
}
