// TEST HARNESS ONLY -- not part of the product, never loaded by pymc3_b200.
//
// Compiles the engine's host/device-neutral state machine (pymc3_b200/csrc/b2_core.cuh,
// b2_models.cuh) for the CPU with a one-lane "thread group", so that the exact sampler logic
// that runs inside the CUDA kernels can be checked against the CPU oracle (oracle/) in the
// `-m "not gpu"` test tier, where no GPU exists.  Built by tests/hostsim/build.py with g++.
#include <cstring>
#include <vector>
#include "../../include/b200nuts.h"
#include "../../pymc3_b200/csrc/b2_models.cuh"

template <typename T>
static int run_t(const B2ModelData& md, int C, const double* q0, const uint64_t* seeds, double step0,
                 const double* mm, const double* mv, double mw, int window, const b2_sampler_opts* o,
                 const b2_trace_out* tr, b2_chain_report* rep) {
    const int D = md.D, Dp = (D + 3) & ~3;
    std::vector<T> vec((size_t)B2_NUM_VEC_SLOTS * C * Dp, (T)0);
    std::vector<double> wm((size_t)2 * C * Dp, 0.0), w2((size_t)2 * C * Dp, 0.0), lp(C, 0.0);
    std::vector<B2ChainState> st(C);
    std::vector<double> lv((size_t)C * 4 * B2_MAX_LEVELS, 0.0);
    std::vector<T> scratch((size_t)(md.family == B2_FAMILY_GLM_LOGIT ? md.N : 1));
    B2View<T> w;
    memset(&w, 0, sizeof(w));
    w.C = C; w.D = D; w.Dp = Dp; w.vec = vec.data(); w.wv_mean = wm.data(); w.wv_m2 = w2.data();
    w.st = st.data(); w.lv = lv.data(); w.logp_eval = lp.data(); w.hot = nullptr; w.lv_hot = nullptr; w.stk_hot = nullptr; w.stk_mask = 0; w.stk_idx = 0;
    w.kind = o->kind; w.iter_base = 0; w.iter_end = o->n_iters; w.iter_cap = o->n_iters; w.tune_until = o->tune_until;
    w.max_treedepth = o->max_treedepth; w.early_max_treedepth = o->early_max_treedepth;
    w.emax = o->Emax; w.target = o->target_accept; w.gamma = o->gamma; w.k = o->k; w.t0 = o->t0;
    w.adapt_step = o->adapt_step_size; w.adapt_mass = o->adapt_mass;
    w.path_length = o->path_length; w.max_steps = o->max_steps; w.hmc_jitter = o->hmc_jitter;
    w.tr_q = (T*)tr->d_q; w.tr_energy = tr->d_energy; w.tr_energy_error = tr->d_energy_error;
    w.tr_max_energy_error = tr->d_max_energy_error; w.tr_mean_tree_accept = tr->d_mean_tree_accept;
    w.tr_step_size = tr->d_step_size; w.tr_step_size_bar = tr->d_step_size_bar; w.tr_model_logp = tr->d_model_logp;
    w.tr_accept = tr->d_accept; w.tr_depth = tr->d_depth; w.tr_tree_size = tr->d_tree_size; w.tr_n_steps = tr->d_n_steps;
    w.tr_diverging = tr->d_diverging; w.tr_tune = tr->d_tune; w.tr_accepted = tr->d_accepted;
    B2HostGroup g;
    B2ModelData m = md;
    m.scratch = scratch.data();
    std::vector<T> q0t(D);
    for (int c = 0; c < C; ++c) {
        for (int i = 0; i < D; ++i) q0t[i] = (T)q0[(size_t)c * D + i];
        b2_init_chain<T, B2HostGroup>(g, w, c, st[c], q0t.data(), seeds[c], step0, mm, mv, mw, window, 0);
        B2ChainState& s = st[c];
        bool active = true;
        while (active) {
            const T* q = w.V(B2_V_QE0 + s.sel, c);
            T* gr = w.V(B2_V_GE0 + s.sel, c);
            const double l = b2_eval_model<T, B2HostGroup>(g, m, q, gr, 0);
            active = b2_advance<T, B2HostGroup>(g, w, c, s, l);
        }
        if (rep) {
            rep[c].phase = s.phase; rep[c].fail_code = s.fail_code; rep[c].iter = s.iter;
            rep[c].n_div_post = s.n_div_post; rep[c].n_maxdepth_post = s.n_maxdepth_post; rep[c].n_post = s.n_post;
            rep[c].n_grad = s.n_grad; rep[c].step_size = exp(s.log_step); rep[c].step_size_bar = exp(s.log_bar);
        }
    }
    return 0;
}

static B2ModelData to_md(const b2_model_desc* d) {
    B2ModelData m;
    memset(&m, 0, sizeof(m));
    m.family = d->family; m.D = d->D; m.N = d->N; m.G = d->G; m.aux0 = d->d_aux0; m.aux1 = d->d_aux1;
    m.X = d->d_X; m.yf = d->d_y; m.floor_u8 = d->d_floor; m.grp_off = d->d_grp_off;
    for (int i = 0; i < 4; ++i) m.hp[i] = d->hp[i];
    return m;
}

extern "C" int hostsim_run(const b2_model_desc* desc, int C, int dtype, const double* q0, const uint64_t* seeds,
                           double step0, const double* mm, const double* mv, double mw, int window,
                           const b2_sampler_opts* o, const b2_trace_out* tr, b2_chain_report* rep) {
    const B2ModelData md = to_md(desc);
    return dtype == B2_F64 ? run_t<double>(md, C, q0, seeds, step0, mm, mv, mw, window, o, tr, rep)
                           : run_t<float>(md, C, q0, seeds, step0, mm, mv, mw, window, o, tr, rep);
}

extern "C" int hostsim_logp(const b2_model_desc* desc, int dtype, const double* q, int n, double* logp, double* grad) {
    B2ModelData m = to_md(desc);
    B2HostGroup g;
    const int D = m.D;
    if (dtype == B2_F64) {
        std::vector<double> scratch(m.N > 0 ? m.N : 1);
        m.scratch = scratch.data();
        for (int p = 0; p < n; ++p) logp[p] = b2_eval_model<double, B2HostGroup>(g, m, q + (size_t)p * D, grad + (size_t)p * D, 0);
    } else {
        std::vector<float> scratch(m.N > 0 ? m.N : 1), qf(D), gf(D);
        m.scratch = scratch.data();
        for (int p = 0; p < n; ++p) {
            for (int i = 0; i < D; ++i) qf[i] = (float)q[(size_t)p * D + i];
            logp[p] = b2_eval_model<float, B2HostGroup>(g, m, qf.data(), gf.data(), 0);
            for (int i = 0; i < D; ++i) grad[(size_t)p * D + i] = gf[i];
        }
    }
    return 0;
}

extern "C" void hostsim_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    b2_philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}
extern "C" double hostsim_normal(uint32_t k0, uint32_t k1, uint32_t t, uint32_t i) { return b2_normal(k0, k1, t, i); }
