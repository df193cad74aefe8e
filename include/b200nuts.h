/* b200nuts.h -- C ABI of the B200-native NUTS/HMC engine (libb200nuts.so).
 *
 * Every entry point replaces one seam of the reference's (PyMC3 v3.8) sampler hot path.
 * The reference has no FFI of its own (it is pure Python over Theano), so the "binding" is
 * the ctypes stub shown in INTEGRATION.md; the Python classes in pymc3_b200/ keep the
 * reference's step-method / trace API on top of these calls.
 *
 * Conventions
 *   - all pointers named d_* are DEVICE pointers owned by the caller (allocated e.g. with
 *     PyTorch); the engine allocates only its private chain state, freed by b2_engine_destroy.
 *   - every call returns 0 on success, <0 for usage errors, >0 for a CUDA error code;
 *     b2_last_error() returns a thread-local message.
 *   - numerical events (divergences, non-finite initial energy, tree-depth hits) are DATA,
 *     reported through the trace arrays / b2_chain_report, never error codes.
 *   - dtype: vectors (positions, momenta, gradients, traces) are float (B2_F32, production)
 *     or double (B2_F64, check build); energies, log-densities and all scalar sampler
 *     arithmetic are double in both.
 */
#ifndef B200NUTS_H
#define B200NUTS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B2_ABI_VERSION 4

enum b2_dtype { B2_F32 = 0, B2_F64 = 1 };

/* model families: the Theano graph of pymc3/model.py:622-631 is replaced by a family id +
 * data pointers.  Free-variable order is creation order (blocking.py:33-59). */
enum b2_family {
    B2_STD_NORMAL = 0,          /* x[D] ~ Normal(mu, sigma)                                  */
    B2_EIGHT_SCHOOLS_NCP = 2,   /* pymc3/examples/gelman_schools.py:26-40                    */
    B2_GLM_LOGIT = 3,           /* pymc3/glm/linear.py:49-101 + glm/families.py:115-119      */
    B2_HIER_LINEAR_NCP = 4,     /* benchmarks/benchmarks/benchmarks.py:25-45                 */
    B2_STOCH_VOL = 5            /* docs/source/notebooks/stochastic_volatility.ipynb cell 10 */
};

typedef struct b2_model_desc {
    int32_t family;
    int32_t D;                  /* free parameters in the sampler's (unconstrained) space   */
    int32_t N;                  /* observations (J for eight schools, T for stoch. vol.)     */
    int32_t G;                  /* groups (hier) / regressors K without intercept (glm)      */
    const double* d_aux0;       /* [D] mu | [J] y | [T] returns                              */
    const double* d_aux1;       /* [D] sigma | [J] sigma                                     */
    const float* d_X;           /* glm: [N, K] row-major fp32 design matrix (no ones column) */
    const float* d_y;           /* glm: [N] 0/1 ; hier: [N] response, sorted by group        */
    const uint8_t* d_floor;     /* hier: [N] 0/1 covariate, sorted by group                  */
    const int32_t* d_grp_off;   /* hier: [G+1] CSR offsets of the group-sorted observations  */
    double hp[4];               /* family hyper-parameters, see pymc3_b200/csrc/b2_models.cuh */
} b2_model_desc;

enum b2_kind { B2_NUTS = 0, B2_HMC = 1 };
enum b2_exec { B2_EXEC_AUTO = 0, B2_EXEC_PERSISTENT = 1, B2_EXEC_LOCKSTEP = 2 };
/* likelihood kernel of B2_GLM_LOGIT: GROUP = one warp/block per chain (small N); SIMT = chain-batched FFMA/DFMA
   tiles (fp64 check build, any shape); TCGEN05 = tensor cores, fp32 build only: up to 127 features fused with the
   lock-step companion kernel (b2_glm_tc.cu), 128..256 features the wide variant (b2_glm_tcw.cu).  AUTO picks
   TCGEN05 where it applies. */
enum b2_glm_path { B2_GLM_AUTO = 0, B2_GLM_GROUP = 1, B2_GLM_SIMT = 2, B2_GLM_TCGEN05 = 3 };

/* ctor arguments of NUTS / HamiltonianMC / BaseHMC (nuts.py:105, hmc.py:53, base_hmc.py:41-59) */
typedef struct b2_sampler_opts {
    int32_t kind;               /* b2_kind                                                   */
    int32_t n_iters;            /* transitions to run in this call (tune + draws)            */
    int32_t tune_until;         /* absolute iteration index at which tuning stops
                                   (sampling.py:919-920, `if i == tune: stop_tuning`)        */
    int32_t max_treedepth;      /* nuts.py:105 (default 10)                                  */
    int32_t early_max_treedepth;/* nuts.py:105 (default 8; first 200 tuning iterations)      */
    double Emax;                /* base_hmc.py:49 (1000)                                     */
    double target_accept;       /* 0.8 NUTS / 0.65 HMC                                       */
    double gamma, k, t0;        /* step_sizes.py:22 (0.05, 0.75, 10)                         */
    int32_t adapt_step_size;    /* base_hmc.py:54                                            */
    int32_t adapt_mass;         /* 1: QuadPotentialDiagAdapt  0: static QuadPotentialDiag    */
    double path_length;         /* hmc.py:53 (2.0)                                           */
    int32_t max_steps;          /* hmc.py:53 (1024)                                          */
    int32_t hmc_jitter;         /* 1: step_rand = unif (hmc.py:26-27,104)                    */
    int32_t exec_mode;          /* b2_exec                                                   */
    int32_t glm_path;           /* b2_glm_path                                               */
    int32_t run_ahead;          /* lock-step runs only: a chain that has finished this call's n_iters may go on
                                   with up to run_ahead more transitions while slower chains catch up (the call
                                   still returns as soon as EVERY chain has done n_iters); the trace buffers must
                                   then hold n_iters + run_ahead rows.  0 = every chain stops at n_iters.        */
} b2_sampler_opts;

/* device trace: replaces NDArray.record per draw (backends/ndarray.py:258-277).
 * Row r of every array belongs to iteration (first iteration of this call + r).
 * Any pointer may be NULL.  Stat names/dtypes follow nuts.py:91-103 and hmc.py:39-51. */
typedef struct b2_trace_out {
    void* d_q;                  /* [n_iters, C, D]  dtype of the engine                      */
    double* d_energy;           /* [n_iters, C] ...                                          */
    double* d_energy_error;
    double* d_max_energy_error; /* NUTS */
    double* d_mean_tree_accept; /* NUTS */
    double* d_step_size;
    double* d_step_size_bar;
    double* d_model_logp;
    double* d_accept;           /* HMC */
    int32_t* d_depth;           /* NUTS */
    int32_t* d_tree_size;       /* NUTS */
    int32_t* d_n_steps;         /* HMC */
    uint8_t* d_diverging;
    uint8_t* d_tune;
    uint8_t* d_accepted;        /* HMC */
} b2_trace_out;

/* per-chain summary after a run (host memory) */
typedef struct b2_chain_report {
    int32_t phase;              /* 3 = done, 4 = failed                                      */
    int32_t fail_code;          /* 1 = bad initial energy (base_hmc.py:138-158)              */
    int32_t iter;               /* iterations completed                                      */
    int32_t n_div_post;         /* divergences after tuning (base_hmc.py:178)                */
    int32_t n_maxdepth_post;    /* nuts.py:182-184                                           */
    int32_t n_post;             /* samples after tuning                                      */
    int64_t n_grad;             /* gradient evaluations performed                            */
    double step_size;           /* exp(log_step)                                             */
    double step_size_bar;       /* exp(log_bar)                                              */
} b2_chain_report;

typedef struct b2_engine b2_engine;

int b2_abi_version(void);
const char* b2_last_error(void);

/* replaces GradientSharedStep.__init__ -> model.logp_dlogp_function (arraystep.py:243-254):
 * binds a model family + data to `n_chains` chains on `device`. */
int b2_engine_create(const b2_model_desc* desc, int32_t n_chains, int32_t dtype, int32_t device,
                     b2_engine** out);
int b2_engine_destroy(b2_engine* e);

/* replaces ValueGradFunction.__call__ (model.py:645-666), batched over n_points <= n_chains:
 * d_q [n_points, D] (engine dtype) -> d_logp [n_points] (double), d_grad [n_points, D]. */
int b2_logp_dlogp(b2_engine* e, const void* d_q, int32_t n_points, double* d_logp, void* d_grad,
                  int32_t glm_path, void* stream);

/* replaces CpuLeapfrogIntegrator.compute_state followed by n_steps x CpuLeapfrogIntegrator.step(epsilon, state)
 * (integration.py:39-47, 49-109) with a static diagonal potential (QuadPotentialDiag.velocity / .energy,
 * quadpotential.py:356-397), for all n_chains chains at once: d_q, d_p [C, D] (engine dtype) -> d_q_out, d_p_out
 * [C, D] and d_energy_out [C] (may be NULL) = kinetic - logp of the end state; d_var [D] is the diagonal of M^-1
 * (velocity = var * p).  epsilon may be negative (time reversal, tests/test_hmc.py:27-46).  Uses the engine's
 * edge slots and mass diagonal as scratch: call b2_set_state again before sampling. */
int b2_leapfrog(b2_engine* e, const void* d_q, const void* d_p, const double* d_var, double epsilon,
                int32_t n_steps, void* d_q_out, void* d_p_out, double* d_energy_out, int32_t glm_path, void* stream);

/* replaces QuadPotentialFull / QuadPotentialFullInv (quadpotential.py:400-479: velocity = cov p, energy = p.cov p / 2,
 * random = chol^-T n) for every chain of the engine.  d_chol [D, D] row-major (engine dtype) is the LOWER Cholesky
 * factor L of the covariance (cov = L L^T); it is copied, the pointer is not retained.  From then on the state
 * machine integrates z = L^-1 q with unit mass -- the same Hamiltonian flow, U-turn products and energies as the
 * dense metric on q (p_z = L^T p_q) -- and evaluates the density at q = L z, mapping its gradient back with L^T:
 * positions handed to b2_set_state / b2_set_position and returned in traces / b2_get_position are z (the host side
 * multiplies by L).  Runs become lock-step, mass adaptation must be off (adapt_mass = 0), mass var = 1.
 * d_chol = NULL switches the dense metric off.  Not available in the stepwise (b2_step_*) run. */
int b2_set_dense_mass(b2_engine* e, const void* d_chol, void* stream);

/* replaces per-chain seeding + start points + init_nuts' potential
 * (sampling.py:410-413, 883-884, 1915-1929; base_hmc.py:93-103):
 * d_q0 [C, D] (engine dtype), d_seeds [C], mass mean/var [D] shared by all chains. */
int b2_set_state(b2_engine* e, const void* d_q0, const uint64_t* d_seeds, double step_size0,
                 const double* d_mass_mean, const double* d_mass_var, double mass_weight,
                 int32_t adaptation_window, void* stream);

/* replaces `point = step.step(point)` being handed a point other than the chain's last
 * state (sampling.py:921, arraystep.py:258-264; CompoundStep / user-driven loops): moves every
 * chain to d_q [C, D] keeping all adaptation state; the gradient is re-evaluated there. */
int b2_set_position(b2_engine* e, const void* d_q, void* stream);

/* replaces the draw loop _iter_sample / _mp_sample (sampling.py:914-936, 1305-1414) for all
 * chains at once: BaseHMC.astep + NUTS/HamiltonianMC._hamiltonian_step + adaptation + record. */
int b2_sample_run(b2_engine* e, const b2_sampler_opts* opts, const b2_trace_out* trace, void* stream);

/* Stepwise form of b2_sample_run for the observation-sharded configuration (BASELINE.json config 5,
 * SURVEY 8e): the reference has no counterpart (it never shards a likelihood); the seam is the same
 * ValueGradFunction.__call__ (model.py:645-666) whose result becomes a sum over ranks.  Every rank
 * holds a row shard of the data and ALL chains; per leapfrog
 *     b2_step_likelihood  -> d_packed[C, D+1] fp64 = (partial logp, partial dlogp) of each pending position
 *     all-reduce(sum) of d_packed over ranks (NCCL; done by the caller)
 *     b2_step_advance     -> unpacks the reduced values (prior counted once: pass prior_copies = world
 *                            size) and advances every chain's NUTS/HMC state machine by one leapfrog.
 * b2_step_active returns how many chains still need gradients; b2_step_end closes the run. */
int b2_step_begin(b2_engine* e, const b2_sampler_opts* opts, const b2_trace_out* trace, void* stream);
int b2_step_likelihood(b2_engine* e, double* d_packed, void* stream);
int b2_step_advance(b2_engine* e, const double* d_packed, int32_t prior_copies, void* stream);
int b2_step_active(b2_engine* e, int32_t* host_count, void* stream);
int b2_step_end(b2_engine* e);

/* adaptation / bookkeeping state back to the host (step.step_size, potential._var, warnings) */
int b2_get_chain_reports(b2_engine* e, b2_chain_report* host_out /* [C] */);
int b2_get_mass_var(b2_engine* e, double* host_out /* [C, D] */);
int b2_get_position(b2_engine* e, double* host_out /* [C, D] */);

/* how many of this library's kernels were launched by the engine so far */
int64_t b2_kernel_launches(b2_engine* e);

/* live CUDA-event timing of the chain-batched likelihood launches in lock-step mode (the
 * reference's analogue is theano profiling, model.py:668-671): total ms and launch count
 * since b2_set_profiling(e, 1). */
int b2_set_profiling(b2_engine* e, int32_t on);
int b2_get_profile(b2_engine* e, double* likelihood_ms, int64_t* likelihood_launches);
int b2_get_profile_advance(b2_engine* e, double* advance_ms);   /* total ms of the advance kernel over the same launches */

#ifdef __cplusplus
}
#endif
#endif
