// Engine object shared by the .cu translation units of libb200nuts.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/b200nuts.h"
#include "b2_models.cuh"

struct b2_engine {
    b2_model_desc desc;
    B2ModelData md;
    int C, D, Dp, dtype, device;
    void* vec;                 // [B2_NUM_VEC_SLOTS][C][Dp] of dtype
    double* wv_mean;           // [2][C][Dp]
    double* wv_m2;
    B2ChainState* st;          // [C]
    double* lv;                // [C][4][B2_MAX_LEVELS] stack-buffer scalars
    double* logp_eval;         // [C]
    void* glm_scratch;         // lazily allocated [C][N] for the group evaluator
    int* d_active;             // device counter
    int* h_active;             // pinned host mirror
    void* glm_ws;              // workspace of the chain-batched GLM kernels (b2_glm_*.cu)
    size_t glm_ws_bytes;
    void* glm_tc;              // state of the tcgen05 GLM path (b2_glm_tc.cu), or null
    void* glm_tcw;             // state of the wide (128 <= K <= 256) tcgen05 GLM path (b2_glm_tcw.cu), or null
    void* hier_ws;             // workspace of the chain-batched hierarchical kernel
    size_t hier_ws_bytes;
    int iter_done;             // iterations completed by every chain so far
    bool state_set;
    int64_t launches;
    int sm_count;
    // stepwise lock-step run (observation sharding): options / trace of the run in progress
    b2_sampler_opts step_opts;
    b2_trace_out step_trace;
    bool stepping;
    // optional live timing of the likelihood launches (bench.py roofline)
    int profile;
    double like_ms;
    int64_t like_n;
    double adv_ms;             // advance / post kernel, same launches
    cudaEvent_t ev[96];        // per batch step b: ev[3b] | likelihood | ev[3b+1] | advance | ev[3b+2]
    // dense mass matrix (b2_set_dense_mass): the state machine runs in z = L^-1 q with unit mass; L is the lower
    // Cholesky factor of the covariance.  dense_L [D][D] row-major, dense_Lt its transpose, dense_q / dense_g [C][Dp].
    void *dense_L, *dense_Lt, *dense_q, *dense_g;
    cudaStream_t own_stream;   // b2_sample_run's stream when the caller hands over the (uncapturable) default stream
    cudaEvent_t own_event;
};

void b2_set_error(const std::string& msg);
#define B2_CUDA_OK(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            b2_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
            return (int)_e;                                                                \
        }                                                                                  \
    } while (0)

// chain-batched likelihood kernels (one launch evaluates every chain's pending position)
// q/g planes: edge 0 at plane pointers A, edge 1 at B; st selects per chain (null -> plane A).
template <typename T>
int b2_glm_simt_launch(b2_engine* e, const T* qA, const T* qB, T* gA, T* gB, int ld,
                       const B2ChainState* st, int n_points, double* logp, cudaStream_t stream);
int b2_glm_tc_launch(b2_engine* e, const float* qA, const float* qB, float* gA, float* gB, int ld,
                     const B2ChainState* st, int n_points, double* logp, cudaStream_t stream);
bool b2_glm_tc_supported(const b2_engine* e);
// wide variant (128 <= K <= 256 features, b2_glm_tcw.cu): likelihood launch only, no fused companion kernel
bool b2_glm_tcw_supported(const b2_engine* e);
void b2_glm_tcw_release(b2_engine* e);
// start of a lock-step / stepwise run: reference position q_ref = mean live position, eta_ref = X . q_ref
int b2_glm_tcw_refresh(b2_engine* e, const float* qA, const float* qB, int ld, const B2ChainState* st, int n, cudaStream_t stream);
int b2_glm_tcw_launch(b2_engine* e, const float* qA, const float* qB, float* gA, float* gB, int ld,
                      const B2ChainState* st, int n, double* logp, cudaStream_t stream);
void b2_glm_tc_release(b2_engine* e);
// lock-step pieces of the tensor-core path: slot maps at the start of a run, then one call per leapfrog
int b2_glm_tc_pack(b2_engine* e, const float* qA, const float* qB, int ld, const B2ChainState* st, int n, cudaStream_t s);
// one leapfrog of every live chain: fused schedule = 2 launches {likelihood(half h) + state machine(other half)},
// two-kernel schedule (B2_TC_FUSED=0) = likelihood + state-machine kernel; `mid` (optional) is recorded between
// the likelihood and the stand-alone state-machine kernel (after both launches in the fused schedule)
int b2_glm_tc_step(b2_engine* e, const void* view_f32, cudaStream_t s, cudaEvent_t mid);
bool b2_glm_tc_is_fused(const b2_engine* e);
template <typename T>
int b2_hier_launch(b2_engine* e, const T* qA, const T* qB, T* gA, T* gB, int ld,
                   const B2ChainState* st, int n_points, double* logp, cudaStream_t stream);
