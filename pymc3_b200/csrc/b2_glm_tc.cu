// tcgen05 / TMA chain-batched GLM likelihood -- placeholder until the kernel lands.
#include "b2_engine.cuh"
bool b2_glm_tc_supported(const b2_engine* e) { (void)e; return false; }
int b2_glm_tc_launch(b2_engine* e, const float*, const float*, float*, float*, int, const B2ChainState*, int,
                     double*, cudaStream_t) {
    (void)e;
    b2_set_error("tcgen05 GLM kernel not built");
    return -6;
}
