// Tensor-core (tcgen05 / TMEM) path of the StateUpdate edge kernel -- placeholder until the kernel lands.
#include "common.cuh"

namespace pesto {

size_t tc_layer_bytes() { return 0; }
void pack_tc_layer(const float *, void *) {}

int launch_state_update_tc(const float *, const void *, int, int, const int32_t *, const float *, const float *, float *,
                           float *, int mode, cudaStream_t, cudaEvent_t *) {
    set_error("state_update: tensor-core mode %d is not available in this build", mode);
    return PESTO_EINVAL;
}

}  // namespace pesto
