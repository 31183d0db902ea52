// tcgen05 / TMA tensor-core path of the implicit GEMM (AGCN_PREC_TF32).  Placeholder until the kernel lands:
// reports "unsupported" so that agcn_conv_fwd falls back to the FFMA kernel.
#include "common.cuh"

int agcn_conv_fwd_tc(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, void*) {
    return AGCN_ERR_UNSUPPORTED;
}
