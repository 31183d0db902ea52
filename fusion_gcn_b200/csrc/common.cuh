// Shared helpers for libagcn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include "../../include/agcn_b200.h"

#define AGCN_API __attribute__((visibility("default")))

namespace agcn {

// thread-local error text, returned by agcn_last_error_string()
char* error_buffer();
int fail(int code, const char* fmt, ...);

void count_launch();

inline int check_launch(const char* what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return AGCN_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;   // B200

// Fused tail of a convolution in eval mode (agcn_conv_fwd_post): y = act(scale * (conv + bias) + shift + res)
struct PostOp {
    const float* scale;    // [cout] or NULL (then shift is ignored): BatchNorm with running statistics folded into the epilogue
    const float* shift;
    const float* res;      // tensor of y's shape or NULL: residual branch added before the activation
    int relu;
};

// A/B switches and limiter probes exist only in builds made with -DAGCN_PROBES
// (python -m fusion_gcn_b200.build --probes); the default library never reads the environment.
#ifdef AGCN_PROBES
inline const char* probe_env(const char* name) { return getenv(name); }
#else
inline const char* probe_env(const char*) { return nullptr; }
#endif

}  // namespace agcn

#define AGCN_REQUIRE(cond, code, ...) do { if (!(cond)) return agcn::fail(code, __VA_ARGS__); } while (0)
