// Multi-tensor optimizer steps: ONE launch updates every parameter tensor of a group.
//
// The reference builds torch.optim.{SGD, Adam, AdamW} over model.parameters() (torch_src/session_helper.py:48-82) and steps them
// after every batch (torch_src/session/procedures/step.py:48-49), under GradScaler when --mixed_precision is set (:67-71).  The
// AGCN model has 274 parameter tensors, most of them a few hundred floats: per-tensor kernels leave hundreds of launches of a few
// microseconds behind the 60 ms step.  Here the tensors stay where PyTorch put them (state-dict / checkpoint layout untouched); a
// device table of (param, grad, state1, state2, numel) rows and a list of (tensor, chunk) work items let one grid walk all of them.
// GradScaler: torch hands optimizers that declare _step_supports_amp_scaling the device scalars `grad_scale` and `found_inf`; the
// unscale (g / scale) and the skip-on-overflow are done here, so no host synchronisation is needed and the step is graph-capturable.
// Arithmetic follows torch/optim/sgd.py (_single_tensor_sgd) and torch/optim/adam.py (_single_tensor_adam, capturable form).
#include "common.cuh"

namespace agcn {

constexpr int kOptChunk = 4096;          // elements per work item
constexpr int kOptThreads = 256;

struct OptRow { float* p; const float* g; float* s1; float* s2; long long n; };

// hyper-parameters arrive as doubles (Python floats) and are rounded to fp32 only where torch rounds them: 1 - beta and the bias
// corrections are formed in double first (1.f - 0.999f is 1.3e-5 away from float(1 - 0.999))
struct SgdHyper { double lr, momentum, dampening, weight_decay; int nesterov, first; };
struct AdamHyper { double lr, beta1, beta2, eps, weight_decay; int decoupled; };

template <typename F>
__device__ __forceinline__ void for_chunk(const OptRow& t, long long base, F f) {
    const long long end = (base + kOptChunk < t.n) ? base + kOptChunk : t.n;
    for (long long i = base + threadIdx.x; i < end; i += kOptThreads) f(i);
}

__global__ void __launch_bounds__(kOptThreads)
sgd_kernel(const OptRow* __restrict__ table, const int2* __restrict__ items, SgdHyper h, const float* __restrict__ lr_dev,
           const float* __restrict__ grad_scale, const float* __restrict__ found_inf) {
    if (found_inf != nullptr && *found_inf != 0.f) return;          // GradScaler: skip the whole step on overflow
    const int2 it = items[blockIdx.x];
    const OptRow t = table[it.x];
    const float lr = lr_dev != nullptr ? *lr_dev : (float)h.lr;
    const float inv_scale = grad_scale != nullptr ? 1.f / *grad_scale : 1.f;
    const float wd = (float)h.weight_decay, mom = (float)h.momentum, damp1 = (float)(1.0 - h.dampening);
    for_chunk(t, (long long)it.y * kOptChunk, [&](long long i) {
        const float p = t.p[i];
        float g = t.g[i] * inv_scale;
        if (wd != 0.f) g = fmaf(wd, p, g);
        if (mom != 0.f) {
            float buf = h.first ? g : fmaf(mom, t.s1[i], damp1 * g);
            t.s1[i] = buf;
            g = h.nesterov ? fmaf(mom, buf, g) : buf;
        }
        t.p[i] = fmaf(-lr, g, p);
    });
}

__global__ void __launch_bounds__(kOptThreads)
adam_kernel(const OptRow* __restrict__ table, const int2* __restrict__ items, AdamHyper h, const float* __restrict__ lr_dev,
            const float* __restrict__ step_dev, const float* __restrict__ grad_scale, const float* __restrict__ found_inf) {
    if (found_inf != nullptr && *found_inf != 0.f) return;
    const int2 it = items[blockIdx.x];
    const OptRow t = table[it.x];
    const double lr_d = lr_dev != nullptr ? (double)*lr_dev : h.lr;
    const float inv_scale = grad_scale != nullptr ? 1.f / *grad_scale : 1.f;
    const double step = (double)*step_dev + 1.0;                     // steps completed so far + this one
    const float step_size = (float)(lr_d / (1.0 - pow(h.beta1, step)));
    const float bc2_sqrt = (float)sqrt(1.0 - pow(h.beta2, step));
    const float b1 = (float)h.beta1, b2 = (float)h.beta2, omb1 = (float)(1.0 - h.beta1), omb2 = (float)(1.0 - h.beta2);
    const float wd = (float)h.weight_decay, eps = (float)h.eps, decay = (float)(1.0 - lr_d * h.weight_decay);
    for_chunk(t, (long long)it.y * kOptChunk, [&](long long i) {
        float p = t.p[i];
        float g = t.g[i] * inv_scale;
        if (h.decoupled) p *= decay;                                 // AdamW
        else if (wd != 0.f) g = fmaf(wd, p, g);
        const float m0 = t.s1[i];
        const float m = fmaf(omb1, g - m0, m0);                      // exp_avg.lerp_(grad, 1 - beta1)
        const float v = fmaf(omb2 * g, g, b2 * t.s2[i]);             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        t.s1[i] = m;
        t.s2[i] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        t.p[i] = fmaf(-step_size, m / denom, p);
    });
    (void)b1;
}

// flat[off ...] <-> tensor copies for the gradient buckets of the data-parallel all-reduce: rows { flat slice, tensor, -, -, numel }
// (OptRow with p = flat slice, g = tensor).  to_flat: flat = tensor;  else: tensor = flat * scale (the 1 / world averaging).
__global__ void __launch_bounds__(kOptThreads)
bucket_copy_kernel(const OptRow* __restrict__ table, const int2* __restrict__ items, int to_flat, float scale) {
    const int2 it = items[blockIdx.x];
    const OptRow t = table[it.x];
    float* tensor = const_cast<float*>(t.g);
    for_chunk(t, (long long)it.y * kOptChunk, [&](long long i) {
        if (to_flat) t.p[i] = tensor[i];
        else tensor[i] = t.p[i] * scale;
    });
}

static int check_table(const void* table, const int* items, int nitems, const char* what) {
    AGCN_REQUIRE(table && items, AGCN_ERR_NULL, "%s: null table / work list", what);
    AGCN_REQUIRE(nitems > 0, AGCN_ERR_BAD_SHAPE, "%s: empty work list", what);
    AGCN_REQUIRE((reinterpret_cast<uintptr_t>(table) & 7u) == 0 && (reinterpret_cast<uintptr_t>(items) & 7u) == 0, AGCN_ERR_MISALIGNED,
                 "%s: table / work list not 8-byte aligned", what);
    return AGCN_OK;
}

}  // namespace agcn

using namespace agcn;

extern "C" AGCN_API int agcn_optim_chunk(void) { return kOptChunk; }

extern "C" AGCN_API int agcn_optim_sgd(const void* table, const int* items, int nitems, double lr, const float* lr_dev,
                                       double momentum, double dampening, double weight_decay, int nesterov, int first_step,
                                       const float* grad_scale, const float* found_inf, void* stream) {
    int rc = check_table(table, items, nitems, "agcn_optim_sgd");
    if (rc) return rc;
    AGCN_REQUIRE(!nesterov || (momentum > 0.0 && dampening == 0.0), AGCN_ERR_UNSUPPORTED,
                 "agcn_optim_sgd: Nesterov momentum requires a momentum and zero dampening");       // torch/optim/sgd.py raises the same
    SgdHyper h{lr, momentum, dampening, weight_decay, nesterov, first_step};
    sgd_kernel<<<nitems, kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const OptRow*>(table), reinterpret_cast<const int2*>(items),
                                                                              h, lr_dev, grad_scale, found_inf);
    return check_launch("agcn_optim_sgd");
}

extern "C" AGCN_API int agcn_optim_adam(const void* table, const int* items, int nitems, double lr, const float* lr_dev,
                                        double beta1, double beta2, double eps, double weight_decay, int decoupled,
                                        const float* step_dev, const float* grad_scale, const float* found_inf, void* stream) {
    int rc = check_table(table, items, nitems, "agcn_optim_adam");
    if (rc) return rc;
    AGCN_REQUIRE(step_dev != nullptr, AGCN_ERR_NULL, "agcn_optim_adam: null step counter");
    AdamHyper h{lr, beta1, beta2, eps, weight_decay, decoupled};
    adam_kernel<<<nitems, kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const OptRow*>(table), reinterpret_cast<const int2*>(items),
                                                                               h, lr_dev, step_dev, grad_scale, found_inf);
    return check_launch("agcn_optim_adam");
}

extern "C" AGCN_API int agcn_bucket_copy(const void* table, const int* items, int nitems, int to_flat, float scale, void* stream) {
    int rc = check_table(table, items, nitems, "agcn_bucket_copy");
    if (rc) return rc;
    bucket_copy_kernel<<<nitems, kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const OptRow*>(table), reinterpret_cast<const int2*>(items),
                                                                                 to_flat, scale);
    return check_launch("agcn_bucket_copy");
}
