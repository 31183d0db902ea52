// Classifier head fused with its loss: logits = fc(features) (torch_src/models/mmargcn/agcn.py:198-199) and
// loss = CrossEntropyLoss(logits, label) (mean reduction; torch_src/session/session.py:53, procedures/step.py:41-42) in ONE launch,
// and the whole backward of the pair (d logits, d fc.weight, d fc.bias, d features) in one more.  In eager PyTorch this tail is a
// GEMM, log_softmax, nll_loss and their five backward kernels for a [N, 60] matrix -- pure launch latency, which is what bounds
// the 8-sequences-per-GPU end of the strong-scaling sweep.  Deterministic (fixed-order sums, no atomics).
#include "common.cuh"

namespace agcn {

constexpr int kHeadThreads = 128;

// block n: logits[n][:] = W x[n] + b; softmax; per-sample loss; dlogits[n][:] = (softmax - onehot(label)) / N
__global__ void __launch_bounds__(kHeadThreads)
linear_ce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, const long long* __restrict__ labels,
                     float* __restrict__ logits, float* __restrict__ dlogits, float* __restrict__ loss_n, int n, int cin, int ncls) {
    extern __shared__ float sm[];             // [cin] features | [ncls] logits
    float* xs = sm;
    float* ls = sm + cin;
    const int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < cin; k += kHeadThreads) xs[k] = x[(long long)i * cin + k];
    __syncthreads();
    for (int c = warp; c < ncls; c += kHeadThreads / 32) {
        float acc = 0.f;
        for (int k = lane; k < cin; k += 32) acc = fmaf(w[(long long)c * cin + k], xs[k], acc);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) ls[c] = acc + (b ? b[c] : 0.f);
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int c = lane; c < ncls; c += 32) mx = fmaxf(mx, ls[c]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        float se = 0.f;
        for (int c = lane; c < ncls; c += 32) se += expf(ls[c] - mx);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) se += __shfl_xor_sync(0xffffffffu, se, d);
        const long long lab = labels[i];
        const float lse = mx + logf(se);
        const float inv_n = 1.f / (float)n;
        for (int c = lane; c < ncls; c += 32) {
            const float p = expf(ls[c] - lse);
            logits[(long long)i * ncls + c] = ls[c];
            dlogits[(long long)i * ncls + c] = (p - (c == lab ? 1.f : 0.f)) * inv_n;
        }
        if (lane == 0) loss_n[i] = (lab >= 0 && lab < ncls) ? lse - ls[lab] : NAN;     // out-of-range label: poison the loss, as a debug aid
    }
}

__global__ void mean_kernel(const float* v, float* out, int n) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += v[i];
        *out = s / (float)n;
    }
}

// blocks [0, ncls): dw[c][:] and db[c];  blocks [ncls, ncls + n): dx[i][:].  g = *gscale (upstream gradient of the scalar loss)
__global__ void __launch_bounds__(kHeadThreads)
linear_ce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ dlogits, const float* __restrict__ gscale,
                     float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dx, int n, int cin, int ncls) {
    const float g = gscale ? *gscale : 1.f;
    if ((int)blockIdx.x < ncls) {
        const int c = blockIdx.x;
        for (int k = threadIdx.x; k < cin; k += kHeadThreads) {
            float acc = 0.f;
            for (int i = 0; i < n; ++i) acc = fmaf(dlogits[(long long)i * ncls + c], x[(long long)i * cin + k], acc);
            dw[(long long)c * cin + k] = acc * g;
        }
        if (threadIdx.x == 0 && db != nullptr) {
            float acc = 0.f;
            for (int i = 0; i < n; ++i) acc += dlogits[(long long)i * ncls + c];
            db[c] = acc * g;
        }
    } else if (dx != nullptr) {
        const int i = blockIdx.x - ncls;
        for (int k = threadIdx.x; k < cin; k += kHeadThreads) {
            float acc = 0.f;
            for (int c = 0; c < ncls; ++c) acc = fmaf(dlogits[(long long)i * ncls + c], w[(long long)c * cin + k], acc);
            dx[(long long)i * cin + k] = acc * g;
        }
    }
}

}  // namespace agcn

using namespace agcn;

extern "C" AGCN_API int agcn_linear_ce_fwd(const float* x, const float* w, const float* bias, const long long* labels,
                                           float* logits, float* dlogits, float* loss_per_sample, float* loss,
                                           int n, int cin, int ncls, void* stream) {
    AGCN_REQUIRE(x && w && labels && logits && dlogits && loss_per_sample && loss, AGCN_ERR_NULL, "agcn_linear_ce_fwd: null pointer");
    AGCN_REQUIRE(n > 0 && cin > 0 && ncls > 0 && (size_t)(cin + ncls) * sizeof(float) <= 96 * 1024, AGCN_ERR_BAD_SHAPE,
                 "agcn_linear_ce_fwd: bad shape n=%d cin=%d ncls=%d", n, cin, ncls);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)(cin + ncls) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(linear_ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "agcn_linear_ce_fwd: %s", cudaGetErrorString(e));
    }
    linear_ce_fwd_kernel<<<n, kHeadThreads, smem, s>>>(x, w, bias, labels, logits, dlogits, loss_per_sample, n, cin, ncls);
    int rc = check_launch("agcn_linear_ce_fwd");
    if (rc) return rc;
    mean_kernel<<<1, 32, 0, s>>>(loss_per_sample, loss, n);
    return check_launch("agcn_linear_ce_fwd(mean)");
}

extern "C" AGCN_API int agcn_linear_ce_bwd(const float* x, const float* w, const float* dlogits, const float* grad_loss,
                                           float* dw, float* dbias, float* dx, int n, int cin, int ncls, void* stream) {
    AGCN_REQUIRE(x && w && dlogits && dw, AGCN_ERR_NULL, "agcn_linear_ce_bwd: null pointer");
    AGCN_REQUIRE(n > 0 && cin > 0 && ncls > 0, AGCN_ERR_BAD_SHAPE, "agcn_linear_ce_bwd: bad shape");
    linear_ce_bwd_kernel<<<ncls + (dx ? n : 0), kHeadThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, w, dlogits, grad_loss, dw, dbias, dx, n, cin, ncls);
    return check_launch("agcn_linear_ce_bwd");
}
