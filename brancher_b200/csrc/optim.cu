// Fused optimiser step over many small parameter tensors: replaces the per-parameter torch.optim launches of
// ProbabilisticOptimizer.update (brancher/optimizers.py:69-73) and the host-side finiteness check + loss bookkeeping of the
// training loop (brancher/inference.py:95-108) with two launches that need no host synchronisation, so that a whole
// iteration (fused ELBO evaluation + this step) can be captured in ONE CUDA graph and replayed.
//
//   brn_opt_step     for every element of every tensor: SGD (momentum, weight decay) or Adam (torch.optim.Adam's update, L2
//                    weight decay) -- skipped when the iteration's loss is not finite, as the reference skips such samples
//   brn_opt_advance  one thread: loss curve entry, step counter, iteration counter, Philox offset of the next iteration
#include "common.cuh"

namespace brn {

__device__ __forceinline__ int find_tensor(const int64_t* __restrict__ prefix, int n, int64_t i) {
    int lo = 0, hi = n - 1;               // prefix[k] = first flat index of tensor k, prefix[n] = total
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (prefix[mid] <= i) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
opt_step_kernel(const brn_opt_tensor* __restrict__ table, const int64_t* __restrict__ prefix, int n_tensors, int64_t total,
                brn_opt_hyper h, const double* __restrict__ loss, const int64_t* __restrict__ counters) {
    const double l = *loss;
    if (!(l - l == 0.0)) return;          // NaN / Inf: skip the update (inference.py:98,106-107)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int k = find_tensor(prefix, n_tensors, i);
    const brn_opt_tensor t = table[k];
    const int64_t j = i - prefix[k];
    float p = t.param[j], g = t.grad[j];
    if (h.weight_decay != 0.f) g = __fmaf_rn(h.weight_decay, p, g);
    const int64_t step = counters[0] + 1;                    // successful steps so far + this one
    if (h.kind == 0) {
        if (h.momentum != 0.f) {
            float b = t.m[j];
            b = step == 1 ? g : __fmaf_rn(h.momentum, b, g);            // torch SGD: buf = g on the first step (dampening 0)
            t.m[j] = b;
            g = b;
        }
        p = __fmaf_rn(-h.lr, g, p);
    } else {
        // torch.optim.Adam (single-tensor path): exp_avg.lerp_(g, 1-b1); exp_avg_sq = b2 v + (1-b2) g g;
        // denom = sqrt(v) / sqrt(1 - b2^t) + eps; p += -(lr / (1 - b1^t)) * m / denom
        float m = t.m[j], v = t.v[j];
        m = __fmaf_rn(g - m, 1.f - h.beta1, m);
        v = __fmaf_rn(g * g, 1.f - h.beta2, v * h.beta2);
        t.m[j] = m;
        t.v[j] = v;
        const double bc1 = 1.0 - pow((double)h.beta1, (double)step), bc2 = 1.0 - pow((double)h.beta2, (double)step);
        const float step_size = (float)((double)h.lr / bc1), bc2_sqrt = (float)sqrt(bc2);
        const float denom = __fsqrt_rn(v) / bc2_sqrt + h.eps;
        p = p - step_size * (m / denom);
    }
    t.param[j] = p;
}

__global__ void opt_advance_kernel(const double* __restrict__ loss, int64_t* __restrict__ counters, float* __restrict__ curve,
                                   int64_t curve_len, unsigned long long* __restrict__ offset_dev) {
    const double l = *loss;
    const int64_t it = counters[1];
    if (curve && it < curve_len) curve[it] = (float)l;
    if (l - l == 0.0) counters[0] += 1;
    else counters[2] += 1;                // skipped iterations
    counters[1] = it + 1;
    if (offset_dev) *offset_dev += 1ull;
}

}  // namespace brn

using namespace brn;

extern "C" int brn_opt_step(const brn_opt_tensor* table_dev, const int64_t* prefix_dev, int n_tensors, int64_t total,
                            const brn_opt_hyper* hyper, const double* loss_dev, int64_t* counters_dev, float* curve_dev,
                            int64_t curve_len, uint64_t* offset_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BRN_CHECK_ARG(table_dev && prefix_dev && hyper && loss_dev && counters_dev, "brn_opt_step: NULL pointer");
    BRN_CHECK_ARG(n_tensors > 0 && total > 0, "brn_opt_step: empty parameter set");
    BRN_CHECK_ARG(hyper->kind == 0 || hyper->kind == 1, "brn_opt_step: unknown optimiser kind %d", hyper->kind);
    opt_step_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(table_dev, prefix_dev, n_tensors, total, *hyper, loss_dev,
                                                                       counters_dev);
    BRN_LAUNCH_OK("opt_step_kernel");
    opt_advance_kernel<<<1, 1, 0, stream>>>(loss_dev, counters_dev, curve_dev, curve_len,
                                            reinterpret_cast<unsigned long long*>(offset_dev));
    BRN_LAUNCH_OK("opt_advance_kernel");
    return 0;
}
