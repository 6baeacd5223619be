// One-shot all-reduce over NVLink / NVSwitch peer memory for the small variational-parameter gradient (636 KB at
// C3, 257 floats at C2 -- latency-bound, not bandwidth-bound).  Every rank copies its partial gradient into a symmetric
// buffer, raises a flag in every peer's flag array, waits for its own flags and then sums the peers' buffers straight out
// of their memory (loads over NVLink), in rank order -- so every rank obtains bit-identical sums.  One barrier per
// reduction: the buffers are double-buffered by the parity of an epoch counter that lives on the device, so the sequence
// can be captured in a CUDA graph and replayed (a rank can only pass the barrier of epoch e + 1 after every peer finished
// reading epoch e, hence nobody overwrites a buffer that is still being read).
// The reference has no collective at all (single process); this replaces what `ncclAllReduce` would do in SURVEY's proposal.
#include "common.cuh"
#include <algorithm>

namespace brn {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// step 1: src -> this rank's symmetric buffer of the coming epoch's parity
// loss != NULL: the fp64 partial loss rides in the buffer's last quad as a (hi, lo) fp32 pair (elements n-4, n-3; the
// caller's src keeps that quad spare) and comes back summed -- no separate collective, no packing kernels.
__global__ void __launch_bounds__(256)
allreduce_stage_kernel(const float* __restrict__ src, float* const* __restrict__ bufs, int rank, int world, int64_t n,
                       const unsigned long long* __restrict__ epoch, const double* __restrict__ loss) {
    const unsigned long long e = *epoch + 1ull;
    float* dst = bufs[(e & 1ull) * world + rank];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = src[i];
        if (loss && i >= n - 4) {
            const double l = *loss;
            const float hi = (float)l;
            v = i == n - 4 ? hi : (i == n - 3 ? (float)(l - (double)hi) : 0.f);
        }
        dst[i] = v;
    }
}

// step 2: flags, wait, reduce.  status[0] counts wait time-outs (a dead peer must not hang the GPU).
__global__ void __launch_bounds__(256)
allreduce_reduce_kernel(float* __restrict__ out, float* const* __restrict__ bufs, unsigned long long* const* __restrict__ flags,
                        int rank, int world, int64_t n, unsigned long long* __restrict__ epoch, unsigned int* __restrict__ ticket,
                        unsigned int* __restrict__ status, double* __restrict__ loss) {
    const unsigned long long e = *epoch + 1ull;
    const int par = (int)(e & 1ull);
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
        __threadfence_system();                                   // the staged copy (previous kernel) is visible before the flag
        st_release_sys(flags[threadIdx.x] + rank, e);              // peer threadIdx.x: "rank's buffer of epoch e is complete"
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const unsigned long long* f = flags[rank] + threadIdx.x;
        long long spins = 0;
        while (ld_acquire_sys(f) < e) {
            if (++spins > (1ll << 24)) { timed_out = 1; break; }   // ~ seconds: give up instead of hanging the device
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (timed_out && threadIdx.x == 0) atomicAdd(status, 1u);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n / 4;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < world; ++j) {
            const float4 v = __ldcv(reinterpret_cast<const float4*>(bufs[par * world + j]) + q);     // never from a stale L1 line
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[q] = acc;
        if (loss && q == n4 - 1) *loss = (double)acc.x + (double)acc.y;       // n % 4 == 0 when a loss rides along
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float acc = 0.f;
        for (int j = 0; j < world; ++j) acc += __ldcv(bufs[par * world + j] + i);
        out[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {              // last block: publish the epoch for the next launch
            *ticket = 0u;
            *epoch = e;
        }
    }
}

}  // namespace brn

using namespace brn;

extern "C" int brn_allreduce_oneshot(const float* src, float* out, int64_t n, float* const* bufs_dev,
                                     unsigned long long* const* flags_dev, int rank, int world, uint64_t* state_dev,
                                     double* loss_inout, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BRN_CHECK_ARG(src && out && bufs_dev && flags_dev && state_dev, "brn_allreduce_oneshot: NULL pointer");
    BRN_CHECK_ARG(n > 0 && world >= 1 && world <= 64 && rank >= 0 && rank < world, "brn_allreduce_oneshot: bad arguments n=%lld rank=%d world=%d",
                  (long long)n, rank, world);
    BRN_CHECK_ARG(((uintptr_t)out & 15) == 0, "brn_allreduce_oneshot: out must be 16-byte aligned");
    BRN_CHECK_ARG(!loss_inout || (n % 4 == 0 && n >= 4), "brn_allreduce_oneshot: a loss needs n %% 4 == 0 (its (hi, lo) pair uses the last quad)");
    unsigned long long* epoch = reinterpret_cast<unsigned long long*>(state_dev);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(state_dev + 1);
    unsigned int* status = ticket + 1;
    const unsigned grid = (unsigned)std::min<int64_t>((n / 4 + 255) / 256 + 1, 148);
    allreduce_stage_kernel<<<grid, 256, 0, stream>>>(src, bufs_dev, rank, world, n, epoch, loss_inout);
    BRN_LAUNCH_OK("allreduce_stage_kernel");
    allreduce_reduce_kernel<<<grid, 256, 0, stream>>>(out, bufs_dev, flags_dev, rank, world, n, epoch, ticket, status, loss_inout);
    BRN_LAUNCH_OK("allreduce_reduce_kernel");
    return 0;
}
