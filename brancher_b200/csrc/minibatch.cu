// Device-side minibatch index sampling WITHOUT replacement -- see include/brancher_cuda.h
// (brn_minibatch_indices).  Replaces the host-side `np.random.choice(range(N), B, replace=False)` of
// EmpiricalDistribution._get_sample (brancher/distributions.py:410-462), which permutes all N row ids on the host for every
// iteration (~140 ms at N = 10^6) and would dwarf any fused ELBO evaluation.
//
// One CTA.  Slot i of the minibatch draws a candidate row uniformly from Philox(seed, offset; round, i); duplicates are
// resolved with the rule "the lower slot keeps the value, the higher slot redraws in the next round", which is exactly
// sequential rejection sampling (slot i's value is uniform over the rows no lower slot holds), hence a uniformly random
// B-subset in random order, and -- unlike a first-come hash insertion -- independent of thread scheduling: the result is a
// pure function of (N, B, seed, offset).  An open-addressing table in shared memory maps value -> lowest slot holding it.
#include "common.cuh"

namespace brn {

constexpr int MB_THREADS = 1024;
constexpr unsigned long long MB_EMPTY = ~0ull;

__device__ __forceinline__ uint32_t mb_hash(uint32_t v) {
    v ^= v >> 16; v *= 0x7feb352du; v ^= v >> 15; v *= 0x846ca68bu; v ^= v >> 16;
    return v;
}

// uniform integer in [0, N) from 64 random bits (multiply-shift; bias < N / 2^64)
__device__ __forceinline__ int64_t mb_draw(uint64_t seed, uint64_t offset, uint32_t round, uint32_t slot, int64_t N) {
    const Philox4 p = philox4x32_10(slot, round, 0x6d62u /* "mb" stream */, (uint32_t)offset, (uint32_t)seed,
                                    (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32));
    const unsigned long long r = ((unsigned long long)p.x << 32) | p.y;
    return (int64_t)__umul64hi(r, (unsigned long long)N);
}

__global__ void __launch_bounds__(MB_THREADS)
minibatch_indices_kernel(int64_t N, int B, uint64_t seed, uint64_t offset, int table_size, int64_t* __restrict__ out,
                         int* __restrict__ rounds_out) {
    extern __shared__ unsigned long long table[];          // (value << 20 | lowest slot) per entry; B <= 2^20
    __shared__ int pending;
    const int t = threadIdx.x;
    // value / round of every slot this thread owns (slots t, t + 1024, ...) live in global `out` (value) and registers
    // cannot hold an unbounded count, so the per-slot round counter is recomputed: a slot redraws at most once per round
    for (int i = t; i < B; i += MB_THREADS) out[i] = mb_draw(seed, offset, 0, (uint32_t)i, N);
    int round = 0;
    while (true) {
        for (int e = t; e < table_size; e += MB_THREADS) table[e] = MB_EMPTY;
        if (t == 0) pending = 0;
        __syncthreads();
        // insert: entry of value v ends up holding the LOWEST slot that has v
        for (int i = t; i < B; i += MB_THREADS) {
            const unsigned long long v = (unsigned long long)out[i];
            const unsigned long long packed = (v << 20) | (unsigned long long)i;
            uint32_t h = mb_hash((uint32_t)v ^ (uint32_t)(v >> 32)) & (uint32_t)(table_size - 1);
            while (true) {
                const unsigned long long cur = table[h];
                if (cur == MB_EMPTY) {
                    const unsigned long long old = atomicCAS(&table[h], MB_EMPTY, packed);
                    if (old == MB_EMPTY) break;
                    if ((old >> 20) == v) { atomicMin(&table[h], packed); break; }
                } else if ((cur >> 20) == v) {
                    atomicMin(&table[h], packed);
                    break;
                }
                h = (h + 1) & (uint32_t)(table_size - 1);
            }
        }
        __syncthreads();
        // losers (a lower slot holds the same value) redraw with the next round's counter
        for (int i = t; i < B; i += MB_THREADS) {
            const unsigned long long v = (unsigned long long)out[i];
            uint32_t h = mb_hash((uint32_t)v ^ (uint32_t)(v >> 32)) & (uint32_t)(table_size - 1);
            while ((table[h] >> 20) != v) h = (h + 1) & (uint32_t)(table_size - 1);
            if ((int)(table[h] & 0xfffffull) != i) {
                out[i] = mb_draw(seed, offset, (uint32_t)(round + 1), (uint32_t)i, N);
                pending = 1;
            }
        }
        __syncthreads();
        const bool again = pending != 0;
        __syncthreads();
        ++round;
        if (!again) break;
    }
    if (t == 0 && rounds_out) *rounds_out = round;
}

}  // namespace brn

using namespace brn;

extern "C" int brn_minibatch_indices(int64_t N, int B, uint64_t seed, uint64_t offset, int64_t* out, int* rounds, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BRN_CHECK_ARG(N > 0 && B >= 0 && (int64_t)B <= N, "brn_minibatch_indices: need 0 <= B <= N (got N=%lld B=%d)", (long long)N, B);
    BRN_CHECK_ARG(B == 0 || out, "brn_minibatch_indices: NULL output");
    // rejection sampling is efficient while the batch is a small part of the data (expected redraws B^2 / 2N per round); a
    // batch of more than half the rows is a permutation problem, not a minibatch
    BRN_CHECK_ARG((int64_t)B * 2 <= N || B <= 1, "brn_minibatch_indices: B=%d exceeds N/2=%lld (use a permutation instead)", B,
                  (long long)(N / 2));
    BRN_CHECK_ARG(B <= 8192, "brn_minibatch_indices: B=%d exceeds the supported maximum 8192", B);
    BRN_CHECK_ARG(N < ((int64_t)1 << 43), "brn_minibatch_indices: N too large");
    if (B == 0) return 0;
    int table_size = 1;
    while (table_size < 2 * B) table_size <<= 1;      // load factor <= 1/2; 8192 slots -> 16384 entries = 128 KB of shared memory
    const size_t smem = sizeof(unsigned long long) * (size_t)table_size;
    BRN_CUDA_OK(cudaFuncSetAttribute(minibatch_indices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    minibatch_indices_kernel<<<1, MB_THREADS, smem, stream>>>(N, B, seed, offset, table_size, out, rounds);
    BRN_LAUNCH_OK("minibatch_indices_kernel");
    return 0;
}
