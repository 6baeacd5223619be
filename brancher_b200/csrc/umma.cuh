// sm_100a primitives for the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma kind::tf32 / commit / ld) and the shared-memory / instruction descriptors.
// Raw inline PTX -- no CUTLASS dependency.  Descriptor bit layouts follow the PTX ISA "tcgen05" matrix
// and instruction descriptor tables (same fields as cute/arch/mma_sm100_desc.hpp documents).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace brn {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Bounded variant: no wait of these kernels legitimately lasts seconds, so a protocol error traps (the launch fails
// loudly) instead of hanging the device.  Costs one register: used where the budget allows.
__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins == (1u << 28)) __trap();
    }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}
// 2-D tile load: coordinates (c0 = innermost element index, c1 = row index); completes on `bar`.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---------------------------------------------------------------- TMEM allocation
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 B (32 tf32), 8-row atoms
// of 1024 B (SBO), 16-byte chunks XOR-swizzled by (row % 8) -- exactly what a TMA box of inner extent
// 128 B with CU_TENSOR_MAP_SWIZZLE_128B writes.  `addr` = tile base (1024-B aligned) + k byte offset.
// SW = swizzle span in bytes = bytes of one K-chunk row: 128 (32 tf32, SWIZZLE_128B, layout code 2) or 64 (16 tf32,
// SWIZZLE_64B, layout code 4); 8-row atoms of 8*SW bytes (SBO).
template <int SW>
__device__ __forceinline__ uint64_t smem_desc_k(uint32_t addr) {
    static_assert(SW == 128 || SW == 64, "unsupported swizzle span");
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);          // start address           bits [0,14)
    d |= (uint64_t)1 << 16;                           // LBO (unused, 16 B)      bits [16,30)
    d |= (uint64_t)((8u * SW) >> 4) << 32;            // SBO = 8 rows            bits [32,46)
    d |= (uint64_t)1 << 46;                           // descriptor version 1    bits [46,48)
    d |= (uint64_t)(SW == 128 ? 2 : 4) << 61;         // SWIZZLE_128B / _64B     bits [61,64)
    return d;
}
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t addr) { return smem_desc_k<128>(addr); }
// MN-major operand, SWIZZLE_128B: the image TMA leaves for a box [K rows][128 bytes along MN] -- 128-byte rows run along MN
// (32 tf32 / 64 fp16 elements), 8 consecutive K indices form one 1024-byte swizzle atom.  LBO = byte distance between
// consecutive 128-byte column blocks along MN (= one box), SBO = between 8-row groups along K (1024 when they are adjacent).
// 16-bit operands: layout 2 = SWIZZLE_128B (8-row atoms).  32-bit (tf32) operands can only be read MN-major through layout 1 =
// SWIZZLE_128B_BASE32B: 128-byte rows whose 32-byte chunks are XOR-ed with (row % 4), 4-row atoms of 512 bytes -- the image
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B produces (CUTLASS: "for mn-major tf32 operands, SW128_32B is the only available smem layout").
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

// Instruction descriptor for kind::tf32, fp32 accumulate, K-major A and B, dense.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4)                   // c_format = F32        bits [4,6)
           | (2u << 7)                 // a_format = TF32       bits [7,10)
           | (2u << 10)                // b_format = TF32       bits [10,13)
           | (0u << 15) | (0u << 16)   // a_major, b_major = K
           | ((uint32_t)(N >> 3) << 17)   // n_dim              bits [17,23)
           | ((uint32_t)(M >> 4) << 24);  // m_dim              bits [24,29)
}

// the same with both operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t idesc_tf32_mn(int M, int N) { return idesc_tf32(M, N) | (1u << 15) | (1u << 16); }

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// kind::f16 with fp16 operands (a_format = b_format = 0), fp32 accumulate: K = 16 elements (32 bytes) per instruction, i.e.
// the same descriptors / byte strides as tf32 at twice the MACs per instruction.
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// true in exactly one lane of the (converged) warp.  Issuing tcgen05 instructions under this predicate instead of
// `lane == 0` lets ptxas emit them once: with a plain lane test it wraps every UTCHMMA in an ELECT / BRA.U.ANY loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// KIND: 0 = tf32 operands (fp32 words), 1 = fp16 operands (two elements per 32-bit word)
template <int KIND>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    if (KIND == 0) mma_tf32_ss(d_tmem, a_desc, b_desc, idesc, accumulate);
    else mma_f16_ss(d_tmem, a_desc, b_desc, idesc, accumulate);
}
template <int KIND>
__host__ __device__ constexpr uint32_t idesc_kind(int M, int N) { return KIND == 0 ? idesc_tf32(M, N) : idesc_f16(M, N); }

// mbarrier arrives once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- TMEM -> registers
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (lane_base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// split an fp32 into two TF32-exact parts (round-to-nearest on the 13 dropped mantissa bits):
//   hi = rn_tf32(x),  lo = rn_tf32(x - hi)      =>  |x - hi - lo| <= 2^-24 |x|
// both are exactly representable in TF32, so the tensor core's own fp32->tf32 conversion (whatever its
// rounding) is the identity on them; the dropped lo*lo product is O(2^-24) relative as well.
__device__ __forceinline__ float rn_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x00001000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rn_tf32(x);
    lo = rn_tf32(x - hi);
}

}  // namespace umma

// Host: encode a 2-D fp32 row-major tensor [rows][cols] (row stride ld elements) for TMA tiles of
// box_rows x box_cols floats (box_cols = 32: 128-byte swizzle, 16: 64-byte swizzle) with zero OOB fill.
int make_tmap_2d_f32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols, bool atom32 = false);
// The same for a 2-D fp16 tensor: cols / ld in ELEMENTS, box_cols = 64 (128-byte swizzle) or 32 (64-byte swizzle) elements.
int make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols);

}  // namespace brn
