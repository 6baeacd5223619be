// tcgen05 "NT" GEMM with 3xTF32 split for fp32-equivalent accuracy:
//     D[m, n] = sum_k A[m, k] * B[n, k]            A: [M][K], B: [N][K], both K-contiguous ("K-major")
// Operands arrive pre-split in HBM as TF32-exact pairs (hi, lo) -- see umma::split_tf32 -- and are moved
// by TMA (128-byte swizzle) into a multi-stage shared-memory ring; one elected thread issues
//     D += A_lo.B_hi^T ; D += A_hi.B_lo^T ; D += A_hi.B_hi^T      (tcgen05.mma kind::tf32, M=128, N=BN, K=8)
// accumulating in TMEM; epilogue warps read the accumulator with tcgen05.ld and hand it to an epilogue
// policy (store / fused gradient reduction).  Persistent: each CTA walks a list of (m-tile, n-tile)
// units; TMA producer, MMA issuer and epilogue are separate warps synchronised only by mbarriers.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace brn {

constexpr int UG_BM = 128;        // rows of A per tile (TMEM lanes)
constexpr int UG_BK = 32;         // fp32 elements per K chunk = 128 B = one swizzle row
constexpr int UG_STAGES = 2;
constexpr int UG_CORR_COL = 256;  // TMEM column offset of the correction accumulator

template <int BN>
struct UmmaSmem {
    static constexpr int A_BYTES = UG_BM * UG_BK * 4;       // 16 KB
    static constexpr int B_BYTES = BN * UG_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int TOTAL = UG_STAGES * STAGE_BYTES + 1024;   // + alignment slack
};

// which (m-tile, n-tile) units a CTA processes, identically enumerated by every warp role
struct UnitIter {
    int m_tiles, n_tiles, mode;
    int u, step, end, mt_fixed;
    __device__ UnitIter(int m_tiles_, int n_tiles_, int mode_) : m_tiles(m_tiles_), n_tiles(n_tiles_), mode(mode_) {
        if (mode == 0) {            // units n-major: concurrently running CTAs share the B (n) tile
            u = blockIdx.x; step = gridDim.x; end = m_tiles * n_tiles; mt_fixed = -1;
        } else {                    // a CTA keeps its m-tile and strides over n-tiles
            mt_fixed = blockIdx.x % m_tiles;
            int g = blockIdx.x / m_tiles, G = gridDim.x / m_tiles;
            u = g; step = G; end = (g < G) ? n_tiles : 0;
        }
    }
    __device__ bool valid() const { return u < end; }
    __device__ void next() { u += step; }
    __device__ int mt() const { return mode == 0 ? u % m_tiles : mt_fixed; }
    __device__ int nt() const { return mode == 0 ? u / m_tiles : u; }
};

template <int BN, class Epi>
__global__ void __launch_bounds__(64 + 32 * Epi::kEpiWarps, 1)
umma_nt_3xtf32_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                      const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                      int m_tiles, int n_tiles, int k_chunks, int mode, typename Epi::Params ep) {
    using SM = UmmaSmem<BN>;
    // two accumulators: [0, BN) main (A_hi.B_hi) and [UG_CORR_COL, +BN) correction (A_lo.B_hi + A_hi.B_lo).
    // The tensor core truncates (round-toward-zero) the fp32 accumulator after every MMA -- measured on
    // B200: relative bias -2.7e-8 per instruction -- so the tiny correction terms are kept out of the main
    // chain (3x fewer truncations of the large partial sums) and added once, in RN, by the epilogue.
    static_assert(BN <= UG_CORR_COL, "BN too large for the dual-accumulator layout");
    constexpr uint32_t TMEM_COLS = 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[UG_STAGES], empty_bar[UG_STAGES], accum_full, accum_empty;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        umma::tma_prefetch_desc(&tmAh); umma::tma_prefetch_desc(&tmAl);
        umma::tma_prefetch_desc(&tmBh); umma::tma_prefetch_desc(&tmBl);
        for (int s = 0; s < UG_STAGES; ++s) { umma::mbar_init(&full_bar[s], 1); umma::mbar_init(&empty_bar[s], 1); }
        umma::mbar_init(&accum_full, 1);
        umma::mbar_init(&accum_empty, Epi::kEpiWarps);
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc(&tmem_base_slot, TMEM_COLS);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (UnitIter it(m_tiles, n_tiles, mode); it.valid(); it.next()) {
                const int m0 = it.mt() * UG_BM, n0 = it.nt() * BN;
                for (int kc = 0; kc < k_chunks; ++kc) {
                    umma::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * SM::STAGE_BYTES;
                    umma::mbar_arrive_expect_tx(&full_bar[stage], SM::STAGE_BYTES);
                    const int k0 = kc * UG_BK;
                    umma::tma_load_2d(st, &tmAh, &full_bar[stage], k0, m0);
                    umma::tma_load_2d(st + SM::A_BYTES, &tmAl, &full_bar[stage], k0, m0);
                    umma::tma_load_2d(st + 2 * SM::A_BYTES, &tmBh, &full_bar[stage], k0, n0);
                    umma::tma_load_2d(st + 2 * SM::A_BYTES + SM::B_BYTES, &tmBl, &full_bar[stage], k0, n0);
                    if (++stage == UG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma::idesc_tf32(UG_BM, BN);
            int stage = 0; uint32_t phase = 0, acc_phase = 0;
            for (UnitIter it(m_tiles, n_tiles, mode); it.valid(); it.next()) {
                umma::mbar_wait(&accum_empty, acc_phase ^ 1);      // epilogue has drained the accumulator
                umma::tc_fence_after();
                for (int kc = 0; kc < k_chunks; ++kc) {
                    umma::mbar_wait(&full_bar[stage], phase);
                    umma::tc_fence_after();
                    const uint32_t st = umma::smem_u32(smem + stage * SM::STAGE_BYTES);
                    const uint32_t ah = st, al = st + SM::A_BYTES, bh = st + 2 * SM::A_BYTES, bl = bh + SM::B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < UG_BK / 8; ++ks) {
                        const uint32_t ko = ks * 32;               // 8 tf32 = 32 bytes along K inside the swizzle row
                        const uint64_t dah = umma::smem_desc_k_sw128(ah + ko), dal = umma::smem_desc_k_sw128(al + ko);
                        const uint64_t dbh = umma::smem_desc_k_sw128(bh + ko), dbl = umma::smem_desc_k_sw128(bl + ko);
                        umma::mma_tf32_ss(tmem_base + UG_CORR_COL, dal, dbh, idesc, (kc | ks) != 0);
                        umma::mma_tf32_ss(tmem_base + UG_CORR_COL, dah, dbl, idesc, true);
                        umma::mma_tf32_ss(tmem_base, dah, dbh, idesc, (kc | ks) != 0);
                    }
                    umma::mma_commit(&empty_bar[stage]);           // smem stage reusable once these MMAs retire
                    if (++stage == UG_STAGES) { stage = 0; phase ^= 1; }
                }
                umma::mma_commit(&accum_full);
                acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int ew = warp - 2;
        Epi epi;
        epi.begin(ep, ew, lane);
        uint32_t acc_phase = 0;
        for (UnitIter it(m_tiles, n_tiles, mode); it.valid(); it.next()) {
            umma::mbar_wait(&accum_full, acc_phase);
            umma::tc_fence_after();
            epi.tile(ep, it.mt(), it.nt(), tmem_base, ew, lane);   // must end with tmem_ld_wait()
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&accum_empty);
            acc_phase ^= 1;
        }
        epi.end(ep, ew, lane);
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// epilogue policies
// ------------------------------------------------------------------------------------------------
// Store the tile to a row-major matrix organised in column blocks (one block per sample for K3's layer 1):
//   column c of n-tile nt  ->  block j = nt*blks_per_tile + c / blk_cols,  h = c % blk_cols
//   out[j * blk_stride + row * ldo + h]   for row < M, j < total_blks, h < blk_valid
struct EpiStoreBlocks {
    static constexpr int kEpiWarps = 4;
    struct Params {
        float* out; int M; int64_t ldo; int blk_cols, blk_valid, blks_per_tile, total_blks; int64_t blk_stride;
    };
    __device__ void begin(const Params&, int, int) {}
    __device__ void end(const Params&, int, int) {}
    __device__ void tile(const Params& p, int mt, int nt, uint32_t tmem_base, int ew, int lane) {
        const int q = (ew + 2) & 3;                     // TMEM lane quarter this warp may access (warp id % 4)
        const int row = mt * UG_BM + q * 32 + lane;
        const int ncols = p.blk_cols * p.blks_per_tile;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            float v[32], w[32];
            umma::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
            umma::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + UG_CORR_COL + c0, w);
            umma::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += w[i];
            if (row < p.M) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    int c = c0 + i;
                    int j = c / p.blk_cols, h = c - j * p.blk_cols;
                    int blk = nt * p.blks_per_tile + j;
                    if (c < ncols && h < p.blk_valid && blk < p.total_blks)
                        p.out[(int64_t)blk * p.blk_stride + (int64_t)row * p.ldo + h] = v[i];
                }
            }
        }
        umma::tmem_ld_wait();
    }
};

int launch_split_tf32(const float* src, int64_t lds, int rows, int cols, float* hi, float* lo, int64_t ldd, float* thi,
                      float* tlo, int64_t ldt, cudaStream_t stream);

// A (hi/lo) [M][K] pitch lda, B (hi/lo) [N][K] pitch ldb; mode 0: units n-major round-robin over CTAs,
// mode 1: every CTA keeps one m-tile and strides over n-tiles (epilogues that accumulate across units).
template <int BN, class Epi>
inline int launch_umma_nt(const float* Ah, const float* Al, int M, int64_t lda, const float* Bh, const float* Bl, int N, int64_t ldb,
                   int K, int mode, int grid_hint, const typename Epi::Params& ep, cudaStream_t stream) {
    CUtensorMap tAh, tAl, tBh, tBl;
    if (int e = make_tmap_2d_f32(&tAh, Ah, M, K, lda, UG_BM)) return e;
    if (int e = make_tmap_2d_f32(&tAl, Al, M, K, lda, UG_BM)) return e;
    if (int e = make_tmap_2d_f32(&tBh, Bh, N, K, ldb, BN)) return e;
    if (int e = make_tmap_2d_f32(&tBl, Bl, N, K, ldb, BN)) return e;
    const int m_tiles = (M + UG_BM - 1) / UG_BM, n_tiles = (N + BN - 1) / BN, k_chunks = (K + UG_BK - 1) / UG_BK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid;
    if (mode == 0) {
        grid = m_tiles * n_tiles < sms ? m_tiles * n_tiles : sms;
    } else {
        int G = sms / m_tiles;
        if (G < 1) G = 1;
        if (G > n_tiles) G = n_tiles;
        grid = G * m_tiles;
    }
    if (grid_hint > 0 && grid_hint < grid && mode == 0) grid = grid_hint;
    auto kern = umma_nt_3xtf32_kernel<BN, Epi>;
    const int smem = UmmaSmem<BN>::TOTAL;
    BRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 64 + 32 * Epi::kEpiWarps, smem, stream>>>(tAh, tAl, tBh, tBl, m_tiles, n_tiles, k_chunks, mode, ep);
    BRN_LAUNCH_OK("umma_nt_3xtf32_kernel");
    return 0;
}


}  // namespace brn
