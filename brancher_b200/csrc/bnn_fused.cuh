// K3, fused forward: layer-1 GEMM on tcgen05 with the whole "mid" stage in its epilogue.
//
//   pre_s[b, h] = sum_p X[b, p] W1_s[h, p]          tcgen05 3xTF32, accumulators in TMEM (M = 128 batch rows,
//                                                   N = 2 samples x 104 hidden units, K = P)
//   epilogue (never leaves the SM):  h = tanh(pre + b1_s), a = W2_s h + b2_s, log-softmax, ll, da,
//                                    dW2_s / db2_s / db1_s, dpre = (W2_s^T da)(1 - h^2)
//   out: dpre^T (TF32-split, K-major B operand of the weight-gradient GEMM), the small per-sample gradients, the loss.
//
// The pre-activations (105 MB at the C3 shape) used to cross HBM twice between the GEMM and a separate mid kernel
// (137 us, latency-bound).  Here the accumulator is read from TMEM with the 16x256b shape, which delivers each warp's
// 32 rows as two m16 tiles in exactly the mma.sync accumulator-fragment layout (thread (g, t): rows g / g+8, columns
// 2t / 2t+1 of every 8-column group).  The three small layer-2 contractions then run on the warp-level tensor cores
// (mma.sync m16n8k8 tf32, 3xTF32 split) chained fragment-to-fragment:
//   A  a[b, c]   = sum_h h[b, h] W2[c, h]     the C fragment of pre IS the A fragment after permuting the contraction
//                                             index (k columns (t, t+4) <-> hidden (2t, 2t+1)); B fragments from smem
//   B  dh[b, h]  = sum_c da[b, c] W2[c, h]    same trick on the class index
//   C  dW2[c, h] = sum_b da[b, c] h[b, h]     contraction over rows: both operands transposed inside the warp with
//                                             shuffles (2 per element), K = the warp's 32 rows, RED into dW2_s
// Everything is per warp except the per-sample weights (W2_s, b1_s, b2_s), which the four warps of a sample stage in
// shared memory once per unit (two 128-thread named barriers per unit).
#pragma once
#include "umma_gemm.cuh"

namespace brn {

// ---- TMEM -> registers, 16 lanes x (8 columns x NUM): regs [4n .. 4n+3] = C fragment of column group n
#define BRN_R4(a, i) "=r"(a[i]), "=r"(a[i + 1]), "=r"(a[i + 2]), "=r"(a[i + 3])
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : BRN_R4(r, 0), BRN_R4(r, 4), BRN_R4(r, 8), BRN_R4(r, 12), BRN_R4(r, 16), BRN_R4(r, 20), BRN_R4(r, 24), BRN_R4(r, 28)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : BRN_R4(r, 0), BRN_R4(r, 4), BRN_R4(r, 8), BRN_R4(r, 12)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : BRN_R4(r, 0) : "r"(taddr) : "memory");
}
#undef BRN_R4

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int FM_W2_PITCH = 104;      // == 8 (mod 32): the 64-bit B-fragment loads of phase A are conflict-free
constexpr int FM_W2_ROWS = 16;

template <int HP, int BK>
struct FwdMidSmem {
    using Ring = UmmaSmem<2 * HP, BK>;
    static constexpr int W2_FLOATS = FM_W2_ROWS * FM_W2_PITCH;
    static constexpr int SAMPLE_FLOATS = W2_FLOATS + HP + 16;          // W2 [16][104], b1 [HP], b2 [16]
    static constexpr int EXTRA_BYTES = 2 * SAMPLE_FLOATS * 4;
    static constexpr int TOTAL = Ring::STAGES * Ring::STAGE_BYTES + 1024 + EXTRA_BYTES;
    static_assert(HP <= FM_W2_PITCH, "hidden width exceeds the staged W2 pitch");
    static_assert(TOTAL <= 227 * 1024, "shared memory budget exceeded");
};

struct FwdMidParams {
    const float* W;        // sampled small variables [S][ldw] (b1, W2, b2 at L.ob1 / L.oW2 / L.ob2)
    float* dW;             // per-sample gradient slots [S][ldw] (b1 / W2 / b2 slots pre-zeroed, accumulated with RED)
    const int32_t* y;
    BnnLayout L;
    int S;
    float inv_S;
    double* loss;
    float* dpT_hi; float* dpT_lo; int64_t ldB;      // [(S + pad) * HP][ldB]
};

// 12 warps = 3 warpgroups: WG0 = {TMA producer, MMA issuer, 2 idle warps} gives registers back (setmaxnreg.dec 40), the 8
// epilogue warps of WG1 / WG2 take them (setmaxnreg.inc 232): each SM sub-partition hosts one warp of every warpgroup,
// 32 x (40 + 2 x 232) = 16128 <= 16384 registers.  (With 10 warps a sub-partition hosts three and the cap is 168.)
constexpr int FM_THREADS = 384, FM_EPI_WARP0 = 4;
__device__ __forceinline__ void setmaxnreg_dec40() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;"); }
__device__ __forceinline__ void setmaxnreg_inc232() { asm volatile("setmaxnreg.inc.sync.aligned.u32 232;"); }

template <int HP, int BK>
__global__ void __launch_bounds__(FM_THREADS, 1)
bnn_fwd_mid_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   int m_tiles, int n_tiles, int k_chunks, int drain_chunks, FwdMidParams p) {
    constexpr int BN = 2 * HP, EW = 8, KS = HP / 8;
    using SM = UmmaSmem<BN, BK>;
    using FS = FwdMidSmem<HP, BK>;
    constexpr int UG_STAGES = SM::STAGES, SW = BK * 4;
    static_assert(HP % 8 == 0 && KS == 13, "the TMEM drain below is written for HP = 104 (x8 + x4 + x1 column groups)");
    constexpr uint32_t TMEM_COLS = 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* extra = reinterpret_cast<float*>(smem + UG_STAGES * SM::STAGE_BYTES);
    __shared__ __align__(8) uint64_t full_bar[UG_STAGES], empty_bar[UG_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        umma::tma_prefetch_desc(&tmAh); umma::tma_prefetch_desc(&tmAl);
        umma::tma_prefetch_desc(&tmBh); umma::tma_prefetch_desc(&tmBl);
        for (int s = 0; s < UG_STAGES; ++s) { umma::mbar_init(&full_bar[s], 1); umma::mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { umma::mbar_init(&acc_full[b], 1); umma::mbar_init(&acc_empty[b], EW); }
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc(&tmem_base_slot, TMEM_COLS);
    // zero the staged-weight region once: pad classes / pad hidden units stay zero for the whole kernel
    for (int i = threadIdx.x; i < 2 * FS::SAMPLE_FLOATS; i += blockDim.x) extra[i] = 0.f;
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp < FM_EPI_WARP0) {
    setmaxnreg_dec40();
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (UnitIter it(m_tiles, n_tiles, k_chunks, 0, 0, m_tiles * n_tiles); it.valid(); it.next()) {
                const int m0 = it.mt() * UG_BM, n0 = it.nt() * BN;
                for (int kc = 0; kc < k_chunks; ++kc) {
                    umma::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * SM::STAGE_BYTES;
                    umma::mbar_arrive_expect_tx(&full_bar[stage], SM::STAGE_BYTES);
                    const int k0 = kc * BK;
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
            int stage = 0; uint32_t phase = 0, blk = 0;
            for (UnitIter it(m_tiles, n_tiles, k_chunks, 0, 0, m_tiles * n_tiles); it.valid(); it.next()) {
                for (int kc0 = 0; kc0 < k_chunks; kc0 += drain_chunks, ++blk) {
                    const uint32_t buf = blk & 1, use = (blk >> 1) & 1;
                    umma::mbar_wait(&acc_empty[buf], use ^ 1);
                    umma::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * UG_BUF_COLS;
                    const int kc1 = min(kc0 + drain_chunks, k_chunks);
                    for (int kc = kc0; kc < kc1; ++kc) {
                        umma::mbar_wait(&full_bar[stage], phase);
                        umma::tc_fence_after();
                        const uint32_t st = umma::smem_u32(smem + stage * SM::STAGE_BYTES);
                        const uint32_t ah = st, al = st + SM::A_BYTES, bh = st + 2 * SM::A_BYTES, bl = bh + SM::B_BYTES;
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks) {
                            const uint32_t ko = ks * 32;
                            const uint64_t dah = umma::smem_desc_k<SW>(ah + ko), dal = umma::smem_desc_k<SW>(al + ko);
                            const uint64_t dbh = umma::smem_desc_k<SW>(bh + ko), dbl = umma::smem_desc_k<SW>(bl + ko);
                            umma::mma_tf32_ss(d_tmem, dal, dbh, idesc, kc != kc0 || ks != 0);
                            umma::mma_tf32_ss(d_tmem, dah, dbl, idesc, true);
                            umma::mma_tf32_ss(d_tmem, dah, dbh, idesc, true);
                        }
                        umma::mma_commit(&empty_bar[stage]);
                        if (++stage == UG_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma::mma_commit(&acc_full[buf]);
                }
            }
        }
    }
    } else {
        // ===================== epilogue warps: drain + mid =====================
        setmaxnreg_inc232();
        const int ew = warp - FM_EPI_WARP0;
        const int q = warp & 3;                   // TMEM lane quarter this warp may access (warp id % 4)
        const int hf = ew >> 2;                   // sample of the pair (column half of the tile)
        const int g = lane >> 2, t = lane & 3;
        const int gt = (ew & 3) * 32 + lane;      // thread index inside the sample's 4-warp group
        const int H = p.L.H, C = p.L.C, B = p.L.B;
        float* W2s = extra + hf * FS::SAMPLE_FLOATS;
        float* b1s = W2s + FS::W2_FLOATS;
        float* b2s = b1s + HP;
        const bool two_q = C > 8;                 // classes 8..15 present
        uint32_t blk = 0;
        for (UnitIter it(m_tiles, n_tiles, k_chunks, 0, 0, m_tiles * n_tiles); it.valid(); it.next()) {
            const int s = it.nt() * 2 + hf;
            const bool s_ok = s < p.S;
            const int b0 = it.mt() * UG_BM;
            // ---- stage this sample's layer-2 weights (overlaps the MMAs of the unit)
            named_bar_sync(1 + hf, 128);                                 // previous unit's readers are done
            if (s_ok) {
                const float* Ws = p.W + (int64_t)s * p.L.ldw;
                for (int c = 0; c < C; ++c)
                    for (int h = gt; h < H; h += 128) W2s[c * FM_W2_PITCH + h] = Ws[p.L.oW2 + (int64_t)c * H + h];
                for (int h = gt; h < H; h += 128) b1s[h] = Ws[p.L.ob1 + h];
                if (gt < C) b2s[gt] = Ws[p.L.ob2 + gt];
            }
            named_bar_sync(1 + hf, 128);

            // ---- drain the accumulator blocks into registers (round-to-nearest adds, see umma_gemm.cuh)
            float acc[2][KS][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int k = 0; k < KS; ++k)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[m][k][i] = 0.f;
            for (int kc0 = 0; kc0 < k_chunks; kc0 += drain_chunks, ++blk) {
                const uint32_t buf = blk & 1, use = (blk >> 1) & 1;
                umma::mbar_wait(&acc_full[buf], use);
                umma::tc_fence_after();
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32 + m * 16) << 16) + buf * UG_BUF_COLS + hf * HP;
                    float v[4 * KS];
                    tmem_ld_16x256b_x8(t0, v);
                    tmem_ld_16x256b_x4(t0 + 64, v + 32);
                    tmem_ld_16x256b_x1(t0 + 96, v + 48);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < KS; ++k)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[m][k][i] += v[4 * k + i];
                }
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&acc_empty[buf]);
            }
            if (!s_ok) continue;                 // odd sample count: the pair's second half is padding (warp-uniform)

            // rows of this thread: r[m][0] = 32q + 16m + g, r[m][1] = +8   (inside the 128-row tile)
            const int rbase = q * 32 + g;
            bool ok[2][2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                ok[m][0] = b0 + rbase + 16 * m < B;
                ok[m][1] = b0 + rbase + 16 * m + 8 < B;
            }

            // ---- phase A: h = tanh(pre + b1) (kept in acc), a = h W2^T
            float ahh[2][2][4], acr[2][2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int qq = 0; qq < 2; ++qq)
#pragma unroll
                    for (int i = 0; i < 4; ++i) ahh[m][qq][i] = acr[m][qq][i] = 0.f;
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const int h0 = 8 * k + 2 * t;
                const float2 bb = *reinterpret_cast<const float2*>(b1s + h0);
                const bool v0 = h0 < H, v1 = h0 + 1 < H;       // pad hidden units: the accumulator columns are garbage
                float bh[2][2], bl[2][2];
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    const float2 w = *reinterpret_cast<const float2*>(W2s + (8 * qq + g) * FM_W2_PITCH + h0);
                    split_tf32_trunc_lo(w.x, bh[qq][0], bl[qq][0]);
                    split_tf32_trunc_lo(w.y, bh[qq][1], bl[qq][1]);
                }
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    float* a = acc[m][k];
                    a[0] = v0 ? tanh_fast(a[0] + bb.x) : 0.f;      // (row g,   h0)
                    a[1] = v1 ? tanh_fast(a[1] + bb.y) : 0.f;      // (row g,   h0 + 1)
                    a[2] = v0 ? tanh_fast(a[2] + bb.x) : 0.f;      // (row g+8, h0)
                    a[3] = v1 ? tanh_fast(a[3] + bb.y) : 0.f;      // (row g+8, h0 + 1)
                    // A fragment (k columns t, t+4 <-> hidden h0, h0+1): (g, h0), (g+8, h0), (g, h0+1), (g+8, h0+1)
                    float hi[4], lo[4];
                    split_tf32_trunc_lo(a[0], hi[0], lo[0]);
                    split_tf32_trunc_lo(a[2], hi[1], lo[1]);
                    split_tf32_trunc_lo(a[1], hi[2], lo[2]);
                    split_tf32_trunc_lo(a[3], hi[3], lo[3]);
                    mma_tf32(ahh[m][0], hi, bh[0][0], bh[0][1]);
                    mma_tf32(acr[m][0], lo, bh[0][0], bh[0][1]);
                    mma_tf32(acr[m][0], hi, bl[0][0], bl[0][1]);
                    if (two_q) {
                        mma_tf32(ahh[m][1], hi, bh[1][0], bh[1][1]);
                        mma_tf32(acr[m][1], lo, bh[1][0], bh[1][1]);
                        mma_tf32(acr[m][1], hi, bl[1][0], bl[1][1]);
                    }
                }
            }

            // ---- log-softmax: rows (g, g+8) of each m-tile, classes {2t, 2t+1, 8+2t, 9+2t}
            float dahi[2][2][4], dalo[2][2][4];          // phase-B A fragments: (r0, 2t), (r1, 2t), (r0, 2t+1), (r1, 2t+1)
            float ll = 0.f;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int lab0 = ok[m][0] ? p.y[b0 + rbase + 16 * m] : -1, lab1 = ok[m][1] ? p.y[b0 + rbase + 16 * m + 8] : -1;
                float a[2][4];
                float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
                for (int qq = 0; qq < 2; ++qq)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = 8 * qq + 2 * t + (i & 1);
                        a[qq][i] = c < C ? ahh[m][qq][i] + acr[m][qq][i] + b2s[c] : -INFINITY;
                        if (i < 2) m0 = fmaxf(m0, a[qq][i]);
                        else m1 = fmaxf(m1, a[qq][i]);
                    }
                m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
                m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
                float ex[2][4];
                float se0 = 0.f, se1 = 0.f;
#pragma unroll
                for (int qq = 0; qq < 2; ++qq)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ex[qq][i] = expf(a[qq][i] - (i < 2 ? m0 : m1));          // exp(-inf) = 0 for the pad classes
                        if (i < 2) se0 += ex[qq][i];
                        else se1 += ex[qq][i];
                    }
                se0 += __shfl_xor_sync(0xffffffffu, se0, 1); se0 += __shfl_xor_sync(0xffffffffu, se0, 2);
                se1 += __shfl_xor_sync(0xffffffffu, se1, 1); se1 += __shfl_xor_sync(0xffffffffu, se1, 2);
                const float lse0 = m0 + logf(se0), lse1 = m1 + logf(se1);
                const float inv0 = 1.f / se0, inv1 = 1.f / se1;
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    float da[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = 8 * qq + 2 * t + (i & 1);
                        const bool okc = (i < 2 ? ok[m][0] : ok[m][1]) && c < C;
                        const int lab = i < 2 ? lab0 : lab1;
                        const float lse = i < 2 ? lse0 : lse1;
                        const float sm_ = ex[qq][i] * (i < 2 ? inv0 : inv1);
                        da[i] = okc ? (c == lab ? 1.f : 0.f) - sm_ : 0.f;            // d ll / d a_c
                        if (okc && c == lab) ll += a[qq][i] - lse;
                    }
                    split_tf32_trunc_lo(da[0], dahi[m][qq][0], dalo[m][qq][0]);
                    split_tf32_trunc_lo(da[2], dahi[m][qq][1], dalo[m][qq][1]);
                    split_tf32_trunc_lo(da[1], dahi[m][qq][2], dalo[m][qq][2]);
                    split_tf32_trunc_lo(da[3], dahi[m][qq][3], dalo[m][qq][3]);
                }
            }

            float* dWs = p.dW + (int64_t)s * p.L.ldw;
            // ---- db2[c] = sum_rows da[row, c]: this thread holds classes 8qq + 2t (+1) of 4 rows
            {
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    // fragment order: [0] = (r0, 2t), [1] = (r1, 2t), [2] = (r0, 2t+1), [3] = (r1, 2t+1); hi + lo == da exactly
                    float c0 = 0.f, c1 = 0.f;
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        c0 += (dahi[m][qq][0] + dalo[m][qq][0]) + (dahi[m][qq][1] + dalo[m][qq][1]);
                        c1 += (dahi[m][qq][2] + dalo[m][qq][2]) + (dahi[m][qq][3] + dalo[m][qq][3]);
                    }
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
                        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                    }
                    const int c = 8 * qq + 2 * t;
                    if (g == 0 && c < C) atomicAdd(&dWs[p.L.ob2 + c], c0);
                    if (g == 0 && c + 1 < C) atomicAdd(&dWs[p.L.ob2 + c + 1], c1);
                }
            }

            // ---- phase B: dh = da W2, dpre = dh (1 - h^2) -> TF32 split, transposed store; db1 column sums
            {
                float* ohi = p.dpT_hi + ((int64_t)s * HP + 2 * t) * p.ldB + b0 + rbase;
                const int64_t lo_off = p.dpT_lo - p.dpT_hi;
                const int64_t ldB = p.ldB;
#pragma unroll
                for (int j = 0; j < KS; ++j) {
                    float bh[2][2], bl[2][2];
#pragma unroll
                    for (int qq = 0; qq < 2; ++qq) {
                        const float w0 = W2s[(8 * qq + 2 * t) * FM_W2_PITCH + 8 * j + g];
                        const float w1 = W2s[(8 * qq + 2 * t + 1) * FM_W2_PITCH + 8 * j + g];
                        split_tf32_trunc_lo(w0, bh[qq][0], bl[qq][0]);
                        split_tf32_trunc_lo(w1, bh[qq][1], bl[qq][1]);
                    }
                    float cs0 = 0.f, cs1 = 0.f;
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        float dh[4] = {0.f, 0.f, 0.f, 0.f};
                        mma_tf32(dh, dalo[m][0], bh[0][0], bh[0][1]);
                        mma_tf32(dh, dahi[m][0], bl[0][0], bl[0][1]);
                        if (two_q) {
                            mma_tf32(dh, dalo[m][1], bh[1][0], bh[1][1]);
                            mma_tf32(dh, dahi[m][1], bl[1][0], bl[1][1]);
                            mma_tf32(dh, dahi[m][1], bh[1][0], bh[1][1]);
                        }
                        mma_tf32(dh, dahi[m][0], bh[0][0], bh[0][1]);
                        // C fragment: (r0, h0), (r0, h0+1), (r1, h0), (r1, h0+1), h0 = 8j + 2t -- same order as acc[m][j]
                        const float* hv = acc[m][j];
                        float dp[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) dp[i] = dh[i] * __fmaf_rn(-hv[i], hv[i], 1.f);
                        float hi[4], lo[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) umma::split_tf32(dp[i], hi[i], lo[i]);
                        // pad hidden units: zero W2 columns -> dp == 0, and the pad rows of dpT exist: no guard needed
                        float* o0 = ohi + (int64_t)(8 * j) * ldB + 16 * m;
                        if (ok[m][0]) {
                            o0[0] = hi[0]; o0[lo_off] = lo[0];
                            o0[ldB] = hi[1]; o0[ldB + lo_off] = lo[1];
                        }
                        if (ok[m][1]) {
                            o0[8] = hi[2]; o0[8 + lo_off] = lo[2];
                            o0[ldB + 8] = hi[3]; o0[ldB + 8 + lo_off] = lo[3];
                        }
                        cs0 += dp[0] + dp[2];
                        cs1 += dp[1] + dp[3];
                    }
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        cs0 += __shfl_xor_sync(0xffffffffu, cs0, o);
                        cs1 += __shfl_xor_sync(0xffffffffu, cs1, o);
                    }
                    const int h0 = 8 * j + 2 * t;
                    if (g == 0 && h0 < H) atomicAdd(&dWs[p.L.ob1 + h0], cs0);
                    if (g == 0 && h0 + 1 < H) atomicAdd(&dWs[p.L.ob1 + h0 + 1], cs1);
                }
            }

            // ---- phase C: dW2[c, h] += sum over this warp's 32 rows of da[row, c] h[row, h]
            // k-step (m, hh) = rows {16m + 8hh + 0..7}.  A[class][k = row]: a0 = da[row t][g], a1 = da[row t][g + 8],
            // a2 = da[row t+4][g], a3 = da[row t+4][g + 8]; B[k = row][n = hidden]: b0 = h[row t][8n + g], b1 = h[row t+4][8n + g].
            // Source of (row r', column 2t' + par) is lane 4 r' + t' -- two shuffles (par = 0 / 1) and a select per element.
            {
                const int srcA = 4 * t + (g >> 1), srcB = 4 * (t + 4) + (g >> 1);
                const bool odd = g & 1;
                float Ahi[4][4], Alo[4][4];
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int ks = 2 * m + hh;
#pragma unroll
                        for (int qq = 0; qq < 2; ++qq) {
                            // da of row half hh: even class -> fragment slot hh, odd class -> slot 2 + hh
                            const float de = dahi[m][qq][hh] + dalo[m][qq][hh], dod = dahi[m][qq][2 + hh] + dalo[m][qq][2 + hh];
                            const float ea = __shfl_sync(0xffffffffu, de, srcA), oa = __shfl_sync(0xffffffffu, dod, srcA);
                            const float eb = __shfl_sync(0xffffffffu, de, srcB), ob = __shfl_sync(0xffffffffu, dod, srcB);
                            split_tf32_trunc_lo(odd ? oa : ea, Ahi[ks][qq], Alo[ks][qq]);             // a0 / a1
                            split_tf32_trunc_lo(odd ? ob : eb, Ahi[ks][2 + qq], Alo[ks][2 + qq]);     // a2 / a3
                        }
                    }
                float* dW2 = dWs + p.L.oW2;
                const bool vec2 = ((H & 1) == 0) && ((p.L.oW2 & 1) == 0) && ((p.L.ldw & 1) == 0);
#pragma unroll
                for (int n = 0; n < KS; ++n) {
                    float cacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int ks = 2 * m + hh;
                            const float he = acc[m][n][2 * hh], ho = acc[m][n][2 * hh + 1];
                            const float ea = __shfl_sync(0xffffffffu, he, srcA), oa = __shfl_sync(0xffffffffu, ho, srcA);
                            const float eb = __shfl_sync(0xffffffffu, he, srcB), ob = __shfl_sync(0xffffffffu, ho, srcB);
                            float bh0, bl0, bh1, bl1;
                            split_tf32_trunc_lo(odd ? oa : ea, bh0, bl0);
                            split_tf32_trunc_lo(odd ? ob : eb, bh1, bl1);
                            mma_tf32(cacc, Alo[ks], bh0, bh1);
                            mma_tf32(cacc, Ahi[ks], bl0, bl1);
                            mma_tf32(cacc, Ahi[ks], bh0, bh1);
                        }
                    // C fragment: (class g, 8n + 2t), (g, 8n + 2t + 1), (g + 8, 8n + 2t), (g + 8, 8n + 2t + 1)
                    const int h0 = 8 * n + 2 * t;
                    if (vec2) {
                        if (h0 < H) {           // H even: h0 + 1 < H as well
                            if (g < C) red_add_v2(dW2 + (int64_t)g * H + h0, cacc[0], cacc[1]);
                            if (g + 8 < C) red_add_v2(dW2 + (int64_t)(g + 8) * H + h0, cacc[2], cacc[3]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c = g + (i >= 2 ? 8 : 0), h = h0 + (i & 1);
                            if (c < C && h < H) atomicAdd(&dW2[(int64_t)c * H + h], cacc[i]);
                        }
                    }
                }
            }
            double tot = (double)ll;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) atomicAdd(p.loss, -tot * (double)p.inv_S);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------
// backward GEMM epilogue: the per-sample weight gradient dW1_s never reaches HBM.  An epilogue thread owns one input
// feature p (tile row) and the HP hidden units of ONE sample; it folds its accumulator row into the sample-axis sums
//     gwT[p][h] += dW1_s[h, p]            gweT[p][h] += dW1_s[h, p] * eps_s[h, p]
// with 16-byte vector REDs (the [p][h] layout makes a thread's hidden units contiguous).  eps_s[h][p] is read coalesced
// across the warp (32 consecutive p).  Linear in the accumulator, so K-split partial sums need no special case.
// ---------------------------------------------------------------------------------------------------
struct EpiSampleReduce {
    struct Params {
        float* gwT; float* gweT;       // [P][HP], zeroed by the caller
        const float* eps; int64_t lde; // eps[s * lde + h * P + p]
        int P, H, HP, S;
    };
    template <int CPT>
    static __device__ __forceinline__ void finish(const Params& p, float (&r)[CPT], int row, int blk, bool, int) {
        if (row >= p.P || blk >= p.S) return;
        const float* e = p.eps + (int64_t)blk * p.lde + row;
        float* ow = p.gwT + (int64_t)row * p.HP;
        float* oe = p.gweT + (int64_t)row * p.HP;
#pragma unroll
        for (int i0 = 0; i0 < CPT; i0 += 8) {        // 8 independent loads in flight per batch
            float ev[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ev[j] = (i0 + j < p.H) ? __ldg(e + (int64_t)(i0 + j) * p.P) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
                const int i = i0 + j;
                if (i >= p.H) break;
                float a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = (i + u < p.H) ? r[i + u] : 0.f;      // pad columns hold garbage
                red_add_v4(ow + i, a[0], a[1], a[2], a[3]);
                red_add_v4(oe + i, a[0] * ev[j], a[1] * ev[j + 1], a[2] * ev[j + 2], a[3] * ev[j + 3]);
            }
        }
    }
};

}  // namespace brn
