// Batched fp32 SIMT GEMM used by the "simt" variants of the likelihood families (and as the
// always-available cross-check of the tcgen05 variants):  C[z][m,n] = sum_k A[z](m,k) * B[z](n,k).
// 128x128x8 CTA tile, 256 threads, 8x8 register micro-tile, double-buffered shared memory,
// 128-bit global loads when the leading dimensions allow it.
#pragma once
#include "common.cuh"

namespace brn {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8, SG_PAD = 4;

// Load a [128 x 8] operand tile into smem laid out [k][m].  KMAJOR: element (m,k) at p[m*ld + k]
// (k contiguous); else at p[k*ld + m] (m contiguous).
template <bool KMAJOR>
__device__ __forceinline__ void sg_load_tile(const float* __restrict__ p, int64_t ld, int m0, int k0, int M, int K,
                                             bool vec_ok, float (&reg)[4]) {
    const int t = threadIdx.x;
    if (KMAJOR) {
        int m = m0 + (t >> 1), k = k0 + (t & 1) * 4;
        if (vec_ok && m < M && k + 3 < K) {
            float4 v = *reinterpret_cast<const float4*>(p + (int64_t)m * ld + k);
            reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) reg[j] = (m < M && k + j < K) ? p[(int64_t)m * ld + k + j] : 0.f;
        }
    } else {
        int k = k0 + (t >> 5), m = m0 + (t & 31) * 4;
        if (vec_ok && k < K && m + 3 < M) {
            float4 v = *reinterpret_cast<const float4*>(p + (int64_t)k * ld + m);
            reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) reg[j] = (k < K && m + j < M) ? p[(int64_t)k * ld + m + j] : 0.f;
        }
    }
}

template <bool KMAJOR>
__device__ __forceinline__ void sg_store_tile(float (*sm)[SG_BM + SG_PAD], const float (&reg)[4]) {
    const int t = threadIdx.x;
    if (KMAJOR) {
        int m = t >> 1, k = (t & 1) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) sm[k + j][m] = reg[j];
    } else {
        int k = t >> 5, m = (t & 31) * 4;
        *reinterpret_cast<float4*>(&sm[k][m]) = make_float4(reg[0], reg[1], reg[2], reg[3]);
    }
}

template <bool A_KMAJOR, bool B_KMAJOR, bool ACCUMULATE>
__global__ void __launch_bounds__(256)
sgemm_batched_kernel(const float* __restrict__ A, int64_t lda, int64_t strideA, const float* __restrict__ B,
                     int64_t ldb, int64_t strideB, float* __restrict__ C, int64_t ldc, int64_t strideC, int M, int N,
                     int K, int a_vec, int b_vec) {
    __shared__ __align__(16) float As[2][SG_BK][SG_BM + SG_PAD];
    __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + SG_PAD];
    const int z = blockIdx.z;
    A += (int64_t)z * strideA;
    B += (int64_t)z * strideB;
    C += (int64_t)z * strideC;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[4], rb[4];
    sg_load_tile<A_KMAJOR>(A, lda, m0, 0, M, K, a_vec, ra);
    sg_load_tile<B_KMAJOR>(B, ldb, n0, 0, N, K, b_vec, rb);
    sg_store_tile<A_KMAJOR>(As[0], ra);
    sg_store_tile<B_KMAJOR>(Bs[0], rb);
    __syncthreads();

    const int nk = (K + SG_BK - 1) / SG_BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            sg_load_tile<A_KMAJOR>(A, lda, m0, (kt + 1) * SG_BK, M, K, a_vec, ra);
            sg_load_tile<B_KMAJOR>(B, ldb, n0, (kt + 1) * SG_BK, N, K, b_vec, rb);
        }
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sg_store_tile<A_KMAJOR>(As[cur ^ 1], ra);
            sg_store_tile<B_KMAJOR>(Bs[cur ^ 1], rb);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= N) continue;
            float* c = C + (int64_t)m * ldc + n;
            if (ACCUMULATE) *c += acc[i][j]; else *c = acc[i][j];
        }
    }
}

inline bool sg_vec_ok(const float* p, int64_t ld, int64_t stride, int contiguous_dim) {
    return ((uintptr_t)p % 16 == 0) && (ld % 4 == 0) && (stride % 4 == 0) && (contiguous_dim % 4 == 0);
}

// C[z] (M x N, ldc) = A[z] (M x K) * B[z]^T (N x K);  *_kmajor says which index is contiguous.
template <bool A_KMAJOR, bool B_KMAJOR>
inline int launch_sgemm_batched(const float* A, int64_t lda, int64_t strideA, const float* B, int64_t ldb,
                                int64_t strideB, float* C, int64_t ldc, int64_t strideC, int M, int N, int K, int batch,
                                cudaStream_t stream) {
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM, batch);
    int a_vec = sg_vec_ok(A, lda, strideA, A_KMAJOR ? K : M);
    int b_vec = sg_vec_ok(B, ldb, strideB, B_KMAJOR ? K : N);
    sgemm_batched_kernel<A_KMAJOR, B_KMAJOR, false><<<grid, 256, 0, stream>>>(A, lda, strideA, B, ldb, strideB, C, ldc,
                                                                               strideC, M, N, K, a_vec, b_vec);
    BRN_LAUNCH_OK("sgemm_batched_kernel");
    return 0;
}

}  // namespace brn
