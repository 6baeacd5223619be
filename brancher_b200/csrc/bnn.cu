// K3: Bayesian neural network P-H-C (tanh, Categorical likelihood) -- ELBO forward + pathwise backward.
// See include/brancher_cuda.h (brn_bnn_elbo_fwd_bwd) for the contract and the reference lines replaced.
//
// Stage plan (shared by the "simt" and "tcgen05" variants; they differ in stages 2 and 4):
//   1. noise + weight sampling          W_s = mu + softplus(rho) * eps_s                (HBM-bound)
//   2. layer-1 GEMM per sample          pre_s[B,H] = X[B,P] . W1_s^T[P,H]               (GEMM)
//   3. "mid": tanh, layer 2, log-softmax, ll, and backward down to d pre_s              (HBM-bound)
//   4. layer-1 weight gradient          dW1_s[H,P] = dpre_s^T[H,B] . X[B,P]             (GEMM)
//   5. sample-axis reduction + K1a      gw = sum_s dW_s, gwe = sum_s dW_s*eps_s, prior/entropy, chain rule
#include "meanfield.cuh"
#include "sgemm.cuh"
#include "umma_gemm.cuh"
#include <stdlib.h>
#include <type_traits>

namespace brn {

constexpr int BNN_MAXC = 16;

struct BnnLayout {
    int B, P, H, C;
    int64_t oW1, ob1, oW2, ob2, numel, ldw;
    __host__ __device__ BnnLayout(int B_, int P_, int H_, int C_) : B(B_), P(P_), H(H_), C(C_) {
        oW1 = 0;
        ob1 = (int64_t)H * P;
        oW2 = ob1 + H;
        ob2 = oW2 + (int64_t)C * H;
        numel = ob2 + C;
        ldw = (numel + 3) / 4 * 4;
    }
};

// One CTA = R rows of the batch for one sample; one thread per row.
// pre (in):  X.W1_s^T without bias;  pre (out): d ll_s / d pre.
// If dpT_hi != NULL the result is written transposed and TF32-split instead of in place:
//   dpT_{hi,lo}[(s*Hp + h) * ldB + b]   (rows h in [H, Hp) zero) -- the K-major B operand of the tcgen05
//   weight-gradient GEMM.
__global__ void bnn_mid_kernel(float* __restrict__ pre, const float* __restrict__ W, float* __restrict__ dW,
                               const int32_t* __restrict__ y, BnnLayout L, int R, float inv_S,
                               double* __restrict__ loss, float* __restrict__ dpT_hi, float* __restrict__ dpT_lo,
                               int Hp, int64_t ldB) {
    extern __shared__ float sm[];
    const int H = L.H, C = L.C, B = L.B, HP = H + 1;
    float* tile = sm;                    // [R][H+1]
    float* W2s = tile + (size_t)R * HP;  // [C][H]
    float* b1s = W2s + C * H;            // [H]
    float* b2s = b1s + H;                // [C]
    float* das = b2s + C;                // [R][C]
    __shared__ double red[32];

    const int s = blockIdx.y, b0 = blockIdx.x * R, t = threadIdx.x, nt = blockDim.x;
    const float* Ws = W + (int64_t)s * L.ldw;
    float* dWs = dW + (int64_t)s * L.ldw;
    float* pre_s = pre + (int64_t)s * B * H;

    if (dpT_hi != nullptr) {     // tcgen05 variant: pre arrives transposed, [S][H][B]
        for (int idx = t; idx < R * H; idx += nt) {
            int h = idx / R, r = idx - h * R;
            tile[r * HP + h] = (b0 + r < B) ? pre_s[(int64_t)h * B + b0 + r] : 0.f;
        }
    } else {
        for (int idx = t; idx < R * H; idx += nt) {
            int r = idx / H, h = idx - r * H;
            tile[r * HP + h] = (b0 + r < B) ? pre_s[(int64_t)(b0 + r) * H + h] : 0.f;
        }
    }
    for (int idx = t; idx < C * H; idx += nt) W2s[idx] = Ws[L.oW2 + idx];
    for (int idx = t; idx < H; idx += nt) b1s[idx] = Ws[L.ob1 + idx];
    for (int idx = t; idx < C; idx += nt) b2s[idx] = Ws[L.ob2 + idx];
    __syncthreads();

    const int r = t;
    const bool valid = (r < R) && (b0 + r < B);
    float da[BNN_MAXC];
    float ll = 0.f;
    if (valid) {
        float a[BNN_MAXC];
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c) a[c] = 0.f;
        for (int h = 0; h < H; ++h) {
            float v = tanhf(tile[r * HP + h] + b1s[h]);
            tile[r * HP + h] = v;
#pragma unroll
            for (int c = 0; c < BNN_MAXC; ++c)
                if (c < C) a[c] = __fmaf_rn(W2s[c * H + h], v, a[c]);
        }
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c)
            if (c < C) { a[c] += b2s[c]; m = fmaxf(m, a[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c)
            if (c < C) se += expf(a[c] - m);
        const float lse = m + logf(se);
        const int label = y[b0 + r];
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c) {
            if (c < C) {
                float p = expf(a[c] - lse);
                da[c] = (c == label ? 1.f : 0.f) - p;      // d ll / d a_c
                if (c == label) ll = a[c] - lse;
            } else da[c] = 0.f;
        }
    } else {
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c) da[c] = 0.f;
    }
    if (r < R)
        for (int c = 0; c < C; ++c) das[r * C + c] = da[c];
    __syncthreads();

    // dW2_s[c,h] += sum_r da_r[c] * h_r[h] ;  db2_s[c] += sum_r da_r[c]
    for (int o = t; o < C * H; o += nt) {
        int c = o / H, h = o - c * H;
        float acc = 0.f;
        for (int rr = 0; rr < R; ++rr) acc = __fmaf_rn(das[rr * C + c], tile[rr * HP + h], acc);
        atomicAdd(&dWs[L.oW2 + o], acc);
    }
    for (int c = t; c < C; c += nt) {
        float acc = 0.f;
        for (int rr = 0; rr < R; ++rr) acc += das[rr * C + c];
        atomicAdd(&dWs[L.ob2 + c], acc);
    }
    __syncthreads();

    if (r < R) {
        for (int h = 0; h < H; ++h) {
            float hv = tile[r * HP + h];
            float dh = 0.f;
#pragma unroll
            for (int c = 0; c < BNN_MAXC; ++c)
                if (c < C) dh = __fmaf_rn(da[c], W2s[c * H + h], dh);
            tile[r * HP + h] = dh * (1.f - hv * hv);
        }
    }
    __syncthreads();

    for (int h = t; h < H; h += nt) {
        float acc = 0.f;
        for (int rr = 0; rr < R; ++rr) acc += tile[rr * HP + h];
        atomicAdd(&dWs[L.ob1 + h], acc);
    }
    if (dpT_hi != nullptr) {
        for (int idx = t; idx < Hp * R; idx += nt) {
            int h = idx / R, rr = idx - h * R;
            if (b0 + rr >= B) continue;
            float hi = 0.f, lo = 0.f;
            if (h < H) umma::split_tf32(tile[rr * HP + h], hi, lo);
            int64_t o = ((int64_t)s * Hp + h) * ldB + b0 + rr;
            dpT_hi[o] = hi;
            dpT_lo[o] = lo;
        }
    } else {
        for (int idx = t; idx < R * H; idx += nt) {
            int rr = idx / H, h = idx - rr * H;
            if (b0 + rr < B) pre_s[(int64_t)(b0 + rr) * H + h] = tile[rr * HP + h];
        }
    }
    double tot = block_sum<double>((double)ll, red);
    if (t == 0) atomicAdd(loss, -tot * (double)inv_S);
}

// W1_s = mu + softplus(rho)*eps_s, TF32-split, in the padded K-major layout [S][Hp][ldP] (rows h >= H zero)
__global__ void sample_w1_split_kernel(const float* __restrict__ mu, const float* __restrict__ rho,
                                       const float* __restrict__ eps, int64_t lde, float* __restrict__ hi,
                                       float* __restrict__ lo, int H, int P, int Hp, int64_t ldP) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y, s = blockIdx.z;
    if (p >= P) return;
    float vh = 0.f, vl = 0.f;
    if (h < H) {
        const int64_t i = (int64_t)h * P + p;
        umma::split_tf32(__fmaf_rn(softplusf(rho[i]), eps[(int64_t)s * lde + i], mu[i]), vh, vl);
    }
    const int64_t o = ((int64_t)s * Hp + h) * ldP + p;
    hi[o] = vh;
    lo[o] = vl;
}

__global__ void softplus_kernel(const float* __restrict__ rho, float* __restrict__ sigma, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sigma[i] = softplusf(rho[i]);
}

// Fused noise + weight sampling + TF32 split for layer 1 (P % 4 == 0): one thread = 4 consecutive weights of one
// sample.  eps comes from Philox (and is also written out for stage 5) or from the injected tensor.
//   Wh/Wl[(s*Hp + h)*ldP + p] = split(mu + sigma*eps) ;  rows h in [H, Hp) = 0
__global__ void __launch_bounds__(256)
sample_w1_fused_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, const float* __restrict__ eps_in,
                       int64_t lde_in, float* __restrict__ eps_out, int64_t lde_out, float* __restrict__ hi,
                       float* __restrict__ lo, int H, int P, int Hp, int64_t ldP, brn_sample_range r, uint32_t var_id) {
    const int64_t qq = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // quad index in the padded [Hp][P] matrix
    const int s = blockIdx.y;
    if (qq * 4 >= (int64_t)H * P) return;      // pad rows h in [H, Hp) are never written (their accumulator columns are discarded)
    const int h = (int)((qq * 4) / P), p = (int)((qq * 4) - (int64_t)h * P);
    float4 vh = make_float4(0.f, 0.f, 0.f, 0.f), vl = vh;
    {
        const int64_t i = (int64_t)h * P + p;
        float4 e;
        if (eps_in) {
            e = *reinterpret_cast<const float4*>(eps_in + (int64_t)s * lde_in + i);
        } else {
            Normal4 n = philox_normal4(r.seed, r.offset, var_id, (uint32_t)(r.s0 + s), (uint32_t)(i >> 2));
            e = make_float4(n.v[0], n.v[1], n.v[2], n.v[3]);
        }
        // stage 5 reads injected noise from the workspace; Philox noise is regenerated there instead (eps_out == NULL)
        if (eps_out) *reinterpret_cast<float4*>(eps_out + (int64_t)s * lde_out + i) = e;
        const float4 m = *reinterpret_cast<const float4*>(mu + i);
        const float4 sg = *reinterpret_cast<const float4*>(sigma + i);
        umma::split_tf32(__fmaf_rn(sg.x, e.x, m.x), vh.x, vl.x);
        umma::split_tf32(__fmaf_rn(sg.y, e.y, m.y), vh.y, vl.y);
        umma::split_tf32(__fmaf_rn(sg.z, e.z, m.z), vh.z, vl.z);
        umma::split_tf32(__fmaf_rn(sg.w, e.w, m.w), vh.w, vl.w);
    }
    const int64_t o = ((int64_t)s * Hp + h) * ldP + p;
    *reinterpret_cast<float4*>(hi + o) = vh;
    *reinterpret_cast<float4*>(lo + o) = vl;
}

// ---------------------------------------------------------------------------------------------------
// "mid" stage, fast version: one CTA = 128 batch rows of one sample, 256 threads = 2 threads per row (each owns
// half of the hidden units).  Everything between the two layer-1 GEMMs happens here in shared memory:
//   h = tanh(pre + b1), a = W2 h + b2, log-softmax, ll, da, dW2 / db2 (reduced over the 128 rows),
//   dh = W2^T da, dpre = dh (1 - h^2), db1.
// TC: `pre` arrives transposed [S][H][B] (tcgen05 GEMM output) and dpre leaves transposed + TF32-split
// ([S*Hp][ldB] hi/lo, rows h >= H zero), both coalesced along the batch axis; else `pre` is [S][B][H] and is
// overwritten in place.
// ---------------------------------------------------------------------------------------------------
constexpr int MID_R = 128;

struct MidSmem {
    int HP1;              // odd row pitch of the tile -> conflict-free column and row access
    size_t tile, w2, b1, b2, das, apart, total;   // offsets in floats
    __host__ __device__ MidSmem(int H, int C) {
        HP1 = H | 1;
        tile = 0;
        w2 = tile + (size_t)MID_R * HP1;
        b1 = w2 + (size_t)C * H;
        b2 = b1 + H;
        das = (b2 + C + 3) / 4 * 4;          // float4-aligned
        apart = das + MID_R * 16;
        total = apart + MID_R * 16;
    }
};

template <bool TC>
__global__ void __launch_bounds__(256)
bnn_mid2_kernel(float* __restrict__ pre, const float* __restrict__ W, float* __restrict__ dW,
                const int32_t* __restrict__ y, BnnLayout L, float inv_S, double* __restrict__ loss,
                float* __restrict__ dpT_hi, float* __restrict__ dpT_lo, int Hp, int64_t ldB) {
    extern __shared__ __align__(16) float sm[];
    const int H = L.H, C = L.C, B = L.B;
    const MidSmem M(H, C);
    const int HP1 = M.HP1;
    float* tile = sm + M.tile;
    float* W2s = sm + M.w2;
    float* b1s = sm + M.b1;
    float* b2s = sm + M.b2;
    float* das = sm + M.das;      // [128][16]
    float* apart = sm + M.apart;  // [128][16]
    __shared__ double red[32];

    const int s = blockIdx.y, b0 = blockIdx.x * MID_R, t = threadIdx.x;
    const int r = t & (MID_R - 1), half = t >> 7;
    const float* Ws = W + (int64_t)s * L.ldw;
    float* dWs = dW + (int64_t)s * L.ldw;
    float* pre_s = pre + (int64_t)s * B * H;
    const bool row_ok = b0 + r < B;

    // ---- P0: stage the tile
    if (TC) {
        for (int h = half; h < H; h += 2) tile[r * HP1 + h] = row_ok ? pre_s[(int64_t)h * B + b0 + r] : 0.f;
    } else {
        for (int idx = t; idx < MID_R * H; idx += 256) {
            int rr = idx / H, h = idx - rr * H;
            tile[rr * HP1 + h] = (b0 + rr < B) ? pre_s[(int64_t)(b0 + rr) * H + h] : 0.f;
        }
    }
    for (int idx = t; idx < C * H; idx += 256) W2s[idx] = Ws[L.oW2 + idx];
    for (int idx = t; idx < H; idx += 256) b1s[idx] = Ws[L.ob1 + idx];
    for (int idx = t; idx < C; idx += 256) b2s[idx] = Ws[L.ob2 + idx];
    __syncthreads();

    // ---- P1: hidden activations + logits (each thread: its half of the hidden units)
    const int Hh = (H + 1) >> 1, h0 = half * Hh, h1 = min(H, h0 + Hh);
    float a[BNN_MAXC];
#pragma unroll
    for (int c = 0; c < BNN_MAXC; ++c) a[c] = 0.f;
    for (int h = h0; h < h1; ++h) {
        float v = tanhf(tile[r * HP1 + h] + b1s[h]);
        tile[r * HP1 + h] = v;
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c)
            if (c < C) a[c] = __fmaf_rn(W2s[c * H + h], v, a[c]);
    }
    if (half == 1) {
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c) apart[r * 16 + c] = a[c];
    }
    __syncthreads();
    float da[BNN_MAXC];
    float ll = 0.f;
    if (half == 0) {
        if (row_ok) {
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < BNN_MAXC; ++c)
                if (c < C) { a[c] += apart[r * 16 + c] + b2s[c]; m = fmaxf(m, a[c]); }
            float se = 0.f;
#pragma unroll
            for (int c = 0; c < BNN_MAXC; ++c)
                if (c < C) se += expf(a[c] - m);
            const float lse = m + logf(se);
            const int label = y[b0 + r];
#pragma unroll
            for (int c = 0; c < BNN_MAXC; ++c) {
                if (c < C) {
                    da[c] = (c == label ? 1.f : 0.f) - expf(a[c] - lse);      // d ll / d a_c
                    if (c == label) ll = a[c] - lse;
                } else da[c] = 0.f;
            }
        } else {
#pragma unroll
            for (int c = 0; c < BNN_MAXC; ++c) da[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c) das[r * 16 + c] = da[c];
    }
    __syncthreads();

    // ---- P2: dW2_s[c,h] += sum_r da_r[c] h_r[h] ; db2_s[c] += sum_r da_r[c]
    {
        const int g = half;                           // class group: classes [8g, 8g+8)
        if (8 * g < C) {
            for (int hh = r; hh < H; hh += MID_R) {
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                for (int rr = 0; rr < MID_R; ++rr) {
                    const float hv = tile[rr * HP1 + hh];
                    const float4 d0 = *reinterpret_cast<const float4*>(&das[rr * 16 + 8 * g]);
                    const float4 d1 = *reinterpret_cast<const float4*>(&das[rr * 16 + 8 * g + 4]);
                    acc[0] = __fmaf_rn(d0.x, hv, acc[0]); acc[1] = __fmaf_rn(d0.y, hv, acc[1]);
                    acc[2] = __fmaf_rn(d0.z, hv, acc[2]); acc[3] = __fmaf_rn(d0.w, hv, acc[3]);
                    acc[4] = __fmaf_rn(d1.x, hv, acc[4]); acc[5] = __fmaf_rn(d1.y, hv, acc[5]);
                    acc[6] = __fmaf_rn(d1.z, hv, acc[6]); acc[7] = __fmaf_rn(d1.w, hv, acc[7]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (8 * g + j < C) atomicAdd(&dWs[L.oW2 + (int64_t)(8 * g + j) * H + hh], acc[j]);
            }
        }
        if (t < C) {
            float acc = 0.f;
            for (int rr = 0; rr < MID_R; ++rr) acc += das[rr * 16 + t];
            atomicAdd(&dWs[L.ob2 + t], acc);
        }
    }
    __syncthreads();

    // ---- P3: dpre = (W2^T da) (1 - h^2)
    if (half == 1) {
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c) da[c] = das[r * 16 + c];
    }
    for (int h = h0; h < h1; ++h) {
        const float hv = tile[r * HP1 + h];
        float dh = 0.f;
#pragma unroll
        for (int c = 0; c < BNN_MAXC; ++c)
            if (c < C) dh = __fmaf_rn(da[c], W2s[c * H + h], dh);
        tile[r * HP1 + h] = dh * (1.f - hv * hv);
    }
    __syncthreads();

    // ---- P4: db1_s[h] += sum_r dpre_r[h] ; write dpre
    for (int hh = r; hh < H; hh += MID_R) {
        float acc = 0.f;
        const int r0 = half * (MID_R / 2);
#pragma unroll 8
        for (int rr = r0; rr < r0 + MID_R / 2; ++rr) acc += tile[rr * HP1 + hh];
        atomicAdd(&dWs[L.ob1 + hh], acc);
    }
    if (TC) {
        if (row_ok) {
            for (int h = half; h < Hp; h += 2) {
                float hi = 0.f, lo = 0.f;
                if (h < H) umma::split_tf32(tile[r * HP1 + h], hi, lo);
                const int64_t o = ((int64_t)s * Hp + h) * ldB + b0 + r;
                dpT_hi[o] = hi;
                dpT_lo[o] = lo;
            }
        }
    } else {
        for (int idx = t; idx < MID_R * H; idx += 256) {
            int rr = idx / H, h = idx - rr * H;
            if (b0 + rr < B) pre_s[(int64_t)(b0 + rr) * H + h] = tile[rr * HP1 + h];
        }
    }
    double tot = block_sum<double>((double)ll, red);
    if (t == 0) atomicAdd(loss, -tot * (double)inv_S);
}


// ---------------------------------------------------------------------------------------------------
// "mid" stage of the tcgen05 variant, instruction-lean version: one CTA = 128 batch rows of one sample, TWO threads
// per row (256 threads; thread `half` owns one half of the hidden units).  `pre` arrives transposed [S][H][B] and
// is read straight from global memory (coalesced along the batch axis, software-pipelined, no staging pass); the
// hidden activations live in a [H][129] shared tile (conflict-free both per row and per hidden unit); W2 is kept
// TRANSPOSED and padded to CT classes ([H][CT], zero pad) so one hidden unit's column is CT/4 broadcast LDS.128
// instead of C scalar loads, and all multiply-adds are packed FFMA2 (two fp32 FMAs per issue slot, sm_100).
// CT = C rounded up to a multiple of 4.
//   P1  h = tanh(pre + b1), a = W2 h            P2  dW2[c,h] = sum_r da[r,c] h[r,h], db2
//   P3  dpre = (W2^T da)(1 - h^2) -> TF32 split, transposed store   P4  db1
// Pad rows h in [H, Hp) of dpT are NOT written: they only feed accumulator columns the GEMM epilogue discards.
// ---------------------------------------------------------------------------------------------------
constexpr int MID3_R = 128, MID3_TP = 129, MID3_THREADS = 256;

struct Mid3Smem {
    size_t tile, w2t, b1, b2, das, apart, total;   // offsets in floats
    __host__ __device__ Mid3Smem(int H, int CT) {
        tile = 0;
        w2t = ((size_t)H * MID3_TP + 3) / 4 * 4;
        b1 = w2t + (size_t)H * CT;
        b2 = b1 + ((size_t)H + 3) / 4 * 4;
        das = b2 + CT;
        apart = das + (size_t)MID3_R * CT;
        total = apart + (size_t)MID3_R * CT;
    }
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int CT>
__global__ void __launch_bounds__(MID3_THREADS)
bnn_mid3_kernel(const float* __restrict__ pre, const float* __restrict__ W, float* __restrict__ dW,
                const int32_t* __restrict__ y, BnnLayout L, float inv_S, double* __restrict__ loss,
                float* __restrict__ dpT_hi, float* __restrict__ dpT_lo, int Hp, int64_t ldB) {
    extern __shared__ __align__(16) float sm[];
    const int H = L.H, C = L.C, B = L.B;
    const Mid3Smem M(H, CT);
    float* tile = sm + M.tile;
    float* W2t = sm + M.w2t;
    float* b1s = sm + M.b1;
    float* b2s = sm + M.b2;
    float* das = sm + M.das;
    float* apart = sm + M.apart;
    __shared__ double red[32];
    constexpr int C4 = CT / 4, C2 = CT / 2;

    const int s = blockIdx.y, b0 = blockIdx.x * MID3_R, t = threadIdx.x;
    const int r = t & (MID3_R - 1), half = t >> 7;            // half is warp-uniform
    const float* Ws = W + (int64_t)s * L.ldw;
    float* dWs = dW + (int64_t)s * L.ldw;
    const bool row_ok = b0 + r < B;
    const float* pcol = pre + (int64_t)s * B * H + (row_ok ? b0 + r : 0);
    const int Hh = (H + 1) >> 1, hbeg = half * Hh, hend = min(H, hbeg + Hh);

    for (int idx = t; idx < H * CT; idx += MID3_THREADS) {
        const int h = idx / CT, c = idx - h * CT;
        W2t[idx] = c < C ? Ws[L.oW2 + (int64_t)c * H + h] : 0.f;
    }
    for (int idx = t; idx < H; idx += MID3_THREADS) b1s[idx] = Ws[L.ob1 + idx];
    if (t < CT) b2s[t] = t < C ? Ws[L.ob2 + t] : 0.f;
    __syncthreads();

    // ---- P1 (this thread: hidden units [hbeg, hend) of row r)
    float2 a2[C2];
#pragma unroll
    for (int c = 0; c < C2; ++c) a2[c] = make_float2(0.f, 0.f);
    // software pipeline over chunks of PF hidden units: the (latency-bound) global loads of chunk k+1 are in flight
    // while chunk k is evaluated; inside a chunk the PF tanh chains are independent (branch-free body -> ILP).
    constexpr int PF = 10;
    auto p1_unit = [&](int h, float xv) {
        const float v = tanhf(xv + b1s[h]);
        tile[h * MID3_TP + r] = v;
        const float2 v2 = make_float2(v, v);
        const float4* w4 = reinterpret_cast<const float4*>(W2t + h * CT);
#pragma unroll
        for (int q = 0; q < C4; ++q) {
            const float4 w = w4[q];
            a2[2 * q + 0] = ffma2(make_float2(w.x, w.y), v2, a2[2 * q + 0]);
            a2[2 * q + 1] = ffma2(make_float2(w.z, w.w), v2, a2[2 * q + 1]);
        }
    };
    {
        const int nmain = (hend - hbeg) / PF * PF, hmain = hbeg + nmain;
        float x[PF];
        if (nmain > 0) {
#pragma unroll
            for (int j = 0; j < PF; ++j) x[j] = pcol[(int64_t)(hbeg + j) * B];
        }
        for (int h0 = hbeg; h0 < hmain; h0 += PF) {
            float xn[PF];
            if (h0 + PF < hmain) {
#pragma unroll
                for (int j = 0; j < PF; ++j) xn[j] = pcol[(int64_t)(h0 + PF + j) * B];
            }
#pragma unroll
            for (int j = 0; j < PF; ++j) p1_unit(h0 + j, x[j]);
#pragma unroll
            for (int j = 0; j < PF; ++j) x[j] = xn[j];
        }
        for (int h = hmain; h < hend; ++h) p1_unit(h, pcol[(int64_t)h * B]);
    }
    if (half == 1) {
#pragma unroll
        for (int q = 0; q < C4; ++q)
            reinterpret_cast<float4*>(apart + r * CT)[q] = make_float4(a2[2 * q].x, a2[2 * q].y, a2[2 * q + 1].x, a2[2 * q + 1].y);
    }
    __syncthreads();
    float da[CT];
    float ll = 0.f;
    if (half == 0) {
        float a[CT];
#pragma unroll
        for (int q = 0; q < C4; ++q) {
            const float4 o = reinterpret_cast<const float4*>(apart + r * CT)[q];
            a[4 * q + 0] = a2[2 * q].x + o.x; a[4 * q + 1] = a2[2 * q].y + o.y;
            a[4 * q + 2] = a2[2 * q + 1].x + o.z; a[4 * q + 3] = a2[2 * q + 1].y + o.w;
        }
        if (row_ok) {
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (c < C) { a[c] += b2s[c]; m = fmaxf(m, a[c]); }
            float se = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (c < C) se += expf(a[c] - m);
            const float lse = m + logf(se);
            const int label = y[b0 + r];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                if (c < C) {
                    da[c] = (c == label ? 1.f : 0.f) - expf(a[c] - lse);      // d ll / d a_c
                    if (c == label) ll = a[c] - lse;
                } else da[c] = 0.f;
            }
        } else {
#pragma unroll
            for (int c = 0; c < CT; ++c) da[c] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < C4; ++q)
            reinterpret_cast<float4*>(das + r * CT)[q] = make_float4(da[4 * q], da[4 * q + 1], da[4 * q + 2], da[4 * q + 3]);
    }
    __syncthreads();
    if (half == 1) {
#pragma unroll
        for (int q = 0; q < C4; ++q) {
            const float4 d = reinterpret_cast<const float4*>(das + r * CT)[q];
            da[4 * q] = d.x; da[4 * q + 1] = d.y; da[4 * q + 2] = d.z; da[4 * q + 3] = d.w;
        }
    }

    // ---- P2: thread = (hidden unit hh, class half g): classes [g*C2, g*C2 + C2)
    for (int idx = t; idx < 2 * H; idx += MID3_THREADS) {
        const int hh = idx >> 1, g = idx & 1;
        float2 acc[C4];
#pragma unroll
        for (int q = 0; q < C4; ++q) acc[q] = make_float2(0.f, 0.f);
        const float* dbase = das + g * C2;
#pragma unroll 4
        for (int rr = 0; rr < MID3_R; ++rr) {
            const float hv = tile[hh * MID3_TP + rr];
            const float2 hv2 = make_float2(hv, hv);
            const float2* d2 = reinterpret_cast<const float2*>(dbase + rr * CT);
#pragma unroll
            for (int q = 0; q < C4; ++q) acc[q] = ffma2(d2[q], hv2, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < C4; ++q) {
            const int c = g * C2 + 2 * q;
            if (c < C) atomicAdd(&dWs[L.oW2 + (int64_t)c * H + hh], acc[q].x);
            if (c + 1 < C) atomicAdd(&dWs[L.oW2 + (int64_t)(c + 1) * H + hh], acc[q].y);
        }
    }
    if (t < C) {
        float acc = 0.f;
        for (int rr = 0; rr < MID3_R; ++rr) acc += das[rr * CT + t];
        atomicAdd(&dWs[L.ob2 + t], acc);
    }
    __syncthreads();

    // ---- P3: thread = (row, hidden half) again
    {
        float* ohi = dpT_hi + (int64_t)s * Hp * ldB + b0 + r;
        float* olo = dpT_lo + (int64_t)s * Hp * ldB + b0 + r;
        float2 da2[C2];
#pragma unroll
        for (int c = 0; c < C2; ++c) da2[c] = make_float2(da[2 * c], da[2 * c + 1]);
        ohi += (int64_t)hbeg * ldB;          // walk the output rows by pointer increments (no 64-bit multiply per h)
        olo += (int64_t)hbeg * ldB;
        auto p3_unit = [&](int h) {
            const float hv = tile[h * MID3_TP + r];
            const float4* w4 = reinterpret_cast<const float4*>(W2t + h * CT);
            float2 p0 = make_float2(0.f, 0.f), p1 = make_float2(0.f, 0.f);     // two independent packed chains
#pragma unroll
            for (int q = 0; q < C4; ++q) {
                const float4 w = w4[q];
                p0 = ffma2(da2[2 * q + 0], make_float2(w.x, w.y), p0);
                p1 = ffma2(da2[2 * q + 1], make_float2(w.z, w.w), p1);
            }
            const float dh = (p0.x + p0.y) + (p1.x + p1.y);
            const float dp = dh * (1.f - hv * hv);
            tile[h * MID3_TP + r] = dp;
            float hi, lo;
            umma::split_tf32(dp, hi, lo);
            if (row_ok) {
                *ohi = hi;
                *olo = lo;
            }
            ohi += ldB;
            olo += ldB;
        };
        const int h4 = hbeg + (hend - hbeg) / 4 * 4;
        for (int h0 = hbeg; h0 < h4; h0 += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) p3_unit(h0 + j);
        }
        for (int h = h4; h < hend; ++h) p3_unit(h);
    }
    __syncthreads();

    // ---- P4: thread = (hidden unit, row half)
    for (int idx = t; idx < 2 * H; idx += MID3_THREADS) {
        const int hh = idx >> 1, r0 = (idx & 1) * (MID3_R / 2);
        float acc = 0.f;
#pragma unroll 8
        for (int rr = r0; rr < r0 + MID3_R / 2; ++rr) acc += tile[hh * MID3_TP + rr];
        atomicAdd(&dWs[L.ob1 + hh], acc);
    }
    double tot = block_sum<double>((double)ll, red);
    if (t == 0) atomicAdd(loss, -tot * (double)inv_S);
}

template <int CT>
static int launch_mid3(const float* pre, const float* W, float* dW, const int32_t* y, const BnnLayout& L, int S, float inv_S,
                       double* loss, float* dph, float* dpl, int Hp, int64_t ldB, cudaStream_t stream) {
    const size_t smem = Mid3Smem(L.H, CT).total * sizeof(float);
    BRN_CUDA_OK(cudaFuncSetAttribute(bnn_mid3_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((L.B + MID3_R - 1) / MID3_R, S);
    bnn_mid3_kernel<CT><<<grid, MID3_THREADS, smem, stream>>>(pre, W, dW, y, L, inv_S, loss, dph, dpl, Hp, ldB);
    BRN_LAUNCH_OK("bnn_mid3_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// "mid" stage of the tcgen05 variant on the warp-level tensor cores (mma.sync m16n8k8 tf32, 3xTF32 split): the three
// small per-sample contractions of layer 2 are matrix products with one tiny dimension (C <= 16),
//   A  a[b, c]    = sum_h h[b, h] W2[c, h]           M = rows,  N = 16 classes, K = HP hidden
//   B  dh[b, h]   = sum_c da[b, c] W2[c, h]          M = rows,  N = HP hidden,  K = 16 classes
//   C  dW2[c, h]  = sum_b da[b, c] h[b, h]           M = 16 classes, N = HP hidden, K = 128 rows
// and cost ~10 k thread instructions per row as scalar FMAs (bnn_mid3_kernel: issue-bound, 140 us at C3).  Here one CTA
// = 128 batch rows of one sample, 8 warps, warp w = rows [16w, 16w+16) = one MMA m-tile.  The fragment layouts are
// chained without any shuffles by permuting the contraction index: the k-columns (t, t+4) of an A fragment are mapped
// to the consecutive pair (2t, 2t+1) of hidden units (phase A) / classes (phase B), which is exactly how the C fragment
// of the previous product holds them.  Only phase C needs transposed operands and goes through shared memory
// (tile[h][b] pitch 132, das[b][c] pitch 24: both conflict-free for the fragment loads).
// Accuracy: fp32-equivalent via the 3-product split; the hi*hi products and the two correction products accumulate in
// separate chains of at most 16 MMAs.
// ---------------------------------------------------------------------------------------------------
constexpr int MID4_R = 128, MID4_TP = 132, MID4_DP = 24, MID4_THREADS = 256;

template <int HP>
struct Mid4Smem {
    static constexpr int KS = HP / 8;
    // offsets in floats (all multiples of 4)
    static constexpr size_t fragA = 0;                                    // [KS][2][32] float4
    static constexpr size_t fragB = fragA + (size_t)KS * 2 * 32 * 4;      // [KS][2][32] float4
    static constexpr size_t tile = fragB + (size_t)KS * 2 * 32 * 4;       // [HP][132]
    static constexpr size_t das = tile + (size_t)HP * MID4_TP;            // [128][24] TF32 hi part of da
    static constexpr size_t das_lo = das + (size_t)MID4_R * MID4_DP;      // [128][24]
    static constexpr size_t b1 = das_lo + (size_t)MID4_R * MID4_DP;       // [HP]
    static constexpr size_t db1 = b1 + HP;                                // [8 warps][HP] per-warp column sums of dpre
    static constexpr size_t b2 = db1 + 8 * HP;                            // [16]
    static constexpr size_t total = b2 + 16;
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// tanh with ~1e-7 absolute error, 13 issue slots, branch-free: odd polynomial below 0.25 (truncation < 1e-8 relative),
// 1 - 2 / (1 + e^{2x}) above (ex2.approx / rcp.approx: absolute error ~2e-7 on a value >= 0.24).
__device__ __forceinline__ float tanh_fast(float x) {
    const float x2 = x * x;
    float p = 0.021869488536155203f;                 //  62/2835
    p = __fmaf_rn(p, x2, -0.053968253968253971f);    // -17/315
    p = __fmaf_rn(p, x2, 0.13333333333333333f);      //   2/15
    p = __fmaf_rn(p, x2, -0.33333333333333333f);     //  -1/3
    const float small = __fmaf_rn(x * x2, p, x);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    const float big = 1.f - __fdividef(2.f, 1.f + e);
    return fabsf(x) < 0.25f ? small : big;
}

// hi = x rounded to TF32; lo = the exact remainder, NOT re-rounded: mma.sync reads only the TF32 bits of an operand
// register, i.e. truncates lo (relative error <= 2^-21 of x, sign uncorrelated with x) -- used for mma.sync operands only.
__device__ __forceinline__ void split_tf32_trunc_lo(float x, float& hi, float& lo) {
    hi = umma::rn_tf32(x);
    lo = x - hi;
}

template <int HP, bool FULL>      // FULL: B is a multiple of 128 (no row guards anywhere)
__global__ void __launch_bounds__(MID4_THREADS, 2)
bnn_mid4_kernel(const float* __restrict__ pre, const float* __restrict__ W, float* __restrict__ dW,
                const int32_t* __restrict__ y, BnnLayout L, float inv_S, double* __restrict__ loss,
                float* __restrict__ dpT_hi, float* __restrict__ dpT_lo, int64_t ldB) {
    using M = Mid4Smem<HP>;
    constexpr int KS = M::KS;
    extern __shared__ __align__(16) float sm[];
    float4* fragA = reinterpret_cast<float4*>(sm + M::fragA);
    float4* fragB = reinterpret_cast<float4*>(sm + M::fragB);
    float* tile = sm + M::tile;
    float* das_hi = sm + M::das;
    float* das_lo = sm + M::das_lo;
    float* b1s = sm + M::b1;
    float* db1s = sm + M::db1;
    float* b2s = sm + M::b2;
    __shared__ double red[32];

    const int H = L.H, C = L.C, B = L.B;
    const int s = blockIdx.y, b0 = blockIdx.x * MID4_R, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const float* Ws = W + (int64_t)s * L.ldw;
    float* dWs = dW + (int64_t)s * L.ldw;
    const float* W2 = Ws + L.oW2;

    // ---- phase A loads first (their latency overlaps the weight staging): pre[h][row] for this thread's 2 rows x 2*KS
    // hidden units.  Rows >= B and hidden units >= H are CLAMPED to valid addresses, not masked: invalid rows get da = 0
    // below, and hidden units >= H meet zero weights in both fragment sets and are never stored.
    const int r0 = 16 * warp + g, r1 = r0 + 8;                 // this thread's two rows inside the CTA block
    const bool ok0 = FULL || b0 + r0 < B, ok1 = FULL || b0 + r1 < B;
    float hA[KS][4];
    {
        const float* pc0 = pre + (int64_t)s * B * H + (ok0 ? b0 + r0 : B - 1);
        const float* pc1 = pre + (int64_t)s * B * H + (ok1 ? b0 + r1 : B - 1);
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const int h0 = min(8 * k + 2 * t, H - 1), h1 = min(8 * k + 2 * t + 1, H - 1);
            hA[k][0] = pc0[(int64_t)h0 * B];
            hA[k][1] = pc1[(int64_t)h0 * B];
            hA[k][2] = pc0[(int64_t)h1 * B];
            hA[k][3] = pc1[(int64_t)h1 * B];
        }
    }

    // ---- stage the per-sample layer-2 weights as ready-made (hi, lo) B fragments
    for (int idx = tid; idx < KS * 2 * 32; idx += MID4_THREADS) {
        const int ln = idx & 31, q = (idx >> 5) & 1, k = idx >> 6, gg = ln >> 2, tt = ln & 3;
        {   // phase A: B[k = hidden, n = class]: b0 = W2[8q + g][8k + 2t], b1 = W2[8q + g][8k + 2t + 1]
            const int c = 8 * q + gg, h0 = 8 * k + 2 * tt;
            const float w0 = (c < C && h0 < H) ? W2[(int64_t)c * H + h0] : 0.f;
            const float w1 = (c < C && h0 + 1 < H) ? W2[(int64_t)c * H + h0 + 1] : 0.f;
            float4 f;
            umma::split_tf32(w0, f.x, f.z);
            umma::split_tf32(w1, f.y, f.w);
            fragA[idx] = f;
        }
        {   // phase B: B[k = class, n = hidden]: b0 = W2[8q + 2t][8k + g], b1 = W2[8q + 2t + 1][8k + g]   (k = n-tile j)
            const int c0 = 8 * q + 2 * tt, h = 8 * k + gg;
            const float w0 = (c0 < C && h < H) ? W2[(int64_t)c0 * H + h] : 0.f;
            const float w1 = (c0 + 1 < C && h < H) ? W2[(int64_t)(c0 + 1) * H + h] : 0.f;
            float4 f;
            umma::split_tf32(w0, f.x, f.z);
            umma::split_tf32(w1, f.y, f.w);
            fragB[idx] = f;
        }
    }
    for (int idx = tid; idx < HP; idx += MID4_THREADS) b1s[idx] = idx < H ? Ws[L.ob1 + idx] : 0.f;
    if (tid < 16) b2s[tid] = tid < C ? Ws[L.ob2 + tid] : 0.f;
    __syncthreads();

    // ---- phase A: h = tanh(pre + b1) (kept in A-fragment registers and in the smem tile), a = h W2^T
    float ahh[2][4], acr[2][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) ahh[q][i] = acr[q][i] = 0.f;
    {
        float* tp = tile + (2 * t) * MID4_TP + r0;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const float2 bb = *reinterpret_cast<const float2*>(b1s + 8 * k + 2 * t);
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = tanh_fast(hA[k][i] + (i < 2 ? bb.x : bb.y));
                hA[k][i] = v;
                split_tf32_trunc_lo(v, hi[i], lo[i]);
            }
            tp[k * 8 * MID4_TP] = hA[k][0];
            tp[k * 8 * MID4_TP + 8] = hA[k][1];
            tp[k * 8 * MID4_TP + MID4_TP] = hA[k][2];
            tp[k * 8 * MID4_TP + MID4_TP + 8] = hA[k][3];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 f = fragA[(k * 2 + q) * 32 + lane];
                mma_tf32(ahh[q], hi, f.x, f.y);
                mma_tf32(acr[q], lo, f.x, f.y);
                mma_tf32(acr[q], hi, f.z, f.w);
            }
        }
    }

    // ---- log-softmax over the classes: row r0 holds classes {2t, 2t+1, 8+2t, 9+2t} in a[q][0..1], row r1 in a[q][2..3]
    float dahi[2][4], dalo[2][4];            // phase-B A fragments: (r0, class 2t), (r1, 2t), (r0, 2t+1), (r1, 2t+1)
    float ll = 0.f;
    {
        const int lab0 = ok0 ? y[b0 + r0] : -1, lab1 = ok1 ? y[b0 + r1] : -1;
        float a[2][4];
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = 8 * q + 2 * t + (i & 1);
                a[q][i] = c < C ? ahh[q][i] + acr[q][i] + b2s[c] : -INFINITY;
                if (i < 2) m0 = fmaxf(m0, a[q][i]);
                else m1 = fmaxf(m1, a[q][i]);
            }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float ex[2][4];
        float se0 = 0.f, se1 = 0.f;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ex[q][i] = expf(a[q][i] - (i < 2 ? m0 : m1));            // exp(-inf) = 0 for the pad classes
                if (i < 2) se0 += ex[q][i];
                else se1 += ex[q][i];
            }
        se0 += __shfl_xor_sync(0xffffffffu, se0, 1); se0 += __shfl_xor_sync(0xffffffffu, se0, 2);
        se1 += __shfl_xor_sync(0xffffffffu, se1, 1); se1 += __shfl_xor_sync(0xffffffffu, se1, 2);
        const float lse0 = m0 + logf(se0), lse1 = m1 + logf(se1);
        const float inv0 = 1.f / se0, inv1 = 1.f / se1;
        float da[2][4];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = 8 * q + 2 * t + (i & 1);
                const bool ok = (i < 2 ? ok0 : ok1) && c < C;
                const int lab = i < 2 ? lab0 : lab1;
                const float lse = i < 2 ? lse0 : lse1;
                const float sm_ = ex[q][i] * (i < 2 ? inv0 : inv1);
                da[q][i] = ok ? (c == lab ? 1.f : 0.f) - sm_ : 0.f;              // d ll / d a_c
                if (ok && c == lab) ll += a[q][i] - lse;
            }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            split_tf32_trunc_lo(da[q][0], dahi[q][0], dalo[q][0]);
            split_tf32_trunc_lo(da[q][2], dahi[q][1], dalo[q][1]);
            split_tf32_trunc_lo(da[q][1], dahi[q][2], dalo[q][2]);
            split_tf32_trunc_lo(da[q][3], dahi[q][3], dalo[q][3]);
            *reinterpret_cast<float2*>(das_hi + r0 * MID4_DP + 8 * q + 2 * t) = make_float2(dahi[q][0], dahi[q][2]);
            *reinterpret_cast<float2*>(das_hi + r1 * MID4_DP + 8 * q + 2 * t) = make_float2(dahi[q][1], dahi[q][3]);
            *reinterpret_cast<float2*>(das_lo + r0 * MID4_DP + 8 * q + 2 * t) = make_float2(dalo[q][0], dalo[q][2]);
            *reinterpret_cast<float2*>(das_lo + r1 * MID4_DP + 8 * q + 2 * t) = make_float2(dalo[q][1], dalo[q][3]);
        }
    }

    // ---- phase B: dh = da W2, dpre = dh (1 - h^2) -> TF32 split, transposed store; db1 column sums
    {
        // output pointers walk down the hidden axis: element (h0 = 2t [+1], row r0 [+8]) of this sample's block
        float* ohi0 = dpT_hi + ((int64_t)s * HP + 2 * t) * ldB + b0 + r0;
        float* olo0 = dpT_lo + ((int64_t)s * HP + 2 * t) * ldB + b0 + r0;
        float* ohi1 = ohi0 + ldB;
        float* olo1 = olo0 + ldB;
        const int64_t step = 8 * ldB;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            float dhh[4] = {0.f, 0.f, 0.f, 0.f}, dcr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 f = fragB[(j * 2 + q) * 32 + lane];
                mma_tf32(dhh, dahi[q], f.x, f.y);
                mma_tf32(dcr, dalo[q], f.x, f.y);
                mma_tf32(dcr, dahi[q], f.z, f.w);
            }
            // C fragment: (r0, h0), (r0, h0+1), (r1, h0), (r1, h0+1) with h0 = 8j + 2t  <->  hA[j][0], [2], [1], [3]
            const int h0 = 8 * j + 2 * t;
            float dp[4];
            dp[0] = (dhh[0] + dcr[0]) * __fmaf_rn(-hA[j][0], hA[j][0], 1.f);
            dp[1] = (dhh[1] + dcr[1]) * __fmaf_rn(-hA[j][2], hA[j][2], 1.f);
            dp[2] = (dhh[2] + dcr[2]) * __fmaf_rn(-hA[j][1], hA[j][1], 1.f);
            dp[3] = (dhh[3] + dcr[3]) * __fmaf_rn(-hA[j][3], hA[j][3], 1.f);
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) umma::split_tf32(dp[i], hi[i], lo[i]);
            // hidden units >= H: dp == 0 there (zero fragB weights) and the pad rows of dpT exist -> no guard needed
            if (ok0) { ohi0[0] = hi[0]; olo0[0] = lo[0]; ohi1[0] = hi[1]; olo1[0] = lo[1]; }
            if (ok1) { ohi0[8] = hi[2]; olo0[8] = lo[2]; ohi1[8] = hi[3]; olo1[8] = lo[3]; }
            ohi0 += step; olo0 += step; ohi1 += step; olo1 += step;
            float c0 = dp[0] + dp[2], c1 = dp[1] + dp[3];        // column sums over this warp's 16 rows
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                c0 += __shfl_xor_sync(0xffffffffu, c0, o);
                c1 += __shfl_xor_sync(0xffffffffu, c1, o);
            }
            if (g == 0) *reinterpret_cast<float2*>(db1s + warp * HP + h0) = make_float2(c0, c1);   // per-warp partial
        }
    }
    __syncthreads();

    // ---- phase C: dW2[c, h] = sum_b da[b, c] h[b, h] over the CTA's 128 rows; warp w owns hidden n-tiles w and w + 8
    auto phase_c = [&](auto ntag) {
        constexpr int NT = decltype(ntag)::value;
        float chh[NT][4], ccr[NT][4];
#pragma unroll
        for (int q = 0; q < NT; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) chh[q][i] = ccr[q][i] = 0.f;
        const float* dh_ = das_hi + t * MID4_DP + g;
        const float* dl_ = das_lo + t * MID4_DP + g;
        const float* tb = tile + (8 * warp + g) * MID4_TP + t;
#pragma unroll 4
        for (int kb = 0; kb < MID4_R / 8; ++kb) {
            // A = da^T: (class g, row 8kb+t), (class g+8, row 8kb+t), (class g, row 8kb+t+4), (class g+8, row 8kb+t+4)
            float ahi[4], alo[4];
            ahi[0] = dh_[kb * 8 * MID4_DP];
            ahi[1] = dh_[kb * 8 * MID4_DP + 8];
            ahi[2] = dh_[(kb * 8 + 4) * MID4_DP];
            ahi[3] = dh_[(kb * 8 + 4) * MID4_DP + 8];
            alo[0] = dl_[kb * 8 * MID4_DP];
            alo[1] = dl_[kb * 8 * MID4_DP + 8];
            alo[2] = dl_[(kb * 8 + 4) * MID4_DP];
            alo[3] = dl_[(kb * 8 + 4) * MID4_DP + 8];
#pragma unroll
            for (int q = 0; q < NT; ++q) {
                float bh0, bl0, bh1, bl1;
                split_tf32_trunc_lo(tb[q * 64 * MID4_TP + kb * 8], bh0, bl0);          // B = h: (row 8kb+t, hidden 8j+g)
                split_tf32_trunc_lo(tb[q * 64 * MID4_TP + kb * 8 + 4], bh1, bl1);
                mma_tf32(chh[q], ahi, bh0, bh1);
                mma_tf32(ccr[q], alo, bh0, bh1);
                mma_tf32(ccr[q], ahi, bl0, bl1);
            }
        }
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            const int h0 = 8 * (warp + 8 * q) + 2 * t;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = g + (i >= 2 ? 8 : 0), h = h0 + (i & 1);
                if (c < C && h < H) atomicAdd(&dWs[L.oW2 + (int64_t)c * H + h], chh[q][i] + ccr[q][i]);
            }
        }
    };
    if (warp + 8 < KS) phase_c(std::integral_constant<int, 2>());
    else if (warp < KS) phase_c(std::integral_constant<int, 1>());
    if (warp == 7 && lane < C) {
        float acc = 0.f;
        for (int rr = 0; rr < MID4_R; ++rr) acc += das_hi[rr * MID4_DP + lane] + das_lo[rr * MID4_DP + lane];
        atomicAdd(&dWs[L.ob2 + lane], acc);
    }
    for (int idx = tid; idx < H; idx += MID4_THREADS) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < MID4_THREADS / 32; ++w) acc += db1s[w * HP + idx];
        atomicAdd(&dWs[L.ob1 + idx], acc);
    }
    double tot = block_sum<double>((double)ll, red);
    if (tid == 0) atomicAdd(loss, -tot * (double)inv_S);
}

template <int HP>
static int launch_mid4(const float* pre, const float* W, float* dW, const int32_t* y, const BnnLayout& L, int S, float inv_S,
                       double* loss, float* dph, float* dpl, int64_t ldB, cudaStream_t stream) {
    const size_t smem = Mid4Smem<HP>::total * sizeof(float);
    dim3 grid((L.B + MID4_R - 1) / MID4_R, S);
    if (L.B % MID4_R == 0) {
        BRN_CUDA_OK(cudaFuncSetAttribute(bnn_mid4_kernel<HP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bnn_mid4_kernel<HP, true><<<grid, MID4_THREADS, smem, stream>>>(pre, W, dW, y, L, inv_S, loss, dph, dpl, ldB);
    } else {
        BRN_CUDA_OK(cudaFuncSetAttribute(bnn_mid4_kernel<HP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bnn_mid4_kernel<HP, false><<<grid, MID4_THREADS, smem, stream>>>(pre, W, dW, y, L, inv_S, loss, dph, dpl, ldB);
    }
    BRN_LAUNCH_OK("bnn_mid4_kernel");
    return 0;
}

}  // namespace brn
#include "bnn_fused.cuh"
namespace brn {

// Fused noise + weight sampling + TF32 split for layer 1 (P % 4 == 0), sample-group version: one thread = 4 consecutive
// weights, walking SG consecutive samples.  sigma = softplus(rho) is evaluated once per thread (no separate pass), and the
// sample-axis noise statistics the closed-form prior / entropy terms need,  e1 = sum_s eps,  e2 = sum_s eps^2,  are
// accumulated in registers and added to e1/e2 [H*P] with one 16-byte RED each per thread -- the statistics stage no
// longer has to re-read the noise.
template <int SG>
__global__ void __launch_bounds__(256)
sample_w1_group_kernel(const float* __restrict__ mu, const float* __restrict__ rho, const float* __restrict__ eps_in,
                       int64_t lde_in, float* __restrict__ eps_out, int64_t lde_out, float* __restrict__ hi,
                       float* __restrict__ lo, int H, int P, int Hp, int64_t ldP, brn_sample_range r, uint32_t var_id,
                       float* __restrict__ e1, float* __restrict__ e2) {
    const int64_t qq = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qq * 4 >= (int64_t)H * P) return;      // pad rows h in [H, Hp) are never written (their accumulator columns are masked)
    const int h = (int)((qq * 4) / P), p = (int)((qq * 4) - (int64_t)h * P);
    const int64_t i = (int64_t)h * P + p;
    const float4 m = *reinterpret_cast<const float4*>(mu + i);
    const float4 rh = *reinterpret_cast<const float4*>(rho + i);
    const float4 sg = make_float4(softplusf(rh.x), softplusf(rh.y), softplusf(rh.z), softplusf(rh.w));
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
    const int s_begin = blockIdx.y * SG, s_end = min(r.s_local, s_begin + SG);
    const int64_t o0 = (int64_t)h * ldP + p;
#pragma unroll 2
    for (int s = s_begin; s < s_end; ++s) {
        float4 e;
        if (eps_in) {
            e = *reinterpret_cast<const float4*>(eps_in + (int64_t)s * lde_in + i);
        } else {
            Normal4 n = philox_normal4(r.seed, r.offset, var_id, (uint32_t)(r.s0 + s), (uint32_t)(i >> 2));
            e = make_float4(n.v[0], n.v[1], n.v[2], n.v[3]);
        }
        *reinterpret_cast<float4*>(eps_out + (int64_t)s * lde_out + i) = e;
        float4 vh, vl;
        umma::split_tf32(__fmaf_rn(sg.x, e.x, m.x), vh.x, vl.x);
        umma::split_tf32(__fmaf_rn(sg.y, e.y, m.y), vh.y, vl.y);
        umma::split_tf32(__fmaf_rn(sg.z, e.z, m.z), vh.z, vl.z);
        umma::split_tf32(__fmaf_rn(sg.w, e.w, m.w), vh.w, vl.w);
        const int64_t o = (int64_t)s * Hp * ldP + o0;
        *reinterpret_cast<float4*>(hi + o) = vh;
        *reinterpret_cast<float4*>(lo + o) = vl;
        a1.x += e.x; a1.y += e.y; a1.z += e.z; a1.w += e.w;
        a2.x = __fmaf_rn(e.x, e.x, a2.x); a2.y = __fmaf_rn(e.y, e.y, a2.y);
        a2.z = __fmaf_rn(e.z, e.z, a2.z); a2.w = __fmaf_rn(e.w, e.w, a2.w);
    }
    red_add_v4(e1 + i, a1.x, a1.y, a1.z, a1.w);
    red_add_v4(e2 + i, a2.x, a2.y, a2.z, a2.w);
}

// e1 = sum_s eps, e2 = sum_s eps^2 from stored noise (shapes the vectorised sampler does not cover)
__global__ void __launch_bounds__(256)
eps_stats_kernel(const float* __restrict__ eps, int64_t lde, int64_t numel, int S, float* __restrict__ e1, float* __restrict__ e2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numel) return;
    float a1 = 0.f, a2 = 0.f;
    for (int s = 0; s < S; ++s) {
        const float e = eps[(int64_t)s * lde + i];
        a1 += e;
        a2 = __fmaf_rn(e, e, a2);
    }
    e1[i] += a1;
    e2[i] += a2;
}

// layer-1 finalisation of the fused path: gwT / gweT [P][HP] (backward GEMM epilogue) + e1 / e2 [H*P] (sampler) ->
// closed-form prior / entropy terms, chain rule to (mu, rho), loss.  One thread per weight, element order (h, p).
__global__ void __launch_bounds__(256)
bnn_w1_finalize_kernel(brn_mf_var v, const float* __restrict__ gwT, const float* __restrict__ gweT, int HP, int P,
                       const float* __restrict__ e1, const float* __restrict__ e2, brn_sample_range r, int with_prior,
                       double* __restrict__ loss) {
    __shared__ double red[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double elbo = 0.0;
    if (i < v.numel) {
        const int h = (int)(i / P), p = (int)(i - (int64_t)h * P);
        const int64_t o = (int64_t)p * HP + h;
        elbo = mf_finalize_element(v, i, gwT[o], gweT[o], e1[i], e2[i], r, with_prior);
    }
    const double tot = block_sum<double>(elbo, red);
    if (threadIdx.x == 0 && with_prior) atomicAdd(loss, -tot);
}

template <int HP, int BK>
static int launch_fwd_mid(const float* Ah, const float* Al, int M, int64_t lda, const float* Bh, const float* Bl, int N, int64_t ldb,
                          int K, int drain_chunks, const FwdMidParams& fp, cudaStream_t stream) {
    constexpr int BN = 2 * HP;
    CUtensorMap tAh, tAl, tBh, tBl;
    if (int e = make_tmap_2d_f32(&tAh, Ah, M, K, lda, UG_BM, BK)) return e;
    if (int e = make_tmap_2d_f32(&tAl, Al, M, K, lda, UG_BM, BK)) return e;
    if (int e = make_tmap_2d_f32(&tBh, Bh, N, K, ldb, BN, BK)) return e;
    if (int e = make_tmap_2d_f32(&tBl, Bl, N, K, ldb, BN, BK)) return e;
    const int m_tiles = (M + UG_BM - 1) / UG_BM, n_tiles = (N + BN - 1) / BN, k_chunks = (K + BK - 1) / BK;
    drain_chunks = drain_chunks * 32 / BK;
    if (drain_chunks < 1) drain_chunks = 2 * 32 / BK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int U = m_tiles * n_tiles, grid = U < sms ? U : sms;
    auto kern = bnn_fwd_mid_kernel<HP, BK>;
    const int smem = FwdMidSmem<HP, BK>::TOTAL;
    BRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, FM_THREADS, smem, stream>>>(tAh, tAl, tBh, tBl, m_tiles, n_tiles, k_chunks, drain_chunks, fp);
    BRN_LAUNCH_OK("bnn_fwd_mid_kernel");
    return 0;
}

static size_t bnn_mid_smem(int R, int H, int C) {
    return sizeof(float) * ((size_t)R * (H + 1) + (size_t)C * H + H + C + (size_t)R * C);
}

constexpr int BNN_UMMA_HP = 104;     // padded hidden width of the instantiated tcgen05 variant (multiple of 8)
constexpr int BNN_UMMA_BK = 16;      // K chunk: 16 floats (64-byte swizzle), 5-stage TMA ring
constexpr int BNN_UMMA_NSAMP = 2;    // samples per MMA N tile (N = 208)

struct BnnWorkspace {
    float *eps, *W, *dW, *pre, *stats, *sigma;
    float *Xh, *Xl, *Xth, *Xtl, *Wh, *Wl, *dph, *dpl;   // tcgen05 variant: TF32-split operands
    float *gwT, *gweT, *e1, *e2;                        // fused path: sample-axis sums of layer 1 (one block, zeroed per call)
    size_t fused_stat_floats;
    int64_t ldP, ldB;
    size_t bytes;
    BnnWorkspace(void* base, const BnnLayout& L, int S) {
        size_t off = 0;
        auto take = [&](size_t nfloat) {
            float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
            off += (nfloat * sizeof(float) + 255) / 256 * 256;
            return p;
        };
        eps = take((size_t)S * L.ldw);
        W = take((size_t)S * L.ldw);
        dW = take((size_t)S * L.ldw);
        pre = take((size_t)S * L.B * L.H);
        stats = take(4 * (size_t)L.ldw);
        sigma = take(L.ldw);
        ldP = (L.P + 3) / 4 * 4;
        ldB = (L.B + 3) / 4 * 4;
        Xh = take((size_t)L.B * ldP); Xl = take((size_t)L.B * ldP);
        Xth = take((size_t)L.P * ldB); Xtl = take((size_t)L.P * ldB);
        const size_t rowsW = (size_t)(S + BNN_UMMA_NSAMP) * BNN_UMMA_HP;
        Wh = take(rowsW * ldP); Wl = take(rowsW * ldP);
        dph = take(rowsW * ldB); dpl = take(rowsW * ldB);
        const size_t gT = ((size_t)L.P * BNN_UMMA_HP + 63) / 64 * 64, ne = ((size_t)L.H * L.P + 63) / 64 * 64;
        gwT = take(2 * gT + 2 * ne);                     // contiguous: [gwT | gweT | e1 | e2]
        gweT = gwT ? gwT + gT : nullptr;
        e1 = gwT ? gweT + gT : nullptr;
        e2 = gwT ? e1 + ne : nullptr;
        fused_stat_floats = 2 * gT + 2 * ne;
        bytes = off;
    }
};

}  // namespace brn

using namespace brn;

extern "C" size_t brn_bnn_workspace_bytes(int B, int P, int H, int C, int s_local) {
    if (B <= 0 || P <= 0 || H <= 0 || C <= 0 || s_local < 0) return 0;
    BnnLayout L(B, P, H, C);
    return BnnWorkspace(nullptr, L, s_local).bytes;
}

extern "C" int brn_bnn_elbo_fwd_bwd(const float* X, const int32_t* y, int B, int P, int H, int C,
                                    const brn_mf_var vars[4], const brn_sample_range* r, void* workspace,
                                    size_t workspace_bytes, int with_prior, double* loss, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BRN_CHECK_ARG(X && y && vars && r && loss, "brn_bnn_elbo_fwd_bwd: NULL pointer");
    BRN_CHECK_ARG(B > 0 && P > 0 && H > 0 && C > 0, "brn_bnn_elbo_fwd_bwd: bad shape B=%d P=%d H=%d C=%d", B, P, H, C);
    BRN_CHECK_ARG(C <= BNN_MAXC, "brn_bnn_elbo_fwd_bwd: C=%d exceeds the supported maximum %d", C, BNN_MAXC);
    BRN_CHECK_ARG(r->s_local >= 0 && r->s_total > 0 && r->s0 >= 0 && r->s0 + r->s_local <= r->s_total,
                  "bad sample range s0=%d s_local=%d s_total=%d", r->s0, r->s_local, r->s_total);
    BnnLayout L(B, P, H, C);
    const int64_t numels[4] = {(int64_t)H * P, H, (int64_t)C * H, C};
    const int64_t offs[4] = {L.oW1, L.ob1, L.oW2, L.ob2};
    for (int v = 0; v < 4; ++v) {
        BRN_CHECK_ARG(vars[v].numel == numels[v], "vars[%d].numel=%lld, expected %lld", v, (long long)vars[v].numel,
                      (long long)numels[v]);
        BRN_CHECK_ARG(vars[v].mu && vars[v].rho && vars[v].dmu && vars[v].drho, "vars[%d]: NULL parameter pointer", v);
        BRN_CHECK_ARG(!with_prior || vars[v].tied || (vars[v].prior_loc && vars[v].prior_scale),
                      "vars[%d]: prior_loc/prior_scale required when not tied", v);
    }
    const int S = r->s_local;
    if (S == 0) return wait_data_ready((cudaStream_t)stream_);
    BnnWorkspace ws(workspace, L, S);
    BRN_CHECK_ARG(workspace && workspace_bytes >= ws.bytes, "workspace too small: %zu < %zu", workspace_bytes, ws.bytes);
    // variant: tcgen05 (TMA + 3xTF32 tensor-core GEMMs) when the padded hidden width matches the instantiated
    // tile, else the fp32 SIMT GEMMs.  BRN_BNN_VARIANT=simt|tcgen05 forces one (tests compare both).
    bool use_tc = (H > BNN_UMMA_HP - 16 && H <= BNN_UMMA_HP);
    if (const char* env = getenv("BRN_BNN_VARIANT")) {
        if (!strcmp(env, "simt")) use_tc = false;
        else if (!strcmp(env, "tcgen05")) {
            BRN_CHECK_ARG(H <= BNN_UMMA_HP, "BRN_BNN_VARIANT=tcgen05 needs H <= %d (got %d)", BNN_UMMA_HP, H);
            use_tc = true;
        }
    }
    set_variant(use_tc ? "tcgen05" : "simt");
    constexpr int HP = BNN_UMMA_HP, NS = BNN_UMMA_NSAMP, BN = HP * NS;
    // TMEM accumulation chain = 4 x 32 K elements between register drains (umma_gemm.cuh): measured on B200 the step is
    // 3.3 % faster than with 2 (profiles/r1g_*) and the whole K3 parity suite, incl. the full C3 shape, stays green.
    int drain = 4;
    if (const char* env = getenv("BRN_UMMA_DRAIN")) drain = atoi(env);

    // ------------------------------------------------------------------------------------------------
    // fused path (default for the tcgen05 variant; BRN_BNN_MID=4 / 3 select the staged pipelines below, kept for A/B
    // measurements and tests): sampler (+ noise statistics) -> forward GEMM with the mid stage in its epilogue ->
    // backward GEMM with the sample-axis reduction in its epilogue -> finalisation.  pre_s and dW1_s never reach HBM.
    // ------------------------------------------------------------------------------------------------
    int mid_sel = 5;
    if (const char* env = getenv("BRN_BNN_MID")) mid_sel = atoi(env);
    if (use_tc && mid_sel >= 5) {
        const float inv_S = 1.0f / (float)r->s_total;
        {
            StageTimer st("bnn.sample_weights", stream);
            BRN_CUDA_OK(cudaMemsetAsync(ws.gwT, 0, sizeof(float) * ws.fused_stat_floats, stream));
            const bool fast = (P % 4 == 0) && ((uintptr_t)vars[0].mu % 16 == 0) && ((uintptr_t)vars[0].rho % 16 == 0) &&
                              (!vars[0].eps || ((uintptr_t)vars[0].eps % 16 == 0));
            if (fast) {
                constexpr int SG = 8;
                dim3 grid((unsigned)(((int64_t)H * P / 4 + 255) / 256), (unsigned)((S + SG - 1) / SG));
                sample_w1_group_kernel<SG><<<grid, 256, 0, stream>>>(vars[0].mu, vars[0].rho, vars[0].eps, numels[0], ws.eps + offs[0],
                                                                     L.ldw, ws.Wh, ws.Wl, H, P, HP, ws.ldP, *r, vars[0].var_id, ws.e1,
                                                                     ws.e2);
                BRN_LAUNCH_OK("sample_w1_group_kernel");
            } else {
                if (vars[0].eps)
                    BRN_CUDA_OK(cudaMemcpy2DAsync(ws.eps + offs[0], L.ldw * sizeof(float), vars[0].eps, numels[0] * sizeof(float),
                                                  numels[0] * sizeof(float), S, cudaMemcpyDeviceToDevice, stream));
                else if (int e = launch_philox_fill(ws.eps + offs[0], L.ldw, numels[0], vars[0].var_id, *r, stream)) return e;
                dim3 grid((P + 255) / 256, HP, S);
                sample_w1_split_kernel<<<grid, 256, 0, stream>>>(vars[0].mu, vars[0].rho, ws.eps + offs[0], L.ldw, ws.Wh, ws.Wl, H, P,
                                                                 HP, ws.ldP);
                BRN_LAUNCH_OK("sample_w1_split_kernel");
                eps_stats_kernel<<<(unsigned)((numels[0] + 255) / 256), 256, 0, stream>>>(ws.eps + offs[0], L.ldw, numels[0], S, ws.e1,
                                                                                         ws.e2);
                BRN_LAUNCH_OK("eps_stats_kernel");
            }
            if (S % NS) {   // the odd tail tile reads one more sample block: keep it finite
                BRN_CUDA_OK(cudaMemsetAsync(ws.Wh + (size_t)S * HP * ws.ldP, 0, sizeof(float) * HP * ws.ldP, stream));
                BRN_CUDA_OK(cudaMemsetAsync(ws.Wl + (size_t)S * HP * ws.ldP, 0, sizeof(float) * HP * ws.ldP, stream));
                BRN_CUDA_OK(cudaMemsetAsync(ws.dph + (size_t)S * HP * ws.ldB, 0, sizeof(float) * HP * ws.ldB, stream));
                BRN_CUDA_OK(cudaMemsetAsync(ws.dpl + (size_t)S * HP * ws.ldB, 0, sizeof(float) * HP * ws.ldB, stream));
            }
            // small variables: noise, sampled values, and zeroed per-sample gradient slots (accumulated into by the epilogue)
            if (int e = launch_sample_multi(vars + 1, offs + 1, 3, ws.eps, ws.W, L.ldw, *r, stream, ws.dW)) return e;
            if (int e = wait_data_ready(stream)) return e;
            if (int e = launch_split_tf32(X, P, B, P, ws.Xh, ws.Xl, ws.ldP, ws.Xth, ws.Xtl, ws.ldB, stream)) return e;
        }
        {
            StageTimer st("bnn.gemm_fwd", stream);      // forward GEMM + mid in its epilogue
            FwdMidParams fp{ws.W, ws.dW, y, L, S, inv_S, loss, ws.dph, ws.dpl, ws.ldB};
            if (int e = launch_fwd_mid<HP, BNN_UMMA_BK>(ws.Xh, ws.Xl, B, ws.ldP, ws.Wh, ws.Wl, S * HP, ws.ldP, P, drain, fp, stream))
                return e;
        }
        {
            StageTimer st("bnn.gemm_bwd", stream);      // dW1_s = dpre_s^T . X, folded over samples in the epilogue
            EpiSampleReduce::Params ep{ws.gwT, ws.gweT, ws.eps + offs[0], L.ldw, P, H, HP, S};
            if (int e = launch_umma_nt<BN, BNN_UMMA_BK, EpiSampleReduce>(ws.Xth, ws.Xtl, P, ws.ldB, ws.dph, ws.dpl, S * HP, ws.ldB, B, 0,
                                                                         drain, ep, stream, true))
                return e;
        }
        StageTimer st5("bnn.reduce_finalize", stream);
        {
            int64_t offs_small[3] = {0, offs[2] - offs[1], offs[3] - offs[1]};
            if (int e = launch_mf_reduce_finalize_multi(vars + 1, offs_small, 3, L.numel - offs[1], ws.eps + offs[1], L.ldw,
                                                        ws.dW + offs[1], L.ldw, ws.stats, *r, with_prior, loss, stream, 0))
                return e;
        }
        bnn_w1_finalize_kernel<<<(unsigned)((numels[0] + 255) / 256), 256, 0, stream>>>(vars[0], ws.gwT, ws.gweT, HP, P, ws.e1, ws.e2,
                                                                                       *r, with_prior, loss);
        BRN_LAUNCH_OK("bnn_w1_finalize_kernel");
        return 0;
    }

    // 1. noise + weights.  Noise ends up in ws.eps [S][ldw] and all sampled weights in ws.W, the four variables back to
    //    back inside a row, so stage 5 is one launch over the concatenated range.  Optional (tcgen05 variant, Philox
    //    mode, BRN_BNN_REGEN_EPS=1): the layer-1 noise (98.6 % of the elements) is never stored -- stage 5 regenerates it
    //    from the same Philox counters (80 MB less written and read per evaluation at the C3 size).
    int64_t regen_numel0 = 0;
    {
        StageTimer st("bnn.sample_weights", stream);
        if (use_tc) {
            const bool fast = (P % 4 == 0) && ((uintptr_t)vars[0].mu % 16 == 0) &&
                              (!vars[0].eps || ((uintptr_t)vars[0].eps % 16 == 0));
            if (fast) {
                softplus_kernel<<<(unsigned)((numels[0] + 255) / 256), 256, 0, stream>>>(vars[0].rho, ws.sigma, numels[0]);
                BRN_LAUNCH_OK("softplus_kernel");
                dim3 grid((unsigned)(((int64_t)H * P / 4 + 255) / 256), S);
                // BRN_BNN_REGEN_EPS=1: do not store the Philox noise of layer 1, regenerate it in stage 5.  Measured on B200
                // (profiles/r1e_*): the sampler is issue-bound, not write-bound (47 us either way) and the regenerating
                // stats kernel is 5 us slower, so storing stays the default; the switch saves 80 MB of workspace.
                const char* rg = getenv("BRN_BNN_REGEN_EPS");
                const bool regen = !vars[0].eps && rg && atoi(rg);
                if (regen) regen_numel0 = numels[0];
                sample_w1_fused_kernel<<<grid, 256, 0, stream>>>(vars[0].mu, ws.sigma, vars[0].eps, numels[0],
                                                                 regen ? nullptr : ws.eps + offs[0], L.ldw, ws.Wh, ws.Wl, H, P,
                                                                 HP, ws.ldP, *r, vars[0].var_id);
                BRN_LAUNCH_OK("sample_w1_fused_kernel");
            } else {
                if (vars[0].eps)
                    BRN_CUDA_OK(cudaMemcpy2DAsync(ws.eps + offs[0], L.ldw * sizeof(float), vars[0].eps, numels[0] * sizeof(float),
                                                  numels[0] * sizeof(float), S, cudaMemcpyDeviceToDevice, stream));
                else if (int e = launch_philox_fill(ws.eps + offs[0], L.ldw, numels[0], vars[0].var_id, *r, stream)) return e;
                dim3 grid((P + 255) / 256, HP, S);
                sample_w1_split_kernel<<<grid, 256, 0, stream>>>(vars[0].mu, vars[0].rho, ws.eps + offs[0], L.ldw, ws.Wh, ws.Wl, H, P,
                                                                 HP, ws.ldP);
                BRN_LAUNCH_OK("sample_w1_split_kernel");
            }
            if (S % NS) {   // the odd tail tile reads one more sample block: keep it finite
                BRN_CUDA_OK(cudaMemsetAsync(ws.Wh + (size_t)S * HP * ws.ldP, 0, sizeof(float) * HP * ws.ldP, stream));
                BRN_CUDA_OK(cudaMemsetAsync(ws.Wl + (size_t)S * HP * ws.ldP, 0, sizeof(float) * HP * ws.ldP, stream));
            }
            // (also zeroes the per-sample gradient slots of b1 / W2 / b2, which the mid stage accumulates into)
            if (int e = launch_sample_multi(vars + 1, offs + 1, 3, ws.eps, ws.W, L.ldw, *r, stream, ws.dW)) return e;
            // first read of the minibatch: everything above overlaps a host->device copy announced by brn_set_data_ready_event
            if (int e = wait_data_ready(stream)) return e;
            if (int e = launch_split_tf32(X, P, B, P, ws.Xh, ws.Xl, ws.ldP, ws.Xth, ws.Xtl, ws.ldB, stream)) return e;
        } else {
            if (int e = launch_sample_multi(vars, offs, 4, ws.eps, ws.W, L.ldw, *r, stream)) return e;
            if (int e = wait_data_ready(stream)) return e;
        }
    }
    // 2. pre_s = X . W1_s^T
    {
        StageTimer st("bnn.gemm_fwd", stream);
        if (use_tc) {
            EpiStore::Params ep;      // pre^T: [S][H][B], coalesced across the warp's rows b
            ep.out = ws.pre; ep.rows = B; ep.row_stride = 1; ep.col_stride = B; ep.blk_stride = (int64_t)B * H;
            ep.blk_valid = H; ep.col_limit = 0; ep.total_blks = S;
            if (int e = launch_umma_nt<BN, BNN_UMMA_BK, EpiStore>(ws.Xh, ws.Xl, B, ws.ldP, ws.Wh, ws.Wl, S * HP, ws.ldP, P, 0, drain, ep,
                                                     stream))
                return e;
        } else {
            if (int e = launch_sgemm_batched<true, true>(X, P, 0, ws.W + L.oW1, P, L.ldw, ws.pre, H, (int64_t)B * H, B, H, P,
                                                         S, stream))
                return e;
        }
    }
    // 3. mid
    if (!use_tc)
        BRN_CUDA_OK(cudaMemset2DAsync(ws.dW + L.ob1, L.ldw * sizeof(float), 0, (L.numel - L.ob1) * sizeof(float), S, stream));
    {
        StageTimer st("bnn.mid", stream);
        const float inv_S = 1.0f / (float)r->s_total;
        const MidSmem ms(H, C);
        const size_t smem2 = ms.total * sizeof(float);
        bool mid4 = true;        // BRN_BNN_MID=3 selects the scalar-FMA kernel (kept for A/B measurements and tests)
        if (const char* env = getenv("BRN_BNN_MID")) mid4 = atoi(env) != 3;
        if (use_tc && mid4) {
            if (int e = launch_mid4<HP>(ws.pre, ws.W, ws.dW, y, L, S, inv_S, loss, ws.dph, ws.dpl, ws.ldB, stream)) return e;
        } else if (use_tc) {
            int e = 0;
            if (C <= 4) e = launch_mid3<4>(ws.pre, ws.W, ws.dW, y, L, S, inv_S, loss, ws.dph, ws.dpl, HP, ws.ldB, stream);
            else if (C <= 8) e = launch_mid3<8>(ws.pre, ws.W, ws.dW, y, L, S, inv_S, loss, ws.dph, ws.dpl, HP, ws.ldB, stream);
            else if (C <= 12) e = launch_mid3<12>(ws.pre, ws.W, ws.dW, y, L, S, inv_S, loss, ws.dph, ws.dpl, HP, ws.ldB, stream);
            else e = launch_mid3<16>(ws.pre, ws.W, ws.dW, y, L, S, inv_S, loss, ws.dph, ws.dpl, HP, ws.ldB, stream);
            if (e) return e;
        } else if (smem2 <= 200 * 1024) {
            dim3 grid((B + MID_R - 1) / MID_R, S);
            BRN_CUDA_OK(cudaFuncSetAttribute(bnn_mid2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            bnn_mid2_kernel<false><<<grid, 256, smem2, stream>>>(ws.pre, ws.W, ws.dW, y, L, inv_S, loss, nullptr, nullptr, HP,
                                                                 ws.ldB);
            BRN_LAUNCH_OK("bnn_mid2_kernel");
        } else {
            // very wide hidden layers: the one-thread-per-row kernel with fewer rows per CTA (SIMT variant only)
            int R = 128;
            while (R > 32 && bnn_mid_smem(R, H, C) > 200 * 1024) R -= 32;
            size_t smem = bnn_mid_smem(R, H, C);
            BRN_CHECK_ARG(smem <= 220 * 1024, "brn_bnn_elbo_fwd_bwd: hidden width H=%d too large for the mid kernel", H);
            BRN_CUDA_OK(cudaFuncSetAttribute(bnn_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grid((B + R - 1) / R, S);
            bnn_mid_kernel<<<grid, R, smem, stream>>>(ws.pre, ws.W, ws.dW, y, L, R, inv_S, loss, nullptr, nullptr, HP, ws.ldB);
            BRN_LAUNCH_OK("bnn_mid_kernel");
        }
    }
    // 4. dW1_s = dpre_s^T . X   (tcgen05: folded over samples in the GEMM epilogue -> gw, gwe directly)
    {
        StageTimer st("bnn.gemm_bwd", stream);
        if (use_tc) {
            if (S % NS) {
                BRN_CUDA_OK(cudaMemsetAsync(ws.dph + (size_t)S * HP * ws.ldB, 0, sizeof(float) * HP * ws.ldB, stream));
                BRN_CUDA_OK(cudaMemsetAsync(ws.dpl + (size_t)S * HP * ws.ldB, 0, sizeof(float) * HP * ws.ldB, stream));
            }
            // K-split tail (see UnitIter): the dW1 blocks of the samples in the split n-tiles take atomic partial sums
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const UmmaSplitPlan plan = umma_plan<BN, BNN_UMMA_BK>(P, S * HP, B, sms, true);
            if (plan.first_split_ntile >= 0) {
                const int s_first = plan.first_split_ntile * NS;
                if (s_first < S)
                    BRN_CUDA_OK(cudaMemset2DAsync(ws.dW + (size_t)s_first * L.ldw + L.oW1, L.ldw * sizeof(float), 0,
                                                  (size_t)H * P * sizeof(float), S - s_first, stream));
            }
            EpiStore::Params ep;      // dW1_s[h][p] = D[p, (s, h)]
            ep.out = ws.dW + L.oW1; ep.rows = P; ep.row_stride = 1; ep.col_stride = P; ep.blk_stride = L.ldw;
            ep.blk_valid = H; ep.col_limit = 0; ep.total_blks = S;
            if (int e = launch_umma_nt<BN, BNN_UMMA_BK, EpiStore>(ws.Xth, ws.Xtl, P, ws.ldB, ws.dph, ws.dpl, S * HP, ws.ldB, B, 0, drain, ep,
                                                     stream, true))
                return e;
        } else {
            if (int e = launch_sgemm_batched<false, false>(ws.pre, H, (int64_t)B * H, X, P, 0, ws.dW + L.oW1, P, L.ldw, H, P,
                                                           B, S, stream))
                return e;
        }
    }
    // 5. reduce over samples + prior/entropy + chain rule: one stats launch + one finalize launch for all four variables
    StageTimer st5("bnn.reduce_finalize", stream);
    if (int e = launch_mf_reduce_finalize_multi(vars, offs, 4, L.numel, ws.eps, L.ldw, ws.dW, L.ldw, ws.stats, *r, with_prior, loss,
                                                stream, regen_numel0))
        return e;
    return 0;
}
