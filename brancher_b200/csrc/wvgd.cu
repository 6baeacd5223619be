// K6: Wasserstein variational gradient descent over (sampler, particle) ensembles -- see include/brancher_cuda.h.
//   wvgd_sample_assign_kernel  reparameterised draws of every sampler + Voronoi owner of every draw (tiled n x P distance pass)
//   wvgd_reduce_kernel         masked means, importance weights (softmax over the accepted draws), gradients; one CTA per sampler
// The per-vector log-likelihoods and their gradients in between come from the K4a kernel (linear.cu).
#include "common.cuh"

namespace brn {

constexpr int WV_TV = 32;      // sample vectors per CTA
constexpr int WV_TJ = 128;     // particles per tile = threads per CTA
constexpr int WV_EC = 64;      // distance elements per shared-memory chunk

__global__ void __launch_bounds__(WV_TJ)
wvgd_sample_assign_kernel(const float* __restrict__ loc, const float* __restrict__ rho, int rho_per_elem,
                          const float* __restrict__ theta, const float* __restrict__ eps_in, int P, int S, int d, int sel_count,
                          int sel_stride, int draw, brn_sample_range r, float* __restrict__ Z, float* __restrict__ eps_out,
                          int32_t* __restrict__ owner) {
    __shared__ float zs[WV_TV][WV_EC];
    __shared__ float ths[WV_EC][WV_TJ + 1];
    __shared__ float red_d[WV_TJ / 32][WV_TV];
    __shared__ int red_j[WV_TJ / 32][WV_TV];
    const int tid = threadIdx.x;
    const int64_t m = (int64_t)P * S, v0 = (int64_t)blockIdx.x * WV_TV;
    const int nv = (int)min((int64_t)WV_TV, m - v0);

    // ---- phase 0: z = loc + sigma * eps for this CTA's vectors (quads of 4 elements: one Philox call each)
    const int quads = (d + 3) / 4;
    for (int idx = tid; idx < nv * quads; idx += WV_TJ) {
        const int v = idx / quads, q = idx - v * quads;
        const int64_t vec = v0 + v;
        const int k = (int)(vec / S), s = (int)(vec - (int64_t)k * S);
        float e[4];
        if (eps_in) {
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = (4 * q + j < d) ? eps_in[vec * d + 4 * q + j] : 0.f;
        } else {
            Normal4 n4 = philox_normal4(r.seed, philox_offset(r), (uint32_t)(2 * k + draw), (uint32_t)(r.s0 + s), (uint32_t)q);
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = n4.v[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int el = 4 * q + j;
            if (el >= d) break;
            const float sg = softplusf(rho_per_elem ? rho[(int64_t)k * d + el] : rho[k]);
            Z[vec * d + el] = __fmaf_rn(sg, e[j], loc[(int64_t)k * d + el]);
            if (eps_out) eps_out[vec * d + el] = e[j];
        }
    }
    __syncthreads();

    // ---- phase 1: squared distances to every particle, tile by tile; thread = one particle of the tile
    float best_d[WV_TV];
    int best_j[WV_TV];
#pragma unroll
    for (int v = 0; v < WV_TV; ++v) { best_d[v] = INFINITY; best_j[v] = 0x7fffffff; }
    for (int jt = 0; jt < P; jt += WV_TJ) {
        float dist[WV_TV];
#pragma unroll
        for (int v = 0; v < WV_TV; ++v) dist[v] = 0.f;
        for (int c0 = 0; c0 < sel_count; c0 += WV_EC) {
            const int nc = min(WV_EC, sel_count - c0);
            __syncthreads();
            for (int idx = tid; idx < WV_TV * WV_EC; idx += WV_TJ) {
                const int v = idx / WV_EC, c = idx - v * WV_EC;
                zs[v][c] = (v < nv && c < nc) ? Z[(v0 + v) * d + (int64_t)(c0 + c) * sel_stride] : 0.f;
            }
            for (int idx = tid; idx < WV_TJ * WV_EC; idx += WV_TJ) {
                const int jj = idx / WV_EC, c = idx - jj * WV_EC;
                ths[c][jj] = (jt + jj < P && c < nc) ? theta[(int64_t)(jt + jj) * d + (int64_t)(c0 + c) * sel_stride] : 0.f;
            }
            __syncthreads();
            for (int c = 0; c < nc; ++c) {
                const float t = ths[c][tid];
#pragma unroll
                for (int v = 0; v < WV_TV; ++v) {
                    const float df = zs[v][c] - t;
                    dist[v] = __fmaf_rn(df, df, dist[v]);
                }
            }
        }
        const int j = jt + tid;
        if (j < P) {
#pragma unroll
            for (int v = 0; v < WV_TV; ++v)
                if (dist[v] < best_d[v]) { best_d[v] = dist[v]; best_j[v] = j; }
        }
    }
    // ---- phase 2: argmin over the CTA's threads (smallest distance, then smallest index: np.argmin)
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int v = 0; v < WV_TV; ++v) {
        float bd = best_d[v];
        int bj = best_j[v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
        }
        if (lane == 0) { red_d[warp][v] = bd; red_j[warp][v] = bj; }
    }
    __syncthreads();
    if (tid < nv) {
        float bd = red_d[0][tid];
        int bj = red_j[0][tid];
#pragma unroll
        for (int w = 1; w < WV_TJ / 32; ++w) {
            const float od = red_d[w][tid];
            const int oj = red_j[w][tid];
            if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
        }
        owner[v0 + tid] = bj;
    }
}

constexpr int WR_THREADS = 256;

__device__ __forceinline__ double block_sum_all(double v, double* scratch, double* bcast) {
    double t = block_sum<double>(v, scratch);
    if (threadIdx.x == 0) *bcast = t;
    __syncthreads();
    t = *bcast;
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(WR_THREADS) wvgd_reduce_kernel(brn_wvgd_args a) {
    extern __shared__ double wsm[];           // [S] log-weights, then weights
    __shared__ double red[32];
    __shared__ double bc;
    const int k = blockIdx.x, tid = threadIdx.x, S = a.S, d = a.d;
    const int lane = tid & 31, warp = tid >> 5, nwarp = WR_THREADS / 32;
    const bool tied = a.prior_loc == nullptr;
    const int32_t* o0 = a.owner0 + (int64_t)k * S;
    const int32_t* o1 = a.owner1 + (int64_t)k * S;
    const int64_t base = (int64_t)k * S * d;
    auto sigma_of = [&](int e) { return softplusf(a.rho_per_elem ? a.rho[(int64_t)k * d + e] : a.rho[k]); };
    auto rho_of = [&](int e) { return a.rho_per_elem ? a.rho[(int64_t)k * d + e] : a.rho[k]; };

    double c0 = 0.0, c1 = 0.0;
    for (int s = tid; s < S; s += WR_THREADS) { c0 += o0[s] == k; c1 += o1[s] == k; }
    const int n0 = (int)block_sum_all(c0, red, &bc), n1 = (int)block_sum_all(c1, red, &bc);
    if (tid == 0) { a.counts[2 * k] = n0; a.counts[2 * k + 1] = n1; }

    // ---- sampler ELBO: value -mean_A[ll + log prior(z)], gradients to (loc, rho)
    double val = 0.0;
    if (n0 > 0) {
        for (int s = tid; s < S; s += WR_THREADS)
            if (o0[s] == k) val += a.ll0[(int64_t)k * S + s];
        for (int64_t idx = tid; idx < (int64_t)S * d; idx += WR_THREADS) {
            const int s = (int)(idx / d), e = (int)(idx - (int64_t)s * d);
            if (o0[s] != k) continue;
            if (tied) {
                const float ep = a.eps0[base + idx];
                val += (double)(-0.5f * ep * ep - logf(sigma_of(e)) - BRN_HALF_LOG_2PI);
            } else {
                const float pa = a.prior_loc[e], pb = a.prior_scale[e], df = a.Z0[base + idx] - pa;
                val += (double)(-(df * df) / (2.f * (pb * pb)) - logf(pb) - BRN_HALF_LOG_2PI);
            }
        }
    }
    const float inv_n0 = n0 > 0 ? 1.0f / (float)n0 : 0.f;
    for (int e = tid; e < d; e += WR_THREADS) {
        float gm = 0.f, gs = 0.f;
        if (n0 > 0) {
            for (int s = 0; s < S; ++s) {
                if (o0[s] != k) continue;
                const int64_t i = base + (int64_t)s * d + e;
                float g = -a.G0[i];                               // d ll / d z
                if (!tied) {
                    const float pb = a.prior_scale[e];
                    g -= (a.Z0[i] - a.prior_loc[e]) / (pb * pb);  // + d log prior / d z
                }
                gm += g;
                gs = __fmaf_rn(g, a.eps0[i], gs);
            }
        }
        const float sg = sigma_of(e);
        float dsig = -(gs * inv_n0);
        if (!tied && n0 > 0) dsig -= 1.0f / sg;                   // -log q(z(phi); phi) = ... + log sigma
        a.dloc[(int64_t)k * d + e] = -(gm * inv_n0);
        a.drho[(int64_t)k * d + e] = dsig * sigmoidf(rho_of(e));
    }
    double loss_k = n0 > 0 ? -block_sum_all(val, red, &bc) / (double)n0 : 0.0;

    // ---- particle loss: importance weights over the accepted draws of the second noise draw
    for (int s = warp; s < S; s += nwarp) {
        double lw = -INFINITY;
        if (o1[s] == k) {
            float corr = 0.f;
            if (!tied && !a.biased) {
                for (int e = lane; e < d; e += 32) {
                    const int64_t i = base + (int64_t)s * d + e;
                    const float pa = a.prior_loc[e], pb = a.prior_scale[e], df = a.Z1[i] - pa, ep = a.eps1[i];
                    corr += (-(df * df) / (2.f * (pb * pb)) - logf(pb)) - (-0.5f * ep * ep - logf(sigma_of(e)));
                }
                corr = warp_sum(corr);
            }
            lw = a.biased ? 0.0 : a.ll1[(int64_t)k * S + s] + (double)corr;
        }
        if (lane == 0) wsm[s] = lw;
    }
    __syncthreads();
    double mx = -INFINITY;
    for (int s = tid; s < S; s += WR_THREADS) mx = fmax(mx, wsm[s]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < nwarp; ++w) mx = fmax(mx, red[w]);
    __syncthreads();
    double zsum = 0.0;
    for (int s = tid; s < S; s += WR_THREADS) {
        const double wv = (o1[s] == k) ? exp(wsm[s] - mx) : 0.0;
        wsm[s] = wv;
        zsum += wv;
    }
    zsum = block_sum_all(zsum, red, &bc);
    const double wnorm = a.biased ? 1.0 / (double)S : (n1 > 0 ? 1.0 / zsum : 0.0);
    double pl = 0.0;
    for (int e = tid; e < d; e += WR_THREADS) {
        const float th = a.theta[(int64_t)k * d + e];
        float g = 0.f, q = 0.f;
        if (n1 > 0) {
            for (int s = 0; s < S; ++s) {
                if (o1[s] != k) continue;
                const float w = (float)(wsm[s] * wnorm);
                const float df = th - a.Z1[base + (int64_t)s * d + e];
                g = __fmaf_rn(w, df, g);
                q = __fmaf_rn(w * df, df, q);
            }
        }
        a.dtheta[(int64_t)k * d + e] = 2.f * g;
        pl += (double)q;
    }
    pl = block_sum_all(pl, red, &bc);
    if (tid == 0) atomicAdd(a.loss, loss_k + pl);
}

}  // namespace brn

using namespace brn;

extern "C" int brn_wvgd_sample_assign(const float* loc, const float* rho, int rho_per_elem, const float* theta, const float* eps,
                                      int P, int S, int d, int F_last, int first_column_only, int draw, const brn_sample_range* r,
                                      float* Z, float* eps_out, int32_t* owner, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BRN_CHECK_ARG(loc && rho && theta && r && Z && owner, "brn_wvgd_sample_assign: NULL pointer");
    BRN_CHECK_ARG(P > 0 && S > 0 && d > 0, "brn_wvgd_sample_assign: bad shape P=%d S=%d d=%d", P, S, d);
    BRN_CHECK_ARG(F_last > 0 && d % F_last == 0, "brn_wvgd_sample_assign: d=%d is not a multiple of the last axis F=%d", d, F_last);
    BRN_CHECK_ARG(draw == 0 || draw == 1, "brn_wvgd_sample_assign: draw must be 0 or 1 (got %d)", draw);
    BRN_CHECK_ARG(eps || eps_out, "brn_wvgd_sample_assign: Philox mode needs eps_out (the reduction reads the noise)");
    BRN_CHECK_ARG(r->s_local == S && r->s0 >= 0, "brn_wvgd_sample_assign: the sample range must cover the S=%d draws", S);
    set_variant("simt");
    StageTimer st("wvgd.sample_assign", stream);
    const int sel_count = first_column_only ? d / F_last : d, sel_stride = first_column_only ? F_last : 1;
    const int64_t m = (int64_t)P * S;
    wvgd_sample_assign_kernel<<<(unsigned)((m + WV_TV - 1) / WV_TV), WV_TJ, 0, stream>>>(loc, rho, rho_per_elem, theta, eps, P, S, d,
                                                                                       sel_count, sel_stride, draw, *r, Z, eps_out,
                                                                                       owner);
    BRN_LAUNCH_OK("wvgd_sample_assign_kernel");
    return 0;
}

extern "C" int brn_wvgd_reduce(const brn_wvgd_args* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BRN_CHECK_ARG(a, "brn_wvgd_reduce: NULL argument block");
    BRN_CHECK_ARG(a->P > 0 && a->S > 0 && a->d > 0, "brn_wvgd_reduce: bad shape P=%d S=%d d=%d", a->P, a->S, a->d);
    BRN_CHECK_ARG(a->loc && a->rho && a->theta && a->owner0 && a->owner1 && a->ll0 && a->ll1 && a->G0 && a->eps0 && a->eps1 &&
                  a->Z0 && a->Z1 && a->dloc && a->drho && a->dtheta && a->counts && a->loss, "brn_wvgd_reduce: NULL pointer");
    BRN_CHECK_ARG((a->prior_loc == nullptr) == (a->prior_scale == nullptr), "prior_loc and prior_scale must both be given or both NULL");
    const size_t smem = sizeof(double) * (size_t)a->S;
    BRN_CHECK_ARG(smem <= 200 * 1024, "brn_wvgd_reduce: S=%d draws per sampler exceed the shared-memory weight buffer", a->S);
    set_variant("simt");
    StageTimer st("wvgd.reduce", stream);
    if (smem > 48 * 1024)
        BRN_CUDA_OK(cudaFuncSetAttribute(wvgd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wvgd_reduce_kernel<<<a->P, WR_THREADS, smem, stream>>>(*a);
    BRN_LAUNCH_OK("wvgd_reduce_kernel");
    return 0;
}
