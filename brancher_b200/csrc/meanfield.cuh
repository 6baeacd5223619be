// Mean-field Normal helpers shared by the likelihood families (K2/K3): noise materialisation,
// weight sampling and the K1a prior+entropy / gradient-finalisation stage.
#pragma once
#include "common.cuh"

namespace brn {

// closed-form prior / entropy terms and the chain rule to (mu, rho) for one element, given its sample-axis statistics;
// returns the element's ELBO contribution (prior + entropy), accumulates -d ELBO / d param into the gradient sinks
__device__ __forceinline__ double mf_finalize_element(const brn_mf_var& v, int64_t i, float gw, float gwe, float e1, float e2,
                                                      const brn_sample_range& r, int with_prior) {
    const float mu = v.mu[i], rho = v.rho[i], sg = softplusf(rho);
    const float inv_S = 1.0f / (float)r.s_total, n = (float)r.s_local, frac = n * inv_S;
    float dE_dmu = gw * inv_S, dE_dsg = gwe * inv_S;
    double elbo = 0.0;
    if (with_prior) {
        const float log_sg = logf(sg);
        const float entropy = 0.5f + BRN_HALF_LOG_2PI + log_sg;
        if (v.tied) {
            elbo = (double)(-0.5f * e2 * inv_S) + (double)(frac * (entropy - log_sg - BRN_HALF_LOG_2PI));
        } else {
            const float a = v.prior_loc[i], b = v.prior_scale[i], inv_b2 = 1.0f / (b * b);
            const float c0 = mu - a;
            const float sd2 = n * c0 * c0 + 2.f * c0 * sg * e1 + sg * sg * e2;
            const float sd = n * c0 + sg * e1;
            const float sde = c0 * e1 + sg * e2;
            elbo = (double)(-0.5f * sd2 * inv_b2 * inv_S) + (double)(frac * (entropy - logf(b) - BRN_HALF_LOG_2PI));
            dE_dmu += -sd * inv_b2 * inv_S;
            dE_dsg += -sde * inv_b2 * inv_S + frac / sg;
        }
    }
    v.dmu[i] += -dE_dmu;
    v.drho[i] += -dE_dsg * sigmoidf(rho);
    return elbo;
}

// out[(s - s0) * ld + i] = N(0,1) for (var_id, global sample s, element i)
int launch_philox_fill(float* out, int64_t ld, int64_t numel, uint32_t var_id, const brn_sample_range& r,
                       cudaStream_t stream);

// W[s*ldw + i] = mu[i] + softplus(rho[i]) * eps[s*lde + i]
int launch_sample_weights(const float* mu, const float* rho, const float* eps, int64_t lde, float* W, int64_t ldw,
                          int64_t numel, int s_local, cudaStream_t stream);

// K1a (see brn_mf_normal_prior_entropy).  eps/lde override var.eps when non-NULL (workspace noise).
int launch_mf_finalize(const brn_mf_var& var, const float* eps, int64_t lde, const float* gw, const float* gwe,
                       const brn_sample_range& r, int with_prior, double* loss, cudaStream_t stream);

// gw[i] = sum_s dW[s*ld+i], gwe[i] = sum_s dW[s*ld+i]*eps[s*lde+i]
int launch_reduce_over_samples(const float* dW, int64_t ld, const float* eps, int64_t lde, float* gw, float* gwe,
                               int64_t numel, int s_local, cudaStream_t stream);

// Stage 5 of the likelihood families, HBM-bound and parallel over (elements x sample chunks):
//   stats kernel : gw = sum_s dW_s, gwe = sum_s dW_s*eps_s, e1 = sum_s eps_s, e2 = sum_s eps_s^2   (atomics into `stats`)
//   finalize     : prior log-prob / entropy in closed form from (e1, e2), chain rule through softplus, loss.
// `stats` is a caller-provided scratch of 4*pad4(numel) floats (zeroed here).  dW may be NULL (prior/entropy only).
int launch_mf_reduce_finalize(const brn_mf_var& var, const float* eps, int64_t lde, const float* dW, int64_t ldd,
                              float* stats, const brn_sample_range& r, int with_prior, double* loss, cudaStream_t stream);

// The same for up to 4 variables stored back to back inside one [S][ld] block (`offs` = element offset of each
// variable inside a row, `total` = end of the last one): one stats launch + one finalize launch.
int launch_mf_reduce_finalize_multi(const brn_mf_var* vars, const int64_t* offs, int nvars, int64_t total, const float* eps,
                                    int64_t lde, const float* dW, int64_t ldd, float* stats, const brn_sample_range& r,
                                    int with_prior, double* loss, cudaStream_t stream, int64_t philox_numel0 = 0);

// eps (injected var.eps or Philox) -> eps_out[s*ld + offs[k] + i], W[...] = mu + softplus(rho)*eps for up to 4 variables
int launch_sample_multi(const brn_mf_var* vars, const int64_t* offs, int nvars, float* eps_out, float* W, int64_t ld,
                        const brn_sample_range& r, cudaStream_t stream, float* grad_zero = nullptr);

}  // namespace brn
