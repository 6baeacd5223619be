// Mean-field Normal helpers shared by the likelihood families (K2/K3): noise materialisation,
// weight sampling and the K1a prior+entropy / gradient-finalisation stage.
#pragma once
#include "common.cuh"

namespace brn {

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
