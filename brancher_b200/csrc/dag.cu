// K1: generic scalar-DAG ELBO -- fused reparameterised sampling, log-probabilities, two-axis reduction and the
// matching backward in ONE launch.  See include/brancher_cuda.h (brn_dag_elbo_fwd_bwd).
//
// The host (brancher_b200/lowering.py: DagPlan) flattens a (joint, posterior) pair whose variables are scalars
// (README AR(1), examples/logNormal_normal.py, examples/multivariate_regression.py ...) into a straight-line SSA
// program of `brn_dag_op`.  One thread evaluates the whole program for one (MC sample s, data row b) pair:
//   forward   values v[slot] in thread-local memory (every thread executes the same op sequence: no divergence)
//   reverse   adjoints in a second local array, walking the program backwards (reverse-mode AD by hand)
//   reduce    d loss / d param: shared-memory atomics per CTA, then one global atomic per parameter per CTA;
//             loss: warp-shuffle + shared reduction in fp64, one atomic per CTA.
// ACC_SAMPLE terms (latent log-probs, entropies) are counted once per sample (row 0 only), ACC_ROW terms (observed
// nodes: summed over the data axis, variables.py:513-514) for every row; loss = -(1/S_total) sum.
#include "common.cuh"

namespace brn {

enum DagOp : int32_t {
    DAG_CONST = BRN_DAG_CONST, DAG_PARAM = BRN_DAG_PARAM, DAG_DATA = BRN_DAG_DATA, DAG_EPS = BRN_DAG_EPS,
    DAG_ADD = BRN_DAG_ADD, DAG_SUB = BRN_DAG_SUB, DAG_MUL = BRN_DAG_MUL, DAG_DIV = BRN_DAG_DIV, DAG_NEG = BRN_DAG_NEG,
    DAG_POWI = BRN_DAG_POWI, DAG_EXP = BRN_DAG_EXP, DAG_LOG = BRN_DAG_LOG, DAG_LOG1P = BRN_DAG_LOG1P,
    DAG_SIGMOID = BRN_DAG_SIGMOID, DAG_SOFTPLUS = BRN_DAG_SOFTPLUS, DAG_TANH = BRN_DAG_TANH, DAG_SIN = BRN_DAG_SIN,
    DAG_COS = BRN_DAG_COS, DAG_RELU = BRN_DAG_RELU, DAG_SQRT = BRN_DAG_SQRT, DAG_ABS = BRN_DAG_ABS,
    DAG_CLAMP_UNIT = BRN_DAG_CLAMP_UNIT, DAG_NORMAL_LP = BRN_DAG_NORMAL_LP, DAG_NORMAL_ENTROPY = BRN_DAG_NORMAL_ENTROPY,
    DAG_ACC_SAMPLE = BRN_DAG_ACC_SAMPLE, DAG_ACC_ROW = BRN_DAG_ACC_ROW
};

constexpr int DAG_MAX_PARAMS = BRN_DAG_MAX_PARAMS;     // shared-memory gradient accumulators

__device__ __forceinline__ float dag_powi(float x, float p) { return powf(x, p); }

template <int MAXS>      // value / adjoint slots per thread (thread-local memory)
__global__ void __launch_bounds__(128)
dag_elbo_kernel(const brn_dag_op* __restrict__ ops, int n_ops, int n_slots, const float* __restrict__ params, int n_params,
                const float* __restrict__ data, int n_cols, int B, const float* __restrict__ eps, int n_eps,
                brn_sample_range r, float* __restrict__ dparams, double* __restrict__ loss) {
    extern __shared__ float sgrad[];          // [n_params]
    __shared__ double red[32];
    float v[MAXS];
    float adj[MAXS];
    for (int i = threadIdx.x; i < n_params; i += blockDim.x) sgrad[i] = 0.f;
    __syncthreads();

    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)r.s_local * B;
    const bool active = gid < total;
    const int s = active ? (int)(gid / B) : 0, b = active ? (int)(gid - (int64_t)s * B) : 0;
    const float inv_S = 1.0f / (float)r.s_total;
    float acc = 0.f;

    if (active) {
        // ---------------- forward
        for (int i = 0; i < n_ops; ++i) {
            const brn_dag_op o = ops[i];
            float x = 0.f;
            switch (o.opcode) {
                case DAG_CONST: x = o.imm; break;
                case DAG_PARAM: x = params[o.a]; break;
                case DAG_DATA: x = data[(int64_t)b * n_cols + o.a]; break;
                case DAG_EPS:
                    x = eps ? eps[(int64_t)s * n_eps + o.a]
                            : philox_normal1(r.seed, r.offset, (uint32_t)o.a, (uint32_t)(r.s0 + s), 0);
                    break;
                case DAG_ADD: x = v[o.a] + v[o.b]; break;
                case DAG_SUB: x = v[o.a] - v[o.b]; break;
                case DAG_MUL: x = v[o.a] * v[o.b]; break;
                case DAG_DIV: x = v[o.a] / v[o.b]; break;
                case DAG_NEG: x = -v[o.a]; break;
                case DAG_POWI: x = dag_powi(v[o.a], o.imm); break;
                case DAG_EXP: x = expf(v[o.a]); break;
                case DAG_LOG: x = logf(v[o.a]); break;
                case DAG_LOG1P: x = log1pf(v[o.a]); break;
                case DAG_SIGMOID: x = sigmoidf(v[o.a]); break;
                case DAG_SOFTPLUS: x = softplusf(v[o.a]); break;
                case DAG_TANH: x = tanhf(v[o.a]); break;
                case DAG_SIN: x = sinf(v[o.a]); break;
                case DAG_COS: x = cosf(v[o.a]); break;
                case DAG_RELU: x = fmaxf(v[o.a], 0.f); break;
                case DAG_SQRT: x = sqrtf(v[o.a]); break;
                case DAG_ABS: x = fabsf(v[o.a]); break;
                case DAG_CLAMP_UNIT: x = fminf(fmaxf(v[o.a], 1.17549435e-38f), 1.0f - 1.1920929e-07f); break;
                case DAG_NORMAL_LP: {      // torch Normal.log_prob: -((x-mu)^2)/(2 sigma^2) - log sigma - log sqrt(2 pi)
                    const float df = v[o.a] - v[o.b], sg = v[o.c];
                    x = -(df * df) / (2.f * (sg * sg)) - logf(sg) - BRN_HALF_LOG_2PI;
                    break;
                }
                case DAG_NORMAL_ENTROPY: x = 0.5f + BRN_HALF_LOG_2PI + logf(v[o.a]); break;
                case DAG_ACC_SAMPLE: if (b == 0) acc += v[o.a]; break;
                case DAG_ACC_ROW: acc += v[o.a]; break;
                default: break;
            }
            v[o.dst] = x;
        }
        // ---------------- reverse: d loss / d slot, loss = -(1/S) * acc
        for (int i = 0; i < n_slots; ++i) adj[i] = 0.f;
        for (int i = n_ops - 1; i >= 0; --i) {
            const brn_dag_op o = ops[i];
            const float g = adj[o.dst];
            switch (o.opcode) {
                case DAG_ACC_SAMPLE: if (b == 0) adj[o.a] -= inv_S; break;
                case DAG_ACC_ROW: adj[o.a] -= inv_S; break;
                case DAG_PARAM: if (g != 0.f) atomicAdd(&sgrad[o.a], g); break;
                case DAG_ADD: adj[o.a] += g; adj[o.b] += g; break;
                case DAG_SUB: adj[o.a] += g; adj[o.b] -= g; break;
                case DAG_MUL: { const float xa = v[o.a], xb = v[o.b]; adj[o.a] += g * xb; adj[o.b] += g * xa; break; }
                case DAG_DIV: {
                    const float inv = 1.f / v[o.b];
                    adj[o.a] += g * inv;
                    adj[o.b] -= g * v[o.dst] * inv;
                    break;
                }
                case DAG_NEG: adj[o.a] -= g; break;
                case DAG_POWI: adj[o.a] += g * o.imm * dag_powi(v[o.a], o.imm - 1.f); break;
                case DAG_EXP: adj[o.a] += g * v[o.dst]; break;
                case DAG_LOG: adj[o.a] += g / v[o.a]; break;
                case DAG_LOG1P: adj[o.a] += g / (1.f + v[o.a]); break;
                case DAG_SIGMOID: { const float y = v[o.dst]; adj[o.a] += g * y * (1.f - y); break; }
                case DAG_SOFTPLUS: adj[o.a] += g * (v[o.a] > 20.f ? 1.f : sigmoidf(v[o.a])); break;
                case DAG_TANH: { const float y = v[o.dst]; adj[o.a] += g * (1.f - y * y); break; }
                case DAG_SIN: adj[o.a] += g * cosf(v[o.a]); break;
                case DAG_COS: adj[o.a] -= g * sinf(v[o.a]); break;
                case DAG_RELU: adj[o.a] += v[o.a] > 0.f ? g : 0.f; break;
                case DAG_SQRT: adj[o.a] += g * 0.5f / v[o.dst]; break;
                case DAG_ABS: adj[o.a] += v[o.a] >= 0.f ? g : -g; break;
                case DAG_CLAMP_UNIT: {
                    const float xa = v[o.a];
                    adj[o.a] += (xa >= 1.17549435e-38f && xa <= 1.0f - 1.1920929e-07f) ? g : 0.f;
                    break;
                }
                case DAG_NORMAL_LP: {
                    const float df = v[o.a] - v[o.b], sg = v[o.c], inv_var = 1.f / (sg * sg);
                    const float t = g * df * inv_var;
                    adj[o.a] -= t;
                    adj[o.b] += t;
                    adj[o.c] += g * (df * df * inv_var - 1.f) / sg;
                    break;
                }
                case DAG_NORMAL_ENTROPY: adj[o.a] += g / v[o.a]; break;
                default: break;     // CONST / DATA / EPS: leaves
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_params; i += blockDim.x) {
        const float gsum = sgrad[i];
        if (gsum != 0.f) atomicAdd(&dparams[i], gsum);
    }
    double tot = block_sum<double>((double)acc, red);
    if (threadIdx.x == 0) atomicAdd(loss, -tot * (double)inv_S);
}

}  // namespace brn

using namespace brn;

extern "C" int brn_dag_elbo_fwd_bwd(const brn_dag_op* ops, int n_ops, int n_slots, const float* params, int n_params,
                                    const float* data, int n_cols, int n_rows, const float* eps, int n_eps,
                                    const brn_sample_range* r, float* dparams, double* loss, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BRN_CHECK_ARG(ops && r && loss, "brn_dag_elbo_fwd_bwd: NULL pointer");
    BRN_CHECK_ARG(n_ops > 0 && n_slots > 0 && n_slots <= BRN_DAG_MAX_SLOTS, "brn_dag_elbo_fwd_bwd: n_slots=%d outside (0, %d]",
                  n_slots, BRN_DAG_MAX_SLOTS);
    BRN_CHECK_ARG(n_params >= 0 && n_params <= DAG_MAX_PARAMS, "brn_dag_elbo_fwd_bwd: n_params=%d exceeds %d", n_params, DAG_MAX_PARAMS);
    BRN_CHECK_ARG(n_params == 0 || (params && dparams), "brn_dag_elbo_fwd_bwd: NULL parameter pointer");
    BRN_CHECK_ARG(n_rows >= 1 && n_cols >= 0 && (n_cols == 0 || data), "brn_dag_elbo_fwd_bwd: bad data block rows=%d cols=%d", n_rows,
                  n_cols);
    BRN_CHECK_ARG(r->s_local >= 0 && r->s_total > 0 && r->s0 >= 0 && r->s0 + r->s_local <= r->s_total,
                  "bad sample range s0=%d s_local=%d s_total=%d", r->s0, r->s_local, r->s_total);
    if (r->s_local == 0) return 0;
    set_variant("simt");
    StageTimer st("dag.fused", stream);
    const int64_t total = (int64_t)r->s_local * n_rows;
    // The value/adjoint arrays live in thread-local memory (2 * n_slots lines of 128 B per warp).  Small problems (the
    // README AR(1): 300 samples) run ONE WARP PER CTA so that (a) the work spreads over as many SMs as there are warps and
    // (b) each SM's L1 holds its warp's whole local frame -- with 128-thread CTAs the frame of ~700 slots spilled to L2 and
    // every interpreted op paid an L2 round trip (563 us per evaluation, measured; profiles/r1e_bench_ar1.json).
    const int block = total <= (int64_t)32 * 148 * 8 ? 32 : 128;
    const unsigned grid = (unsigned)((total + block - 1) / block);
    const size_t smem = sizeof(float) * (size_t)(n_params > 0 ? n_params : 1);
#define BRN_DAG_LAUNCH(MAXS)                                                                                             \
    dag_elbo_kernel<MAXS><<<grid, block, smem, stream>>>(ops, n_ops, n_slots, params, n_params, data, n_cols, n_rows, eps, \
                                                       n_eps, *r, dparams, loss)
    if (n_slots <= 128) BRN_DAG_LAUNCH(128);
    else if (n_slots <= 512) BRN_DAG_LAUNCH(512);
    else BRN_DAG_LAUNCH(BRN_DAG_MAX_SLOTS);
#undef BRN_DAG_LAUNCH
    BRN_LAUNCH_OK("dag_elbo_kernel");
    return 0;
}
