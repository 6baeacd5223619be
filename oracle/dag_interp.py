"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU (numpy, fp64 by default) interpreter of the scalar-DAG programs `brancher_b200.lowering.DagPlan` emits for
K1 (`brn_dag_elbo_fwd_bwd`, include/brancher_cuda.h).  It restates, op by op, the arithmetic the reference delegates
to torch for these model classes -- Normal.rsample `loc + eps*scale`, Normal.log_prob
`-((x-mu)^2)/(2 sigma^2) - log sigma - log sqrt(2 pi)`, Normal.entropy `1/2 + 1/2 log 2 pi + log sigma`,
softplus with threshold 20, SigmoidTransform's clamp (reference call sites distributions.py:108,122,166,180) -- and
the reductions of the reference's graph walk: observed nodes summed over the data axis (variables.py:513-514),
`.mean()` over the sample axis (gradient_estimators.py:44), loss = -ELBO (inference.py:140-144).
Gradients: reverse-mode over the same program.  Pinned by tests/golden/{ar1_readme,lognormal_normal,
multivariate_regression}.npz (outputs of the live reference) in tests/test_dag_program.py.
"""
import numpy as np

OPS = ["CONST", "PARAM", "DATA", "EPS", "ADD", "SUB", "MUL", "DIV", "NEG", "POWI", "EXP", "LOG", "LOG1P", "SIGMOID", "SOFTPLUS",
       "TANH", "SIN", "COS", "RELU", "SQRT", "ABS", "CLAMP_UNIT", "NORMAL_LP", "NORMAL_ENTROPY", "ACC_SAMPLE", "ACC_ROW"]
HALF_LOG_2PI = 0.5 * np.log(2 * np.pi)
TINY32, EPS32 = float(np.finfo(np.float32).tiny), float(np.finfo(np.float32).eps)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def run(ops, n_slots, params, data, eps, s_total=None, dtype=np.float64):
    """ops: iterable of (opcode, dst, a, b, c, imm); params [n_params]; data [B, n_cols] or None; eps [S, n_eps].
    Returns (loss, dparams)."""
    params = np.asarray(params, dtype=dtype)
    eps = np.asarray(eps, dtype=dtype)
    S = eps.shape[0]
    data = np.zeros((1, 0), dtype) if data is None else np.asarray(data, dtype=dtype)
    B = data.shape[0]
    s_total = S if s_total is None else s_total
    v = [None] * n_slots
    full = lambda x: np.broadcast_to(np.asarray(x, dtype=dtype), (S, B))
    acc = np.zeros((S, B), dtype)
    row0 = np.zeros((S, B), dtype)
    row0[:, 0] = 1
    for code, dst, a, b, c, imm in ops:
        op = OPS[code]
        if op == "CONST": x = full(imm)
        elif op == "PARAM": x = full(params[a])
        elif op == "DATA": x = full(data[None, :, a])
        elif op == "EPS": x = full(eps[:, a][:, None])
        elif op == "ADD": x = v[a] + v[b]
        elif op == "SUB": x = v[a] - v[b]
        elif op == "MUL": x = v[a] * v[b]
        elif op == "DIV": x = v[a] / v[b]
        elif op == "NEG": x = -v[a]
        elif op == "POWI": x = v[a] ** imm
        elif op == "EXP": x = np.exp(v[a])
        elif op == "LOG": x = np.log(v[a])
        elif op == "LOG1P": x = np.log1p(v[a])
        elif op == "SIGMOID": x = _sigmoid(v[a])
        elif op == "SOFTPLUS": x = np.where(v[a] > 20, v[a], np.log1p(np.exp(np.minimum(v[a], 20))))
        elif op == "TANH": x = np.tanh(v[a])
        elif op == "SIN": x = np.sin(v[a])
        elif op == "COS": x = np.cos(v[a])
        elif op == "RELU": x = np.maximum(v[a], 0)
        elif op == "SQRT": x = np.sqrt(v[a])
        elif op == "ABS": x = np.abs(v[a])
        elif op == "CLAMP_UNIT": x = np.clip(v[a], TINY32, 1.0 - EPS32)
        elif op == "NORMAL_LP": x = -((v[a] - v[b]) ** 2) / (2 * v[c] ** 2) - np.log(v[c]) - HALF_LOG_2PI
        elif op == "NORMAL_ENTROPY": x = 0.5 + HALF_LOG_2PI + np.log(v[a])
        elif op == "ACC_SAMPLE": acc = acc + v[a] * row0; x = full(0.0)
        elif op == "ACC_ROW": acc = acc + v[a]; x = full(0.0)
        else: raise ValueError(op)
        v[dst] = np.asarray(x, dtype=dtype)
    loss = -acc.sum() / s_total
    adj = [np.zeros((S, B), dtype) for _ in range(n_slots)]
    dparams = np.zeros_like(params)
    for code, dst, a, b, c, imm in reversed(list(ops)):
        op, g = OPS[code], adj[dst]
        if op == "ACC_SAMPLE": adj[a] = adj[a] - row0 / s_total
        elif op == "ACC_ROW": adj[a] = adj[a] - 1.0 / s_total
        elif op == "PARAM": dparams[a] += g.sum()
        elif op == "ADD": adj[a] = adj[a] + g; adj[b] = adj[b] + g
        elif op == "SUB": adj[a] = adj[a] + g; adj[b] = adj[b] - g
        elif op == "MUL": adj[a] = adj[a] + g * v[b]; adj[b] = adj[b] + g * v[a]
        elif op == "DIV": adj[a] = adj[a] + g / v[b]; adj[b] = adj[b] - g * v[dst] / v[b]
        elif op == "NEG": adj[a] = adj[a] - g
        elif op == "POWI": adj[a] = adj[a] + g * imm * v[a] ** (imm - 1)
        elif op == "EXP": adj[a] = adj[a] + g * v[dst]
        elif op == "LOG": adj[a] = adj[a] + g / v[a]
        elif op == "LOG1P": adj[a] = adj[a] + g / (1 + v[a])
        elif op == "SIGMOID": adj[a] = adj[a] + g * v[dst] * (1 - v[dst])
        elif op == "SOFTPLUS": adj[a] = adj[a] + g * np.where(v[a] > 20, 1.0, _sigmoid(v[a]))
        elif op == "TANH": adj[a] = adj[a] + g * (1 - v[dst] ** 2)
        elif op == "SIN": adj[a] = adj[a] + g * np.cos(v[a])
        elif op == "COS": adj[a] = adj[a] - g * np.sin(v[a])
        elif op == "RELU": adj[a] = adj[a] + g * (v[a] > 0)
        elif op == "SQRT": adj[a] = adj[a] + g * 0.5 / v[dst]
        elif op == "ABS": adj[a] = adj[a] + g * np.where(v[a] >= 0, 1.0, -1.0)
        elif op == "CLAMP_UNIT": adj[a] = adj[a] + g * ((v[a] >= TINY32) & (v[a] <= 1.0 - EPS32))
        elif op == "NORMAL_LP":
            df, iv = v[a] - v[b], 1.0 / v[c] ** 2
            adj[a] = adj[a] - g * df * iv
            adj[b] = adj[b] + g * df * iv
            adj[c] = adj[c] + g * (df * df * iv - 1.0) / v[c]
        elif op == "NORMAL_ENTROPY": adj[a] = adj[a] + g / v[a]
    return float(loss), dparams
