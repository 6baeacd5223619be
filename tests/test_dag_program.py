"""K1 scalar-DAG family.
CPU: the host compile (brancher_b200.lowering.DagPlan) evaluated by the numpy interpreter (oracle/dag_interp.py) must
reproduce the live reference's loss and gradients (tests/golden/*.npz) -- this pins BOTH the compile and the interpreter.
GPU: the same programs through brn_dag_elbo_fwd_bwd vs the goldens and vs the fp64 interpreter."""
import numpy as np
import pytest
import torch

import model_zoo as zoo
from helpers import load_golden, assert_close

CASES = {
    "ar1_readme": (zoo.ar1, dict(seed=6, T=20)),
    "lognormal_normal": (zoo.lognormal_normal, dict(seed=10, N=20)),
    "multivariate_regression": (zoo.multivariate_regression, dict(seed=11, n=50)),
    "robust_regression": (zoo.robust_regression, dict(seed=14, n=40)),        # Laplace priors, Cauchy likelihood
    "scalar_logistic": (zoo.scalar_logistic, dict(seed=15, n=30)),            # observed Binomial(1, logits) node
    "op_zoo": (zoo.op_zoo, dict(seed=16, n=24)),                              # every unary / binary op of the family
}


def build(tag, device):
    from brancher_b200 import config, lowering
    config.set_device(device)
    ns = zoo.namespace("brancher_b200")
    builder, kw = CASES[tag]
    model, Q, d = builder(ns, **kw)
    plan = lowering.get_plan(model, model.posterior_model)
    names = {id(v._value): v.name for v in model.posterior_model.flatten()
             if getattr(v, "learnable", False) and isinstance(getattr(v, "_value", None), torch.nn.Parameter)}
    return ns, model, plan, names


def interp(plan, names, g, dtype=np.float64):
    from oracle import dag_interp
    P = plan.prog
    eps = np.stack([g["eps"][n] for n in P.eps_names], 1)
    pv = np.array([g["param"][names[id(p)]].reshape(()) for p in P.params])
    from brancher_b200.lowering import _observed_tensor
    cols = [np.broadcast_to(_observed_tensor(c).cpu().numpy().reshape(-1), (plan.n_rows,)) for c in P.columns]
    data = np.stack(cols, 1) if cols else None
    loss, dp = dag_interp.run(P.ops, P.n_slots, pv, data, eps, dtype=dtype)
    return loss, {names[id(p)]: dp[i] for i, p in enumerate(P.params)}


@pytest.mark.parametrize("tag", sorted(CASES))
def test_compiled_program_matches_reference_on_cpu(tag):
    g = load_golden(tag)
    ns, model, plan, names = build(tag, "cpu")
    assert plan.family == "dag (K1)"
    assert set(names.values()) == set(g["param"])            # same learnable names as the reference
    for p in plan.prog.params:                               # same initial values (same construction script)
        np.testing.assert_allclose(float(p.detach()), float(g["param"][names[id(p)]]), rtol=1e-6, atol=1e-7)
    loss, grads = interp(plan, names, g)
    assert_close(loss, g["raw"]["loss"], tag + " loss", rtol=1e-5, atol=1e-6)
    sc = max(abs(float(v)) for v in g["grad"].values())
    for k, v in g["grad"].items():
        assert_close(grads[k], v.reshape(()), tag + " grad " + k, rtol=1e-5, atol=1e-6, scale=sc)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_device_table_layout(tag):
    """Host logic of the device table (lowering.DagProgram.device_ops): common subexpressions are merged, the
    sample-independent ops sit in the UNIFORM_HEADER / LEVEL segment with every operand defined in an EARLIER level, the
    per-sample ops keep their order, and walking the table in device order (markers skipped) reproduces the plain
    program's loss and gradients in the numpy interpreter."""
    from oracle import dag_interp
    from brancher_b200.lowering import _DAG, _observed_tensor
    g = load_golden(tag)
    ns, model, plan, names = build(tag, "cpu")
    P = plan.prog
    inv = {v: k for k, v in _DAG.items()}
    pure = [o for o in P.ops if inv[o[0]] not in ("EPS", "DATA", "ACC_SAMPLE", "ACC_ROW")]
    norm = lambda o: (o[0],) + (tuple(sorted(o[2:4])) if inv[o[0]] in ("ADD", "MUL") else tuple(o[2:4])) + tuple(o[4:])
    assert len({norm(o) for o in pure}) == len(pure), "duplicate pure op survived CSE"
    dev = P.device_ops()
    assert dev[0][0] == 26
    n_uniform, n_eps_ops, n_data_ops = dev[0][2], dev[0][3], dev[0][4]
    seg, rest = dev[1:1 + n_uniform], dev[1 + n_uniform:]
    # per-sample segment: noise draws, then data loads, then the body in program order; an ACC whose operand is produced by
    # a body op is folded into that op as a flag (bit 6: ACC_SAMPLE, bit 7: ACC_ROW of the opcode byte)
    assert all(inv[o[0]] == "EPS" for o in rest[:n_eps_ops])
    assert all(inv[o[0]] == "DATA" for o in rest[n_eps_ops:n_eps_ops + n_data_ops])
    body = rest[n_eps_ops + n_data_ops:]
    assert all(inv[o[0] & 0x3f] not in ("EPS", "DATA") for o in body)
    expanded, extra = [], P.n_slots            # fused accumulations written out again (fresh dst slots)
    for o in body:
        expanded.append((o[0] & 0x3f,) + tuple(o[1:]))
        for bit, acc in ((0x40, "ACC_SAMPLE"), (0x80, "ACC_ROW")):
            if o[0] & bit:
                expanded.append((_DAG[acc], extra, o[1], 0, 0, 0.0))
                extra += 1
    key = lambda o: (o[0],) + tuple(o[2:])      # an op up to its destination slot
    uniform_dst = {o[1] for o in seg if o[0] < 26}
    plain_body = [o for o in P.ops if o[1] not in uniform_dst and inv[o[0]] not in ("EPS", "DATA")]
    assert sorted(key(o) for o in expanded) == sorted(key(o) for o in plain_body), "per-sample ops changed"
    kept = [o[1] for o in expanded if o[1] < P.n_slots]
    assert kept == [o[1] for o in plain_body if o[1] in set(kept)], "per-sample ops must keep their order"
    defined, i, prev = set(), 0, 0
    while i < len(seg):
        assert seg[i][0] == 27 and seg[i][3] == prev
        cnt = seg[i][2]
        level_ops = seg[i + 1:i + 1 + cnt]
        for o in level_ops:
            op = inv[o[0]]
            assert op not in ("EPS", "DATA", "ACC_SAMPLE", "ACC_ROW")
            ins = () if op in ("CONST", "PARAM") else ((o[2], o[3], o[4]) if op == "NORMAL_LP" else
                                                       ((o[2], o[3]) if op in ("ADD", "SUB", "MUL", "DIV") else (o[2],)))
            assert all(x in defined for x in ins), "operand of a uniform op is not defined in an earlier level"
        defined |= {o[1] for o in level_ops}
        prev, i = cnt, i + 1 + cnt
    walk = [o for o in seg if o[0] < 26] + list(rest[:n_eps_ops + n_data_ops]) + expanded
    eps = np.stack([g["eps"][n] for n in P.eps_names], 1)
    pv = np.array([g["param"][names[id(p)]].reshape(()) for p in P.params])
    cols = [np.broadcast_to(_observed_tensor(c).cpu().numpy().reshape(-1), (plan.n_rows,)) for c in P.columns]
    data = np.stack(cols, 1) if cols else None
    l0, g0 = dag_interp.run(P.ops, P.n_slots, pv, data, eps)
    l1, g1 = dag_interp.run(walk, extra, pv, data, eps)
    np.testing.assert_allclose(l1, l0, rtol=1e-12)
    np.testing.assert_allclose(g1, g0, rtol=1e-10, atol=1e-12)


def test_unsupported_scalar_graphs_raise():
    from brancher_b200 import config, lowering
    config.set_device("cpu")
    ns = zoo.namespace("brancher_b200")
    # local latent per data row: z_b ~ N(0,1) with an amortised q that depends on the observed rows
    x = ns.DeterministicVariable(np.linspace(0, 1, 5), name="x", is_observed=True)
    z = ns.NormalVariable(0., 1., "z")
    y = ns.NormalVariable(z * x, 1., "y")
    model = ns.ProbabilisticModel([y])
    y.observe(np.zeros((5, 1, 1), "float32"))
    Qz = ns.NormalVariable(ns.DeterministicVariable(0., "a", learnable=True) * x, 1., "z", learnable=True)
    model.set_posterior_model(ns.ProbabilisticModel([Qz]))
    with pytest.raises(lowering.UnsupportedModelError):
        lowering.get_plan(model, model.posterior_model)
    # Binomial with total_count != 1 (needs lgamma terms) and a Cauchy q variable (needs a uniform noise stream): not lowered
    x2 = ns.DeterministicVariable(np.linspace(-1, 1, 6), name="x", is_observed=True)
    w = ns.NormalVariable(0., 1., "w")
    k = ns.BinomialVariable(3, logits=w * x2, name="k")
    m2 = ns.ProbabilisticModel([k])
    k.observe(np.ones((6, 1, 1), "float32"))
    m2.set_posterior_model(ns.ProbabilisticModel([ns.NormalVariable(0., 1., "w", learnable=True)]))
    with pytest.raises(lowering.UnsupportedModelError):
        lowering.get_plan(m2, m2.posterior_model)
    w3 = ns.NormalVariable(0., 1., "w")
    y3 = ns.NormalVariable(w3 * x2, 1., "y")
    m3 = ns.ProbabilisticModel([y3])
    y3.observe(np.zeros((6, 1, 1), "float32"))
    m3.set_posterior_model(ns.ProbabilisticModel([ns.CauchyVariable(0., 1., "w", learnable=True)]))
    with pytest.raises(lowering.UnsupportedModelError):
        lowering.get_plan(m3, m3.posterior_model)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CASES))
def test_dag_kernel_matches_reference(tag):
    from brancher_b200 import lowering
    g = load_golden(tag)
    ns, model, plan, names = build(tag, "cuda:0")
    S = next(iter(g["eps"].values())).shape[0]
    with lowering.inject_noise(g["eps"]):
        loss = ns.inference.ReverseKL().compute_loss(model, model.posterior_model, None, S)
    loss.backward()
    l64, g64 = interp(plan, names, g)
    assert_close(float(loss.detach()), g["raw"]["loss"], tag + " loss vs reference", rtol=1e-5, atol=1e-6)
    assert_close(float(loss.detach()), l64, tag + " loss vs fp64 interpreter")
    sc = max(abs(float(v)) for v in g["grad"].values())
    for p in plan.prog.params:
        k = names[id(p)]
        got = float(p.grad.reshape(()))
        assert_close(got, g["grad"][k].reshape(()), tag + " grad %s vs reference" % k, rtol=1e-5, atol=1e-6, scale=sc)
        assert_close(got, g64[k], tag + " grad %s vs fp64 interpreter" % k, scale=sc)


@pytest.mark.gpu
def test_dag_philox_mode_and_shard_invariance():
    """Philox noise: re-materialise the same eps with brn_philox_normal_fill and compare with the injected-noise run;
    evaluate in two sample shards and check the partial sums add up."""
    from brancher_b200 import _cuda as cu, config, lowering
    ns, model, plan, names = build("multivariate_regression", "cuda:0")
    P, S, dev = plan.prog, 37, torch.device("cuda:0")
    ops = torch.from_numpy(P.table().view(np.uint8)).to(dev)
    pvec = torch.stack([p.detach().reshape(()) for p in P.params])
    data = torch.stack([lowering._observed_tensor(c).reshape(-1).float().expand(plan.n_rows) for c in P.columns], 1).contiguous()
    r = cu.sample_range(S, seed=11, offset=3)
    loss, gr = cu.dag_elbo_fwd_bwd(ops, ops.numel() // 24, P.n_slots, pvec, data, plan.n_rows, None, len(P.eps_names), r)
    eps = torch.cat([cu.philox_normal(1, k, r, dev) for k in range(len(P.eps_names))], 1).contiguous()
    loss2, gr2 = cu.dag_elbo_fwd_bwd(ops, ops.numel() // 24, P.n_slots, pvec, data, plan.n_rows, eps, len(P.eps_names), r)
    assert_close(loss.item(), loss2.item(), "philox vs injected loss", rtol=1e-6, atol=1e-6)
    assert_close(gr.cpu().numpy(), gr2.cpu().numpy(), "philox vs injected grads", rtol=1e-6, atol=1e-6, scale=float(gr2.abs().max()))
    lsum, gsum = torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros_like(gr)
    for s0, n in [(0, 20), (20, 17)]:
        rr = cu.sample_range(S, s0=s0, s_local=n, seed=11, offset=3)
        l, g_ = cu.dag_elbo_fwd_bwd(ops, ops.numel() // 24, P.n_slots, pvec, data, plan.n_rows, None, len(P.eps_names), rr)
        lsum += l; gsum += g_
    assert_close(lsum.item(), loss.item(), "sharded loss", rtol=1e-6, atol=1e-6)
    assert_close(gsum.cpu().numpy(), gr.cpu().numpy(), "sharded grads", rtol=1e-5, atol=1e-6, scale=float(gr.abs().max()))


@pytest.mark.gpu
def test_ar1_perform_inference_readme_config():
    """BASELINE config C1: README AR(1), T=20, 300 MC samples, SGD -- loss must decrease."""
    from brancher_b200 import config
    config.set_device("cuda:0")
    ns = zoo.namespace("brancher_b200")
    model, Q, d = zoo.ar1(ns, 6, 20)
    ns.inference.perform_inference(model, number_iterations=60, number_samples=300, optimizer="SGD", lr=0.001)
    curve = model.diagnostics["loss curve"]
    assert curve.shape == (60,) and np.isfinite(curve).all()
    assert curve[-10:].mean() < curve[:10].mean()


def _plain_table(P):
    """Device table WITHOUT the uniform segment: the logical program as is."""
    t = np.zeros(len(P.ops), dtype=P.table().dtype)
    for i, o in enumerate(P.ops):
        t[i] = o
    return t


@pytest.mark.gpu
@pytest.mark.parametrize("name,S", [("ar1_readme", 300), ("ar1_readme", 45000), ("multivariate_regression", 33),
                                    ("multivariate_regression", 4000), ("lognormal_normal", 7)])
def test_dag_uniform_segment_matches_plain_program(name, S):
    """The hoisted, lane-parallel uniform segment (UNIFORM_HEADER / LEVEL layout) against the plain per-thread walk of the
    same program, same Philox noise: small S runs the shared-memory-frame kernel (one warp per CTA, uniform phases active),
    large S the thread-local-frame kernel (markers skipped, every thread evaluates the uniform ops itself)."""
    from brancher_b200 import _cuda as cu, lowering
    ns, model, plan, names = build(name, "cuda:0")
    P, dev = plan.prog, torch.device("cuda:0")
    layout = P.table()
    assert layout[0]["opcode"] == 26 and (layout["opcode"] == 27).sum() >= 1        # header + level markers present
    pvec = torch.stack([p.detach().reshape(()) for p in P.params])
    data = None
    if P.columns:
        data = torch.stack([lowering._observed_tensor(c).reshape(-1).float().expand(plan.n_rows) for c in P.columns], 1).contiguous()
    r = cu.sample_range(S, seed=5, offset=2)
    out = []
    for t in (layout, _plain_table(P)):
        ops = torch.from_numpy(t.view(np.uint8)).to(dev)
        loss, gr = cu.dag_elbo_fwd_bwd(ops, ops.numel() // 24, P.n_slots, pvec, data, plan.n_rows, None, len(P.eps_names), r)
        out.append((loss.item(), gr.cpu().numpy()))
    assert_close(out[0][0], out[1][0], "%s S=%d loss: uniform segment vs plain" % (name, S), rtol=2e-6, atol=1e-6)
    assert_close(out[0][1], out[1][1], "%s S=%d grads: uniform segment vs plain" % (name, S), rtol=1e-5, atol=1e-6,
                 scale=float(np.abs(out[1][1]).max()))


def test_bernoulli_node_lowers_like_binomial_one():
    """BernulliVariable(logits=...) (distributions.py:578-592) and BinomialVariable(1, logits=...) (:561-575) have the same
    log-probability: the scalar-DAG lowering must emit the same program for both (host logic, CPU)."""
    from brancher_b200 import config, lowering
    config.set_device("cpu")
    ns = zoo.namespace("brancher_b200")
    import importlib
    SV = importlib.import_module("brancher_b200.standard_variables")

    def build_with(make_k):
        xv = np.linspace(-2., 2., 12)
        x = ns.DeterministicVariable(xv, name="x", is_observed=True)
        w = ns.NormalVariable(0., 1., name="w")
        b = ns.NormalVariable(0., 1., name="b")
        k = make_k(w * x + b)
        model = ns.ProbabilisticModel([k])
        model.set_posterior_model(ns.ProbabilisticModel([ns.NormalVariable(0.3, 0.6, name="w", learnable=True),
                                                          ns.NormalVariable(-0.2, 0.8, name="b", learnable=True)]))
        k.observe((xv > 0).astype("float32").reshape(12, 1, 1))
        return lowering.get_plan(model, model.posterior_model).prog.ops

    ops_bin = build_with(lambda l: ns.BinomialVariable(1, logits=l, name="k"))
    ops_ber = build_with(lambda l: SV.BernulliVariable(logits=l, name="k"))
    assert ops_bin == ops_ber and len(ops_bin) > 10
