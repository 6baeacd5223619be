"""TEST INFRASTRUCTURE ONLY -- never imported by the product package `brancher_b200`.

CPU restatement (torch CPU tensors + autograd, fp32 or fp64) of the reference's
Monte-Carlo ELBO / pathwise-gradient hot path and of its SVGD direction.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package.

Parity status: PINNED -- the reference has no golden vectors of its own,
so every function here is checked against outputs of the reference itself, generated in
the build container by `tests/golden/make_golden.py` (which imports `/root/reference`) and
committed under `tests/golden/*.npz`.  `tests/test_oracle_golden.py` replays them.
"""
