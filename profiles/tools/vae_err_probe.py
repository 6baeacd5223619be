#!/usr/bin/env python
"""Error of the K5 CUDA path against the fp64 oracle per tensor, relative to the tensor's scale.
    python profiles/tools/vae_err_probe.py B S [scale]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from brancher_b200 import _cuda as cu
from oracle import elbo_oracle as O
import test_vae_cuda as V

B, S = int(sys.argv[1]), int(sys.argv[2])
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
D, L = 784, 2
X, enc, dec, eps = V.random_vae(31, B, D, L, (256, 512), (512, 256), S, scale=scale)
net = V.make_net(cu, enc, dec)
loss = cu.vae_elbo_fwd_bwd(V.dev(X), net, cu.sample_range(S), eps=V.dev(eps)).item()
l64, g64 = O.vae_elbo(X, enc, dec, eps, dtype=torch.float64, row_chunk=512)
g = V.grads_of(net)
print("B=%d S=%d scale=%g loss rel err %.1e" % (B, S, scale, abs(loss - l64) / abs(l64)))
print("  " + "  ".join("%s %.1e" % (k, np.abs(g[k].reshape(g64[k].shape) - g64[k]).max() / np.abs(g64[k]).max()) for k in sorted(g64)))
net2 = V.make_net(cu, enc, dec)
loss2 = cu.vae_elbo_fwd_bwd(V.dev(X), net2, cu.sample_range(S), eps=V.dev(eps)).item()
g2 = V.grads_of(net2)
print("  run-to-run: " + "  ".join("%s %.1e" % (k, np.abs(g[k] - g2[k]).max() / np.abs(g64[k]).max()) for k in sorted(g64)))
# fp32 oracle for comparison
l32, g32 = O.vae_elbo(X, enc, dec, eps, row_chunk=512)
print("  fp32 oracle: " + "  ".join("%s %.1e" % (k, np.abs(g32[k] - g64[k]).max() / np.abs(g64[k]).max()) for k in sorted(g64)))
