"""Probe for MN-major tf32 operands (launch_umma_tn_plain through the stand-alone GEMM entry, BRN_GEMM_BN=-300): error structure
against torch for a few shapes.  History: with plain SWIZZLE_128B boxes / layout type 2 every product was wrong (max err/scale
0.3-0.8 at every shape, K = 8 included); with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B / layout type 1 / SBO = 512 all shapes agree
to 1e-7 .. 4e-7 of scale (SBO = 1024 there is an illegal address).  The kernel had a BRN_MN_DESC=kstep,lbo,sbo knob for this."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from brancher_b200 import _cuda as cu
os.environ["BRN_GEMM_BN"] = "-300"
torch.manual_seed(0)
for desc in ["default"]:
    for (M, N, K) in ((128, 128, 8), (128, 128, 16), (128, 128, 64), (256, 256, 64), (200, 136, 300)):
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
        D = cu.gemm_nt_3xtf32(A, B); torch.cuda.synchronize()
        ref = (A.double() @ B.double().T)
        err = (D.double() - ref).abs()
        scale = (A.abs().double() @ B.abs().double().T).max().item()
        bad = err > 2e-6 * scale
        rows_bad = bad.any(1).nonzero().flatten().tolist(); cols_bad = bad.any(0).nonzero().flatten().tolist()
        print(desc, (M, N, K), "max err/scale %.2e" % (err.max().item() / scale), "bad %d/%d" % (bad.sum().item(), bad.numel()),
              "bad rows", rows_bad[:6], len(rows_bad), "bad cols", cols_bad[:6], len(cols_bad), flush=True)
