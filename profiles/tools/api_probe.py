import sys, os, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import model_zoo as zoo
from brancher_b200 import config, inference
dev = torch.device("cuda:0"); config.set_device(dev)
ns = zoo.namespace("brancher_b200")
model = zoo.bnn(ns, 0, B=1024, P=784, H=100, C=10)[0]
for k in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    inference.perform_inference(model, number_iterations=200 if k else 3, number_samples=256, optimizer="Adam", lr=1e-3, inference_method=inference.ReverseKL())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(k, "%.1f ms total" % (1e3 * dt), inference.last_loop, flush=True)
