"""Diagnostic (GPU box): accuracy of the f16x3 path vs K and vs the accumulator-truncation compensation
(env ADDK_H3_COMP, read once per process): python tools/h3_accuracy.py"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from test_gpu_gemm import _gemm, _gemm_h3, _Twin
g = torch.Generator(device="cuda").manual_seed(0)
print("ADDK_H3_COMP =", os.environ.get("ADDK_H3_COMP"))
for relu_in in (False, True):
    for K in (32, 256, 512, 1024, 2048, 16384):
        M, N = 256, 256
        A = torch.randn(M, K, device="cuda", generator=g); B = torch.randn(N, K, device="cuda", generator=g)
        if relu_in: A = torch.relu(A); B = B.abs()
        ref = A.double() @ B.double().t()
        res = {}
        for prec in ("fp32", "tf32x3", "f16x3"):
            out = torch.empty(M, N, device="cuda")
            if prec == "f16x3": _gemm_h3(A, B, out, M, N, K, 0, 1, _Twin(A), _Twin(B))
            else: _gemm(A, B, out, M, N, K, 0, 1, prec)
            d = out.double() - ref
            res[prec] = (float(d.norm() / ref.norm()), float((d * torch.sign(ref)).mean() / ref.abs().mean()))
        print("relu_in", relu_in, "K", K, {k: "rel %.2e bias %.2e" % v for k, v in res.items()})
