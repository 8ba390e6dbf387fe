"""Diagnostic (GPU box): accumulation error of the tcgen05 tf32 path vs K, with tf32-exact inputs (products exact)."""
import ctypes as C, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from test_gpu_gemm import _gemm
g = torch.Generator(device="cuda").manual_seed(0)
def tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
for relu_in in (False, True):
    for K in (32, 64, 128, 256, 512, 1024, 2048, 4096, 16384):
        M, N = 256, 256
        A = tf32(torch.randn(M, K, device="cuda", generator=g)); B = tf32(torch.randn(N, K, device="cuda", generator=g))
        if relu_in: A = torch.relu(A)
        ref = A.double() @ B.double().t()
        res = {}
        for prec in ("fp32", "tf32", "tf32x3"):
            out = torch.empty(M, N, device="cuda")
            _gemm(A, B, out, M, N, K, 0, 1, prec)
            d = out.double() - ref
            res[prec] = (float(d.norm() / ref.norm()), float((d * torch.sign(ref)).mean() / ref.abs().mean()))
        print("relu_in", relu_in, "K", K, {k: "rel %.2e bias %.2e" % v for k, v in res.items()})
# does the tensor core truncate or round fp32 operands to tf32?
A = torch.full((128, 32), 1.0 + 2**-11 + 2**-12, device="cuda"); B = torch.zeros(64, 32, device="cuda"); B[:, 0] = 1.0
out = torch.empty(128, 64, device="cuda")
_gemm(A, B, out, 128, 64, 32, 0, 1, "tf32")
print("operand 1+2^-11+2^-12 ->", float(out[0, 0]) - 1.0, "(truncate: 0, round-nearest: 2^-10 =", 2**-10, ")")
