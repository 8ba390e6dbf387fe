"""GPU box: every dense-layer shape of one optimizer step (4096 envs: M = 16384 / 16385 rows) through addk_gemm, operand
twins ready, 7 timed launches each (median), with the ideal time at the measured sustained tensor peak and the clock stamps
of CTA 0 of the persistent kernel.      python tools/gemm_shapes.py [f16x3|bf16] [M]"""
import ctypes as C, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import _lib
prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
MB = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
dev = "cuda:0"
L = _lib.lib()
H3, BF = prec == "f16x3", prec == "bf16"
PEAK = 1392.7 / (3 if H3 else 1)      # TFLOP/s algorithmic ceiling of the scheme at the sustained bf16 rate
dt16 = torch.float16 if H3 else torch.bfloat16

class Op:
    def __init__(self, rows, cols, ld=None, scale=1.0):
        ld = ld or cols
        self.t = torch.randn(rows, ld, device=dev) * scale
        self.n = (self.t.numel() + 7) & ~7
        self.p = torch.zeros(2 * self.n if H3 else self.n, device=dev, dtype=dt16)
        self.s = torch.zeros(2, device=dev, dtype=torch.int32)
        if BF:
            self.p[:self.t.numel()] = self.t.flatten().to(dt16)
        self.ld = ld

def run(name, M, N, K, ta, tb, A, B, bias=False, relu=0, mask=False, split=1, ldc=None, reps=7, planes=False):
    # planes (f16x3, PLANES=1): the output exists only as fp16 planes written by the epilogue with the sticky scale, and
    # the ReLU mask is read from the hi plane of a 16-bit tensor -- the variant the optimizer step runs for h1, h2, g2, g1, e1, u1
    planes = planes and H3 and PLANES
    ldc = ldc or N
    out = torch.empty(split, M, ldc, device=dev) if split > 1 else torch.empty(M, ldc, device=dev)
    out16 = torch.empty(M, ldc, device=dev, dtype=dt16) if (BF and split == 1) else None
    bvec = torch.zeros(N, device=dev) if bias else None
    mk = torch.randn(M, ldc, device=dev) if mask else None
    mk16 = mk.to(torch.float16) if (mask and planes) else None
    cpl = torch.zeros(2 * M * ldc, device=dev, dtype=torch.float16) if planes else None
    cslot = torch.zeros(2, device=dev, dtype=torch.int32) if planes else None
    use_bits = BITS and H3 and (mask or relu) and N % 128 == 0 and split == 1      # ReLU masks as bit planes (in or out)
    bits = torch.randint(-2**31, 2**31 - 1, (M, ldc // 32), device=dev, dtype=torch.int32) if use_bits else None
    ready = [0]
    def prep():
        if planes:
            _lib.check(L.addk_f16x3_prep(_lib.stream(), C.c_void_p(cslot.data_ptr()), C.c_int(1)), "prep")
    def launch():
        a = _lib.AddkGemmArgs(A=A.t.data_ptr(), lda=A.ld, B=B.t.data_ptr(), ldb=B.ld, C=out.data_ptr(), ldc=ldc, M=M, N=N, K=K,
                              bias=bvec.data_ptr() if bias else None, a_mean=None, a_std=None,
                              relu_mask_src=mk.data_ptr() if mask else None, ld_mask=ldc if mask else 0, trans_a=ta, trans_b=tb,
                              relu=relu, split_k=split, accumulate=0, slab_stride=0, A16=A.p.data_ptr(), B16=B.p.data_ptr(),
                              C16=out16.data_ptr() if out16 is not None else None)
        if H3:
            a.a16_plane, a.b16_plane, a.a_amax, a.b_amax = A.n, B.n, A.s.data_ptr(), B.s.data_ptr()
            a.a16_ready = a.b16_ready = ready[0]
        if BF and NO_F32 and split == 1 and N > 128:
            a.no_f32 = 1
        if planes:
            a.C16, a.c16_plane, a.c_amax, a.no_f32 = cpl.data_ptr(), M * ldc, cslot.data_ptr(), 1
            if mask:
                a.relu_mask_src16, a.relu_mask_src = mk16.data_ptr(), None
        if use_bits:
            a.ld_bits = ldc // 32
            if mask:
                a.relu_bits_in = bits.data_ptr()
            else:
                a.relu_bits_out = bits.data_ptr()
        _lib.check(L.addk_gemm(_lib.stream(), C.byref(a), C.c_int(_lib.PRECISIONS[prec])), "addk_gemm " + name)
    prep(); launch(); ready[0] = 1; prep(); launch()
    torch.cuda.synchronize()
    flush = torch.empty(64 * 1024 * 1024, device=dev)      # 256 MB: evict L2 between launches
    ev = []
    for _ in range(reps):
        flush.zero_()
        prep()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); ev.append((e0, e1))
    torch.cuda.synchronize()
    us = sorted(a.elapsed_time(b) for a, b in ev)[reps // 2] * 1e3
    kern = L.addk_debug_last_gemm_kernel()
    stamps = ""
    if kern in (21, 22, 31, 32):
        dbg = torch.zeros(16, dtype=torch.int64, device=dev)
        L.addk_debug_set_stamp_buffer(C.c_void_p(dbg.data_ptr()))
        prep(); launch(); torch.cuda.synchronize()
        L.addk_debug_set_stamp_buffer(C.c_void_p(0))
        t = dbg.tolist()
        stamps = " | CTA0 kcyc: mma %d (wait acc %d, wait full %d) workers %d (wait acc_full %d, drains %d, epilogue %d)" % tuple(x // 1000 for x in t[:7])
    fl = 2.0 * M * N * K
    ideal = fl / PEAK / 1e6
    print("%-34s M=%-6d N=%-5d K=%-6d k%-2d %7.1f us  %6.1f TF  ideal %5.1f us (%.0f%%)%s" % (name, M, N, K, kern, us, fl / us / 1e6, ideal, 100 * ideal / us, stamps))
    return us, ideal

R = MB + 1
tot = [0.0, 0.0]
QUICK = os.environ.get("QUICK") == "1"
NO_F32 = os.environ.get("NO_F32", "1") == "1"
PLANES = os.environ.get("PLANES", "1") == "1"
BITS = os.environ.get("BITS", "1") == "1"
def acc(r, times=1):
    tot[0] += r[0] * times; tot[1] += r[1] * times
X272, H1024, H512 = Op(R, 264, 272), Op(R, 1024), Op(R, 512)
W0, W1, W2 = Op(1024, 264, 272, 0.05), Op(1024, 1024, scale=0.03), Op(512, 1024, scale=0.03)
G1024, G512 = Op(R, 1024, scale=1e-5), Op(R, 512, scale=1e-5)
Dn, Wd0, Wd1 = Op(R, 114, 128), Op(1024, 114, 128, 0.05), Op(512, 1024, scale=0.03)
Wm, Dm = Op(29, 512, scale=0.01), Op(R, 29, 32, 1e-5)
print("# %s, M = %d; actor/critic chain x2 unless noted" % (prec, MB))
acc(run("fwd0 [M,264]x[1024] bias relu", MB, 1024, 264, 0, 1, X272, W0, bias=True, relu=1, planes=True), 2)
acc(run("fwd1 1024x1024 bias relu", MB, 1024, 1024, 0, 1, H1024, W1, bias=True, relu=1, planes=True), 2)
acc(run("dgrad1 1024->1024 mask", MB, 1024, 1024, 0, 0, G1024, W1, mask=True, planes=True), 2)
if QUICK:
    sys.exit(0)
acc(run("fwd2 1024->512 bias relu", MB, 512, 1024, 0, 1, H1024, W2, bias=True, relu=1), 2)
acc(run("head fwd 512->29 (actor)", MB, 29, 512, 0, 1, H512, Wm, bias=True, ldc=32), 1)
acc(run("head wgrad 29x512 split9 (actor)", 29, 512, MB, 1, 0, Dm, H512, split=9), 1)
acc(run("head dgrad [M,29]->512 mask (actor)", MB, 512, 29, 0, 0, Dm, Wm, mask=True), 1)
acc(run("wgrad2 512x1024 split9", 512, 1024, MB, 1, 0, G512, H1024, split=9), 2)
acc(run("dgrad2 512->1024 mask", MB, 1024, 512, 0, 0, G512, W2, mask=True, planes=True), 2)
acc(run("wgrad1 1024x1024 split9", 1024, 1024, MB, 1, 0, G1024, H1024, split=9), 2)
acc(run("dgrad1 1024->1024 mask", MB, 1024, 1024, 0, 0, G1024, W1, mask=True, planes=True), 2)
acc(run("wgrad0 1024x264 split9", 1024, 264, MB, 1, 0, G1024, X272, split=9, ldc=264), 2)
print("# discriminator chain (R = M + 1 rows)")
acc(run("d fwd0 [R,114]x1024", R, 1024, 128, 0, 1, Dn, Wd0, bias=True, relu=1, planes=True))
acc(run("d fwd1 1024->512", R, 512, 1024, 0, 1, H1024, Wd1, bias=True, relu=1))
acc(run("d u1 = u2.W1 mask 512->1024", R, 1024, 512, 0, 0, G512, Wd1, mask=True, planes=True), 2)
acc(run("d gx = u1.W0 1024->128", R, 128, 1024, 0, 0, G1024, Wd0))
acc(run("d wgrad0 1024x114 split9", 1024, 114, R, 1, 0, G1024, Dn, split=9, ldc=114), 2)
acc(run("d dv1 = dg.W0^T 128->1024 mask", R, 1024, 128, 0, 1, Dn, Wd0, mask=True, planes=True))
acc(run("d wgrad1 512x1024 split9", 512, 1024, R, 1, 0, G512, H1024, split=9), 2)
acc(run("d du2 = dv1.W1^T 1024->512 mask", R, 512, 1024, 0, 1, H1024, Wd1, mask=True))
print("# sum over one optimizer step: %.0f us measured, %.0f us ideal (%.0f%%)" % (tot[0], tot[1], 100 * tot[1] / tot[0]))
