"""GPU: the dense-layer contraction (addk_gemm) in every operand layout the MLP forward / backward uses, for the
exact-fp32 CUDA-core kernel and the tcgen05 tensor-core modes, against a float64 torch reference.

Tolerances (norm-wise relative error vs float64): fp32 2e-6; tf32x3 (3-pass split with the main accumulator drained
into fp32 registers every 256 k, the fp32-parity tensor-core mode) 2e-6 as well -- measured 3.3e-7 at any K; tf32
(single pass, operands truncated to 10 mantissa bits, accumulator truncation -2.05e-9 x K) 2e-3.
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-6, "tf32x3": 2e-6, "tf32": 2e-3, "f16x3": 2e-6}
K_SGEMM, K_BF16_TILE, K_BF16_PERSISTENT, K_H3_TILE = 0, 20, 21, 30     # addk_debug_last_gemm_kernel()
K_H3_PERSISTENT = (31, 32)      # persistent f16x3 kernel: one CTA per SM (31) or cta_group::2 CTA pairs (32, the default)


def _skip_without_legacy(prec):
    from add_gym_b200 import _lib
    if prec in ("tf32x3", "tf32") and not _lib.has_legacy_kernels():
        pytest.skip("precision %s: superseded kernels, only in libaddk_legacy.so (make LEGACY=1, ADDK_LIB=...)" % prec)


def _last_kernel():
    from add_gym_b200 import _lib
    return int(_lib.lib().addk_debug_last_gemm_kernel())


def _gemm(A, B, Cout, M, N, K, ta, tb, prec, bias=None, relu=0, mask=None, split=1, accumulate=0, lda=None, ldb=None):
    from add_gym_b200 import _lib
    a = _lib.AddkGemmArgs(A=A.data_ptr(), lda=lda or A.stride(0), B=B.data_ptr(), ldb=ldb or B.stride(0),
                          C=Cout.data_ptr(), ldc=Cout.stride(-2), M=M, N=N, K=K,
                          bias=bias.data_ptr() if bias is not None else None, a_mean=None, a_std=None,
                          relu_mask_src=mask.data_ptr() if mask is not None else None,
                          ld_mask=mask.stride(0) if mask is not None else 0, trans_a=ta, trans_b=tb, relu=relu,
                          split_k=split, accumulate=accumulate, slab_stride=0, A16=None, B16=None, C16=None)
    _lib.check(_lib.lib().addk_gemm(_lib.stream(), C.byref(a), C.c_int(_lib.PRECISIONS[prec])), "addk_gemm")


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


SHAPES = [
    # M, N, K   (forward / dgrad / wgrad shapes of the three MLPs, plus ragged edges)
    (512, 1024, 264), (300, 512, 1024), (129, 1024, 1024), (1024, 264, 2048), (1024, 114, 1500), (256, 256, 32),
    (128, 64, 40), (16385 // 8, 512, 1024),
    (29, 512, 2048), (2048, 29, 512), (2048, 512, 29), (1024, 114, 1000), (40, 24, 20),   # heads: tiles mostly out of bounds
]


@pytest.mark.parametrize("prec", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_gemm_layouts(prec, ta, tb):
    _skip_without_legacy(prec)
    g = torch.Generator(device="cuda").manual_seed(1)
    for (M, N, K) in SHAPES:
        pad = lambda n: (n + 3) & ~3
        A = torch.randn((K, pad(M)) if ta else (M, pad(K)), device="cuda", generator=g)
        B = torch.randn((N, pad(K)) if tb else (K, pad(N)), device="cuda", generator=g)
        Aop = (A[:, :M].t() if ta else A[:, :K]).double()
        Bop = (B[:, :K].t() if tb else B[:, :N]).double()
        ref = Aop @ Bop
        out = torch.full((M, pad(N)), float("nan"), device="cuda")
        _gemm(A, B, out, M, N, K, ta, tb, prec)
        torch.cuda.synchronize()
        e = _rel(out[:, :N], ref)
        assert e <= TOL[prec], "%s ta=%d tb=%d %s: rel err %.3e" % (prec, ta, tb, (M, N, K), e)
        if pad(N) != N:
            assert torch.isnan(out[:, N:]).all(), "columns beyond N must not be written"


@pytest.mark.parametrize("prec", ["fp32", "tf32x3", "tf32"])
def test_gemm_epilogues_and_split_k(prec):
    _skip_without_legacy(prec)
    g = torch.Generator(device="cuda").manual_seed(2)
    M, N, K = 1000, 512, 1024
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    _gemm(A, W, out, M, N, K, 0, 1, prec, bias=bias, relu=1)
    ref = torch.relu(A.double() @ W.double().t() + bias.double())
    assert _rel(out, ref) <= TOL[prec]
    # ReLU-mask epilogue (input gradient through a ReLU layer) and accumulate
    h = torch.randn(M, K, device="cuda", generator=g)
    dY = torch.randn(M, N, device="cuda", generator=g)
    dX = torch.empty(M, K, device="cuda")
    _gemm(dY, W, dX, M, K, N, 0, 0, prec, mask=h)
    ref = (dY.double() @ W.double()) * (h > 0).double()
    assert _rel(dX, ref) <= TOL[prec]
    base = torch.randn(M, K, device="cuda", generator=g)
    acc = base.clone()
    _gemm(dY, W, acc, M, K, N, 0, 0, prec, accumulate=1)
    assert _rel(acc, base.double() + dY.double() @ W.double()) <= TOL[prec]
    # split-K weight gradient: dW[N,K] = dY^T X as 8 partial slabs
    S = 8
    slabs = torch.full((S, N, K), float("nan"), device="cuda")
    _gemm(dY, A, slabs, N, K, M, 1, 0, prec, split=S)
    torch.cuda.synchronize()
    ref = dY.double().t() @ A.double()
    assert _rel(slabs.sum(0), ref) <= TOL[prec]


def test_tensor_core_split_is_fp32_class_on_mlp_scale_data():
    """tf32x3 on data shaped like the MLP activations (non-negative post-ReLU inputs, 1/sqrt(K) weights): the error
    must stay an order of magnitude inside the 1e-5 parity bar so the stacked layers still meet it."""
    _skip_without_legacy("tf32x3")
    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 4096, 1024, 1024
    A = torch.relu(torch.randn(M, K, device="cuda", generator=g))
    W = (torch.rand(N, K, device="cuda", generator=g) * 2 - 1) / 32.0
    out = torch.empty(M, N, device="cuda")
    _gemm(A, W, out, M, N, K, 0, 1, "tf32x3")
    e = _rel(out, A.double() @ W.double().t())
    assert e <= 1e-6, e


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_gemm_bf16_layouts(ta, tb):
    """precision "bf16": bf16 twins of the operands through TMA + tcgen05 kind::f16, fp32 accumulate.  Checked against
    the float64 product of the SAME bf16 values (so only the accumulation differs): 1e-5; and the fp32 + bf16 outputs
    must agree to bf16 rounding."""
    from add_gym_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(4)
    pad = lambda n: (n + 7) & ~7
    for (M, N, K) in [(512, 1024, 264), (300, 512, 1024), (1024, 264, 2048), (256, 256, 64), (129, 200, 72), (29, 512, 2048),
                      (2048, 29, 512)]:
        A = torch.randn((K, pad(M)) if ta else (M, pad(K)), device="cuda", generator=g)
        B = torch.randn((N, pad(K)) if tb else (K, pad(N)), device="cuda", generator=g)
        A16, B16 = A.to(torch.bfloat16), B.to(torch.bfloat16)
        Aop = (A16[:, :M].t() if ta else A16[:, :K]).double()
        Bop = (B16[:, :K].t() if tb else B16[:, :N]).double()
        ref = Aop @ Bop
        out = torch.full((M, pad(N)), float("nan"), device="cuda")
        out16 = torch.zeros((M, pad(N)), device="cuda", dtype=torch.bfloat16)
        a = _lib.AddkGemmArgs(A=A.data_ptr(), lda=A.stride(0), B=B.data_ptr(), ldb=B.stride(0), C=out.data_ptr(), ldc=out.stride(0),
                              M=M, N=N, K=K, bias=None, a_mean=None, a_std=None, relu_mask_src=None, ld_mask=0, trans_a=ta,
                              trans_b=tb, relu=0, split_k=1, accumulate=0, slab_stride=0, A16=A16.data_ptr(),
                              B16=B16.data_ptr(), C16=out16.data_ptr())
        _lib.check(_lib.lib().addk_gemm(_lib.stream(), C.byref(a), C.c_int(_lib.PRECISIONS["bf16"])), "addk_gemm")
        torch.cuda.synchronize()
        e = _rel(out[:, :N], ref)
        assert e <= 1e-5, "bf16 ta=%d tb=%d %s: rel err %.3e" % (ta, tb, (M, N, K), e)
        assert torch.equal(out16[:, :N], out[:, :N].to(torch.bfloat16)), "bf16 copy of the output"


# ---------------------------------------------------------------------------------------------------------------------
# precision "f16x3": fp16 hi/lo planes (scaled by the tensor's max|x|) on the fp16 tensor pipe, fp32-class accuracy
# ---------------------------------------------------------------------------------------------------------------------
class _Twin:
    """fp16 hi/lo planes + the max|x| word of one fp32 operand, as the C-ABI expects them."""

    def __init__(self, t):
        self.n = (t.numel() + 7) & ~7
        self.planes = torch.zeros(2 * self.n, device=t.device, dtype=torch.float16)
        self.amax = torch.zeros(2, device=t.device, dtype=torch.int32)      # slot: {sticky scale word, max|x|}


def _gemm_h3(A, B, Cout, M, N, K, ta, tb, tw_a, tw_b, ready=(0, 0), bias=None, relu=0, mask=None, split=1, accumulate=0,
             c_amax=None, c16=None, bits_out=None, bits_in=None, no_f32=0, colpart=None):
    from add_gym_b200 import _lib
    a = _lib.AddkGemmArgs(A=A.data_ptr(), lda=A.stride(0), B=B.data_ptr(), ldb=B.stride(0), C=Cout.data_ptr(),
                          ldc=Cout.stride(-2), M=M, N=N, K=K, bias=bias.data_ptr() if bias is not None else None,
                          a_mean=None, a_std=None, relu_mask_src=mask.data_ptr() if mask is not None else None,
                          ld_mask=mask.stride(0) if mask is not None else 0, trans_a=ta, trans_b=tb, relu=relu,
                          split_k=split, accumulate=accumulate, slab_stride=0, A16=tw_a.planes.data_ptr(),
                          B16=tw_b.planes.data_ptr(), a16_plane=tw_a.n, b16_plane=tw_b.n,
                          a_amax=tw_a.amax.data_ptr(), b_amax=tw_b.amax.data_ptr(), a16_ready=ready[0], b16_ready=ready[1],
                          c_amax=c_amax.data_ptr() if c_amax is not None else None,
                          C16=c16.planes.data_ptr() if c16 is not None else None, c16_plane=c16.n if c16 is not None else 0,
                          no_f32=no_f32)
    if colpart is not None:     # [ceil(M / 32), N] column sums per 32-row block of the output
        a.colsum_partials = colpart.data_ptr()
    bits = bits_out if bits_out is not None else bits_in
    if bits is not None:      # ReLU bit plane [rows, ld_bits] int32
        a.ld_bits = bits.stride(0)
        a.relu_bits_out = bits.data_ptr() if bits_out is not None else None
        a.relu_bits_in = bits.data_ptr() if bits_in is not None else None
    _lib.check(_lib.lib().addk_gemm(_lib.stream(), C.byref(a), C.c_int(_lib.PRECISIONS["f16x3"])), "addk_gemm")


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_gemm_f16x3_layouts(ta, tb):
    g = torch.Generator(device="cuda").manual_seed(5)
    pad = lambda n: (n + 7) & ~7
    for (M, N, K) in SHAPES:
        A = torch.randn((K, pad(M)) if ta else (M, pad(K)), device="cuda", generator=g)
        B = torch.randn((N, pad(K)) if tb else (K, pad(N)), device="cuda", generator=g)
        Aop = (A[:, :M].t() if ta else A[:, :K]).double()
        Bop = (B[:, :K].t() if tb else B[:, :N]).double()
        ref = Aop @ Bop
        out = torch.full((M, pad(N)), float("nan"), device="cuda")
        tw_a, tw_b = _Twin(A), _Twin(B)
        _gemm_h3(A, B, out, M, N, K, ta, tb, tw_a, tw_b)
        torch.cuda.synchronize()
        if min(M, N, K) >= 16:   # smaller shapes run the tf32x3 fallback and leave the twins untouched
            used = A[:, :M] if ta else A[:, :K]
            # (a few alignment-padding columns may be scanned too: they belong to the operand's buffer)
            assert int(used.abs().max().view(torch.int32).item()) <= int(tw_a.amax[1].item()) <= \
                int(A.abs().max().view(torch.int32).item()), "max|A| word"
            assert bool((tw_a.planes != 0).any()), "the fp16 planes were not written: the call fell back"
        e = _rel(out[:, :N], ref)
        assert e <= TOL["f16x3"], "f16x3 ta=%d tb=%d %s: rel err %.3e" % (ta, tb, (M, N, K), e)
        if pad(N) != N:
            assert torch.isnan(out[:, N:]).all(), "columns beyond N must not be written"


def test_gemm_f16x3_epilogues_split_k_and_reuse():
    g = torch.Generator(device="cuda").manual_seed(6)
    M, N, K = 1000, 512, 1024
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    tA, tW = _Twin(A), _Twin(W)
    _gemm_h3(A, W, out, M, N, K, 0, 1, tA, tW, bias=bias, relu=1)
    ref = torch.relu(A.double() @ W.double().t() + bias.double())
    assert _rel(out, ref) <= TOL["f16x3"]
    # the twins now hold A and W: a second call may skip the conversion and must give the same bits
    out2 = torch.empty(M, N, device="cuda")
    _gemm_h3(A, W, out2, M, N, K, 0, 1, tA, tW, ready=(1, 1), bias=bias, relu=1)
    assert torch.equal(out, out2)
    # the epilogue can leave max|C| behind, so that the next layer's conversion of C skips its max pass (ready = 2)
    tOut = _Twin(out)
    out3 = torch.empty(M, N, device="cuda")
    _gemm_h3(A, W, out3, M, N, K, 0, 1, tA, tW, ready=(1, 1), bias=bias, relu=1, c_amax=tOut.amax)
    torch.cuda.synchronize()
    assert torch.equal(out3, out)
    assert int(tOut.amax[1].item()) == int(out.abs().max().view(torch.int32).item()), "max|C| word written by the epilogue"
    W2 = torch.randn(256, N, device="cuda", generator=g) * 0.05
    y = torch.empty(M, 256, device="cuda")
    tW2 = _Twin(W2)
    _gemm_h3(out, W2, y, M, 256, N, 0, 1, tOut, tW2, ready=(2, 0))
    assert _rel(y, out.double() @ W2.double().t()) <= TOL["f16x3"]
    # ... or write C's planes itself with the slot's sticky scale (prep -> layer -> repair); the next layer passes ready = 1.
    # First round: no history (W = 0) -> the repair pass writes the planes; second round: the epilogue's planes stand,
    # also when the data shrinks 8x (still inside the window); a 2^20 jump forces the repair again.
    from add_gym_b200 import _lib
    L = _lib.lib()
    slot = C.c_void_p(tOut.amax.data_ptr())
    for scale in (1.0, 1.0, 0.125, 2.0 ** 20):
        As = A * scale
        tOut.planes.zero_()
        _lib.check(L.addk_f16x3_prep(_lib.stream(), slot, C.c_int(1)), "prep")
        W_before = int(tOut.amax[0].item())
        _gemm_h3(As, W, out3, M, N, K, 0, 1, _Twin(As), tW, ready=(0, 1), bias=None, relu=1, c_amax=tOut.amax, c16=tOut)
        _lib.check(L.addk_f16x3_repair(_lib.stream(), C.c_void_p(out3.data_ptr()), C.c_longlong(M), C.c_int(N), C.c_int(N),
                                       C.c_void_p(tOut.planes.data_ptr()), C.c_longlong(tOut.n), slot), "repair")
        _gemm_h3(out3, W2, y, M, 256, N, 0, 1, tOut, tW2, ready=(1, 1))
        torch.cuda.synchronize()
        assert _rel(y, out3.double() @ W2.double().t()) <= TOL["f16x3"], scale
        assert bool((tOut.planes != 0).any())
        if scale == 0.125:
            assert W_before != 0, "sticky word carried over"
    h = torch.randn(M, K, device="cuda", generator=g)
    dY = torch.randn(M, N, device="cuda", generator=g) * 1e-6          # gradient-sized values: far below fp16's range unscaled
    dX = torch.empty(M, K, device="cuda")
    tdY = _Twin(dY)
    _gemm_h3(dY, W, dX, M, K, N, 0, 0, tdY, tW, ready=(0, 1), mask=h)   # W's twin is layout-independent
    ref = (dY.double() @ W.double()) * (h > 0).double()
    assert _rel(dX, ref) <= TOL["f16x3"]
    base = torch.randn(M, K, device="cuda", generator=g) * 1e-6
    acc = base.clone()
    _gemm_h3(dY, W, acc, M, K, N, 0, 0, tdY, tW, ready=(1, 1), accumulate=1)
    assert _rel(acc, base.double() + dY.double() @ W.double()) <= TOL["f16x3"]
    S = 8
    slabs = torch.full((S, N, K), float("nan"), device="cuda")
    _gemm_h3(dY, A, slabs, N, K, M, 1, 0, tdY, tA, ready=(1, 1), split=S)
    torch.cuda.synchronize()
    ref = dY.double().t() @ A.double()
    assert _rel(slabs.sum(0), ref) <= TOL["f16x3"]


def test_gemm_f16x3_dynamic_range_and_mlp_scale():
    """Operands whose elements span 12 orders of magnitude (one scale per tensor): the norm-wise error stays fp32-class;
    and the MLP-shaped case must stay an order of magnitude inside the 1e-5 parity bar."""
    g = torch.Generator(device="cuda").manual_seed(7)
    M, N, K = 2048, 1024, 1024
    A = torch.randn(M, K, device="cuda", generator=g) * torch.exp(torch.randn(M, K, device="cuda", generator=g) * 4.0) * 1e-5
    W = torch.randn(N, K, device="cuda", generator=g) * torch.exp(torch.randn(N, K, device="cuda", generator=g) * 4.0)
    out = torch.empty(M, N, device="cuda")
    _gemm_h3(A, W, out, M, N, K, 0, 1, _Twin(A), _Twin(W))
    assert _rel(out, A.double() @ W.double().t()) <= TOL["f16x3"]
    A = torch.relu(torch.randn(4096, K, device="cuda", generator=g))
    W = (torch.rand(N, K, device="cuda", generator=g) * 2 - 1) / 32.0
    out = torch.empty(4096, N, device="cuda")
    _gemm_h3(A, W, out, 4096, N, K, 0, 1, _Twin(A), _Twin(W))
    e = _rel(out, A.double() @ W.double().t())
    assert e <= 1e-6, e
    Z = torch.zeros(256, K, device="cuda")                              # all-zero operand: finite scale, zero output
    out = torch.full((256, N), float("nan"), device="cuda")
    _gemm_h3(Z, W, out, 256, N, K, 0, 1, _Twin(Z), _Twin(W))
    assert torch.equal(out, torch.zeros_like(out))


# ---------------------------------------------------------------------------------------------------------------------
# The PERSISTENT f16x3 kernel (gemm_tc_h3p_kernel<256>: what the benchmarked 16384-row minibatch runs, 64 % of an
# optimizer step) in every operand layout and epilogue the update uses, at the update's own shapes, against float64.
# Each case asserts through addk_debug_last_gemm_kernel() that the persistent kernel is the one that ran.
# ---------------------------------------------------------------------------------------------------------------------
MB = 16384          # minibatch rows of BASELINE configs[1]; the discriminator chain runs MB + 1 rows

H3P_CASES = [
    # name, M, N, K, ta, tb, epilogue
    ("forward 1024x1024 bias+relu", MB, 1024, 1024, 0, 1, "bias_relu"),
    ("forward first layer K=272", MB, 1024, 272, 0, 1, "bias_relu"),
    ("forward 1024->512", MB, 512, 1024, 0, 1, "bias_relu"),
    ("disc forward 16385 rows K=128", MB + 1, 1024, 128, 0, 1, "bias_relu"),
    ("input gradient dY.W masked", MB, 1024, 1024, 0, 0, "mask"),
    ("input gradient 512->1024 masked, 16385 rows", MB + 1, 1024, 512, 0, 0, "mask"),
    ("input gradient plain", MB, 512, 1024, 0, 0, None),
    ("input gradient accumulate", MB, 1024, 512, 0, 0, "accumulate"),
    ("weight gradient dY^T.X split-K 9", 1024, 1024, MB, 1, 0, "split9"),
    ("weight gradient 512x1024 split-K 9, 16385 rows", 512, 1024, MB + 1, 1, 0, "split9"),
    ("weight gradient first layer N=272 split-K 9", 1024, 272, MB, 1, 0, "split9"),
    ("A^T.B^T (both operands transposed)", 8192, 1024, 1024, 1, 1, None),
    # ragged last n-tile: contracted at its effective width (N = 32 .. 224 instead of 256; each CTA of the pair then
    # stages its half of THAT width)
    ("weight gradient first layer N=264 split-K 9", 1024, 264, MB, 1, 0, "split9"),
    ("forward ragged N=400 (K-major B, second tile 160 wide)", MB, 400, 512, 0, 1, "bias_relu"),
    ("input gradient ragged N=456 (MN-major B, second tile 224 wide)", MB, 456, 512, 0, 0, None),
    ("forward ragged N=296 (second tile 64 wide)", MB, 296, 256, 0, 1, None),
]


@pytest.mark.parametrize("case", H3P_CASES, ids=[c[0] for c in H3P_CASES])
def test_gemm_f16x3_persistent_kernel_layouts_and_epilogues(case):
    name, M, N, K, ta, tb, epi = case
    g = torch.Generator(device="cuda").manual_seed(8)
    A = torch.randn((K, M) if ta else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tb else (K, N), device="cuda", generator=g) * 0.05
    if ta:      # a gradient-sized operand: far below fp16's range unscaled
        A = A * 1e-5
    tw_a, tw_b = _Twin(A), _Twin(B)
    Aop = (A.t() if ta else A).double()
    Bop = (B.t() if tb else B).double()
    ref = Aop @ Bop
    if epi == "split9":
        S = 9
        slabs = torch.full((S, M, N), float("nan"), device="cuda")
        _gemm_h3(A, B, slabs, M, N, K, ta, tb, tw_a, tw_b, split=S)
        torch.cuda.synchronize()
        assert _last_kernel() in K_H3_PERSISTENT, name
        assert not bool(torch.isnan(slabs).any()), "every slab element must be written"
        assert _rel(slabs.double().sum(0), ref) <= TOL["f16x3"], name
        return
    out = torch.full((M, N), float("nan"), device="cuda")
    kw = {}
    if epi == "bias_relu":
        bias = torch.randn(N, device="cuda", generator=g)
        kw = dict(bias=bias, relu=1)
        ref = torch.relu(ref + bias.double())
    elif epi == "mask":
        h = torch.randn(M, N, device="cuda", generator=g)
        kw = dict(mask=h)
        ref = ref * (h > 0).double()
    elif epi == "accumulate":
        base = torch.randn(M, N, device="cuda", generator=g) * float(ref.abs().mean())
        out = base.clone()
        kw = dict(accumulate=1)
        ref = ref + base.double()
    slot = torch.zeros(2, device="cuda", dtype=torch.int32)
    _gemm_h3(A, B, out, M, N, K, ta, tb, tw_a, tw_b, c_amax=slot, **kw)
    torch.cuda.synchronize()
    assert _last_kernel() in K_H3_PERSISTENT, name
    e = _rel(out, ref)
    assert e <= TOL["f16x3"], "%s: rel err %.3e" % (name, e)
    assert int(slot[1].item()) == int(out.abs().max().view(torch.int32).item()), "max|C| word left by the epilogue"
    # twins reused (ready = 1): bit-identical
    out2 = torch.full((M, N), float("nan"), device="cuda") if epi != "accumulate" else base.clone()
    _gemm_h3(A, B, out2, M, N, K, ta, tb, tw_a, tw_b, ready=(1, 1), **kw)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("M,planes_only", [(MB, False), (MB + 1, False), (MB + 1, True), (5000, True)])
def test_gemm_f16x3_relu_bit_planes(M, planes_only):
    """ReLU masks as bit planes: a forward layer (bias + ReLU, fp32 or planes-only output) leaves one bit per output
    element behind; the input-gradient layer masked by those bits equals the layer masked by the fp32 tensor itself
    (bit for bit) and float64 masked by (forward output > 0).  Ragged row counts exercise the edge path of both sides."""
    from add_gym_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(11)
    K, N = 512, 1024
    X = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    tX, tW = _Twin(X), _Twin(W)
    h = torch.full((M, N), float("nan"), device="cuda")
    bits = torch.full((M, N // 32), -1, device="cuda", dtype=torch.int32)
    kw = {}
    if planes_only:
        tH = _Twin(h)
        slot = C.c_void_p(tH.amax.data_ptr())
        for _ in range(2):      # second round: the sticky scale is in force and the epilogue's planes stand
            _lib.check(_lib.lib().addk_f16x3_prep(_lib.stream(), slot, C.c_int(1)), "prep")
            _gemm_h3(X, W, h, M, N, K, 0, 1, tX, tW, bias=bias, relu=1, c_amax=tH.amax, c16=tH, bits_out=bits, no_f32=1)
    else:
        _gemm_h3(X, W, h, M, N, K, 0, 1, tX, tW, bias=bias, relu=1, bits_out=bits)
    torch.cuda.synchronize()
    assert _last_kernel() in K_H3_PERSISTENT
    href = torch.relu(X.double() @ W.double().t() + bias.double())
    if planes_only:
        assert bool(torch.isnan(h).all()), "planes-only: the fp32 output must not be written"
        s = 2.0 ** (14 - (((int(tH.amax[0].item()) >> 23) & 0xFF) - 127))
        hval = (tH.planes[:M * N].view(M, N).double() + tH.planes[tH.n:tH.n + M * N].view(M, N).double()) / s
        assert _rel(hval, href) <= TOL["f16x3"]
    else:
        hval = h.double()
    # the bit plane against the kernel's own output: bit set <=> element > 0 (elements within rounding of zero excluded
    # when the comparison is against float64)
    # layout (include/addk.h): column 64 g + 4 i + k of a row -> word 2 g + (k >> 1), bit 16 (k & 1) + i
    col = torch.arange(N, device="cuda")
    gi, ii, kk = col // 64, (col % 64) // 4, col % 4
    word = bits[:, (2 * gi + (kk >> 1))]
    got = ((word >> (16 * (kk & 1) + ii).to(torch.int32)) & 1).bool()
    if not planes_only:
        assert torch.equal(got, h > 0)
    clear = href.abs() > 1e-4
    assert torch.equal(got[clear], (href > 0)[clear])
    # consumer: dX = (dY . W2) * mask with W2 [N2 = 256... here the mask has N columns, so the product must be [M, N]
    K2 = 512
    dY = torch.randn(M, K2, device="cuda", generator=g) * 1e-5
    W2 = torch.randn(K2, N, device="cuda", generator=g) * 0.05
    tdY, tW2 = _Twin(dY), _Twin(W2)
    out_bits = torch.full((M, N), float("nan"), device="cuda")
    nblk = (M + 31) // 32
    part = torch.full((nblk + 2, N), float("nan"), device="cuda")
    _gemm_h3(dY, W2, out_bits, M, N, K2, 0, 0, tdY, tW2, bits_in=bits, colpart=part)
    torch.cuda.synchronize()
    assert _last_kernel() in K_H3_PERSISTENT
    ref = (dY.double() @ W2.double()) * got.double()
    assert _rel(out_bits, ref) <= TOL["f16x3"]
    # column sums per 32-row block, left behind by the same epilogue (the bias gradient of a gradient tensor)
    pad = torch.zeros(nblk * 32, N, device="cuda", dtype=torch.float64)
    pad[:M] = out_bits.double()
    blocks = pad.view(nblk, 32, N).sum(1)
    assert not bool(torch.isnan(part[:nblk]).any()) and bool(torch.isnan(part[nblk:]).all())
    assert float((part[:nblk].double() - blocks).abs().max()) <= 1e-5 * float(blocks.abs().max())
    assert _rel(part[:nblk].double().sum(0), out_bits.double().sum(0)) <= 1e-6
    maskf = got.float()                                     # the same mask as an fp32 tensor through the old path
    out_mask = torch.full((M, N), float("nan"), device="cuda")
    _gemm_h3(dY, W2, out_mask, M, N, K2, 0, 0, tdY, tW2, ready=(1, 1), mask=maskf)
    assert torch.equal(out_bits, out_mask)


def test_gemm_small_shapes_run_the_one_tile_kernel():
    """The dispatch rule the parity tests rely on: <= 37 wide tiles -> 64-wide one-tile-per-CTA kernel."""
    g = torch.Generator(device="cuda").manual_seed(9)
    A = torch.randn(256, 1024, device="cuda", generator=g)
    B = torch.randn(1024, 1024, device="cuda", generator=g)
    out = torch.empty(256, 1024, device="cuda")
    _gemm_h3(A, B, out, 256, 1024, 1024, 0, 1, _Twin(A), _Twin(B))
    assert _last_kernel() == K_H3_TILE
    A = torch.randn(8192, 1024, device="cuda", generator=g)
    out = torch.empty(8192, 1024, device="cuda")
    _gemm_h3(A, B, out, 8192, 1024, 1024, 0, 1, _Twin(A), _Twin(B))
    assert _last_kernel() in K_H3_PERSISTENT
