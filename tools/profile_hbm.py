"""Profiling driver (GPU box): the HBM-bound kernels of the path one by one, through the C-ABI, against the measured
HBM peak.  Each kernel is launched back to back between one CUDA-event pair, rotating over operand sets whose total
exceeds the 126 MB L2, so `GB/s` = ALGORITHMIC bytes per launch / average launch time.
    python tools/profile_hbm.py [reps]"""
import ctypes as C
import json
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from add_gym_b200 import _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
L = _lib.lib()
dev = "cuda:0"
peak = 6446.9
try:
    peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows_out = []


def timed(name, nbytes, launch, nsets):
    try:
        for i in range(nsets):                     # warm-up: every operand set once
            launch(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            launch(i % nsets)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        gbs = nbytes / us / 1e3
        rows_out.append({"kernel": name, "us": us, "algorithmic_MB": nbytes / 1e6, "GBps": gbs, "frac": gbs / peak})
        print("%-52s %8.1f us  %8.1f MB  %7.0f GB/s  %5.1f %% of %.0f" % (name, us, nbytes / 1e6, gbs, 100 * gbs / peak, peak), flush=True)
    except Exception as e:                        # one kernel failing must not hide the others
        print("%-52s FAILED: %s" % (name, e), flush=True)


# AdamW on the flat parameter vector: 28 B / parameter (read p, g, m, v; write p, m, v)
P = 4349983
sets = [[torch.randn(P, device=dev) * s for s in (0.1, 0.01, 0.0, 0.0)] for _ in range(4)]
for s in sets:
    s[3].abs_()
timed("adamw_kernel (4,349,983 parameters)", 28 * P,
      lambda i: _lib.check(L.addk_adamw(_lib.stream(), _lib.ptr(sets[i][0]), _lib.ptr(sets[i][1]), _lib.ptr(sets[i][2]),
                                        _lib.ptr(sets[i][3]), C.c_longlong(P), C.c_int(1), C.c_double(1e-4), C.c_double(0.9),
                                        C.c_double(0.999), C.c_double(1e-8), C.c_double(0.0), C.c_double(1.0)), "addk_adamw"), 4)
del sets

# TD(lambda) returns + advantages: 24 B / sample (reward, next value, value, done; target, advantage)
for N in (4096, 32768):
    T = 32
    ns = 10 if N == 32768 else 48
    td = [(torch.randn(T, N, device=dev), torch.randn(T, N, device=dev), torch.randn(T, N, device=dev),
           (torch.rand(T, N, device=dev) < 0.01).int(), torch.empty(T, N, device=dev), torch.empty(T, N, device=dev)) for _ in range(ns)]
    timed("td_lambda_kernel (T=32, N=%d)" % N, 24 * T * N,
          lambda i: _lib.check(L.addk_td_lambda(_lib.stream(), _lib.ptr(td[i][0]), _lib.ptr(td[i][1]), _lib.ptr(td[i][2]),
                                                _lib.ptr(td[i][3]), C.c_int(T), C.c_int(N), C.c_float(0.99), C.c_float(0.95),
                                                C.c_float(0.0), C.c_float(0.0), _lib.ptr(td[i][4]), _lib.ptr(td[i][5])),
                               "addk_td_lambda"), ns)
    del td

# f16x3 pre-pass on a 16384 x 1024 operand: max|x| pass 4 B / element, split pass 4 B read + 4 B written / element
R, Cc = 16384, 1024
xs = [torch.randn(R, Cc, device=dev) for _ in range(3)]
planes = [torch.empty(2 * R * Cc, dtype=torch.int16, device=dev) for _ in range(3)]
slots = torch.zeros(3, 2, dtype=torch.int32, device=dev)
slot_ptr = lambda i: C.c_void_p(slots.data_ptr() + 8 * i)
conv = lambda i: _lib.check(L.addk_f16x3_convert(_lib.stream(), _lib.ptr(xs[i]), C.c_longlong(R), C.c_int(Cc), C.c_int(Cc),
                                                 _lib.ptr(planes[i]), C.c_longlong(R * Cc), slot_ptr(i)), "addk_f16x3_convert")
timed("h3_amax_kernel + h3_split_kernel (16384x1024)", 12 * R * Cc, conv, 3)
timed("h3_split_kernel alone (16384x1024)", 8 * R * Cc,
      lambda i: _lib.check(L.addk_f16x3_split(_lib.stream(), _lib.ptr(xs[i]), C.c_longlong(R), C.c_int(Cc), C.c_int(Cc),
                                              _lib.ptr(planes[i]), C.c_longlong(R * Cc), slot_ptr(i), None, None), "addk_f16x3_split"), 3)
del xs, planes

# normalizer column statistics over one iteration's observations: 4 B / element
n, dim = 131072, 264
obs = [torch.randn(n, dim, device=dev) for _ in range(2)]
out = torch.zeros(2 * dim, dtype=torch.float64, device=dev)
timed("column_stats (131072 x 264, sum + sum of squares)", 4 * n * dim,
      lambda i: _lib.check(L.addk_column_stats(_lib.stream(), _lib.ptr(obs[i]), None, C.c_longlong(n), C.c_int(dim), C.c_int(0),
                                               _lib.ptr(out)), "addk_column_stats"), 2)
del obs

# runtime motion lookup: per query 12 B in (id, time), one 284-byte table row out as six tensors + 284 B read
S, D, n = 906000, 29, 6 * 32768
table = torch.randn(S, 72, device=dev)
start = torch.zeros(1, dtype=torch.long, device=dev)
ids = torch.zeros(n, dtype=torch.long, device=dev)
times = [torch.rand(n, device=dev) * (S / 100.0) for _ in range(3)]
outs = [[torch.empty(n, w, device=dev) for w in (3, 4, 3, 3, D, D)] for _ in range(3)]
timed("motion_gather_kernel (196,608 lookups, 261 MB table)", n * (12 + 284 + 284),
      lambda i: _lib.check(L.addk_motion_gather(_lib.stream(), _lib.ptr(table), C.c_int(72), C.c_int(D), C.c_longlong(S),
                                                _lib.ptr(start), C.c_float(100.0), _lib.ptr(ids), _lib.ptr(times[i]), C.c_int(n),
                                                *[_lib.ptr(o) for o in outs[i]], None), "addk_motion_gather"), 3)

os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump({"hbm_peak_GBps": peak, "reps": reps, "kernels": rows_out}, open(os.path.join(REPO, "gpurun_out", "hbm_kernels.json"), "w"), indent=1)
