"""GPU box: fill most of the HBM with NaN bit patterns and exit, so that the next process gets this memory back un-zeroed
(cudaMalloc does not clear).  Run before the GPU tests / bench.py to catch reads of uninitialised device memory:
    python tools/dirty_hbm.py && python -m pytest tests -m gpu -x -q && python tools/dirty_hbm.py && python bench.py"""
import torch
free, total = torch.cuda.mem_get_info()
n = int(free * 0.92) // 4
x = torch.full((n,), float("nan"), device="cuda", dtype=torch.float32)
torch.cuda.synchronize()
print("dirtied %.1f GB of %.1f" % (n * 4 / 1e9, total / 1e9))
