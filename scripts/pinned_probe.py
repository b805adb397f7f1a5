"""Diagnostic: host->device copy bandwidth of pinned buffers by size and by how they were pinned."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hector_slam_b200 import parallel
torch.cuda.init()
print("bound:", parallel.bind_process_to_gpu_numa_node(0) is not None)
def bw(h, label):
    d = torch.empty(h.numel(), dtype=h.dtype, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{label:40s} {h.numel()*4/1e6:7.1f} MB  {h.numel()*4/dt/1e9:6.1f} GB/s", flush=True)
for mb in (8, 16, 17.7, 24, 31, 32, 33, 35.4, 48, 64, 65, 128):
    n = int(mb * 1e6 / 4)
    bw(torch.empty(n, dtype=torch.float32).pin_memory(), f"pin_memory() {mb} MB")
for mb in (17.7, 35.4, 64):
    n = int(mb * 1e6 / 4)
    bw(torch.empty(n, dtype=torch.float32, pin_memory=True), f"empty(pin_memory=True) {mb} MB")
a = torch.empty(int(35.4e6 / 4), dtype=torch.float32)
r = torch.cuda.cudart().cudaHostRegister(a.data_ptr(), a.numel() * 4, 0)
bw(a, f"cudaHostRegister (rc={int(r)}) 35.4 MB")
# second allocation of the same size after the first is still alive
x1 = torch.empty(int(35.4e6 / 4), dtype=torch.float32).pin_memory(); x2 = torch.empty(int(35.4e6 / 4), dtype=torch.float32).pin_memory()
bw(x1, "35.4 MB first of two"); bw(x2, "35.4 MB second of two")
