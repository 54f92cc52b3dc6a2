"""Does a copy-engine D2H (3.2 GB) overlap with the GRM compute?  (decides how bound outputs leave the device)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissect_b200 import GrmEngine
from dissect_b200.engine import PinnedArray
from dissect_b200.synth import synth_packed_cuda

n, m = 20000, 98304
pk, _ = synth_packed_cuda(n, m, 2, missing=0.01)
host_pk = PinnedArray(tuple(pk.shape), np.uint8); host_pk.array[:] = pk.cpu().numpy()
bpr = pk.shape[1]
eng = GrmEngine(); eng.set_individuals(n)
dev = torch.zeros(400_000_000, dtype=torch.float64, device="cuda")      # 3.2 GB
host = torch.empty(400_000_000, dtype=torch.float64, pin_memory=True)
side = torch.cuda.Stream()

def compute():
    eng.begin(m); eng.push_bed_host(host_pk.ptr, pitch=bpr, nrows=m); eng.finish()

for label, do_copy, do_compute in (("compute only", False, True), ("copy only", True, False), ("both", True, True), ("both", True, True)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if do_copy:
        with torch.cuda.stream(side):
            for i in range(8):
                host[i*50_000_000:(i+1)*50_000_000].copy_(dev[i*50_000_000:(i+1)*50_000_000], non_blocking=True)
    if do_compute:
        compute()
    torch.cuda.synchronize()
    print(f"{label:14s} {1e3*(time.perf_counter()-t0):8.2f} ms")
