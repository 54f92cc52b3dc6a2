"""e2e probe: where does the host-buffer path spend its time?  (H2D + compute) vs bound (streamed) outputs
vs fetch-after-finish, at the bench geometry."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissect_b200 import GrmEngine  # noqa: E402
from dissect_b200.engine import PinnedArray  # noqa: E402
from dissect_b200.synth import synth_packed_cuda  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--m", type=int, default=98304)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    n, m = a.n, a.m
    pk, _ = synth_packed_cuda(n, m, 2, missing=0.01)
    host_pk = PinnedArray(tuple(pk.shape), np.uint8)
    host_pk.array[:] = pk.cpu().numpy()
    bpr = pk.shape[1]
    A = PinnedArray((n, n), np.float64, order="F")
    N = PinnedArray((n, n), np.float64, order="F")
    eng = GrmEngine(profile=True)
    eng.set_individuals(n)

    def run(label, bind, fetch):
        eng.unbind_outputs()
        if bind:
            eng.bind_output("A", A.ptr, n)
            eng.bind_output("N", N.ptr, n)
        ts = []
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.begin(m)
            eng.push_bed_host(host_pk.ptr, pitch=bpr, nrows=m)
            t1 = time.perf_counter()
            eng.finish()
            t2 = time.perf_counter()
            if fetch:
                eng.fetch_into("A", A.ptr, n)
                eng.fetch_into("N", N.ptr, n)
            t3 = time.perf_counter()
            ts.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
        best = min(ts, key=lambda x: x[3])
        kt = eng.kernel_times()
        print(f"{label:28s} push {best[0]*1e3:7.2f}  finish {best[1]*1e3:7.2f}  fetch {best[2]*1e3:7.2f}  total "
              f"{best[3]*1e3:7.2f} ms   fetch-class ms/rep {kt['fetch'][0]/a.reps:.2f} launches {kt['fetch'][1]}  "
              f"d2h ms/rep {kt['d2h'][0]/a.reps:.2f} copies {kt['d2h'][1]}")

    run("compute only (H2D in)", False, False)
    run("bound outputs (streamed)", True, False)
    run("fetch after finish", False, True)
    run("bound outputs (streamed)", True, False)
    eng.close()


if __name__ == "__main__":
    main()
