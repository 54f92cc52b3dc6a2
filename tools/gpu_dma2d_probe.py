"""Copy-engine D2H rate for strided (2-D) copies vs contiguous ones (decides how streamed outputs are cut)."""
import time
import torch
from cuda import cudart

n = 20000
dev = torch.zeros(n * 2048, dtype=torch.float64, device="cuda")
err, host = cudart.cudaHostAlloc(n * n * 8, 0)
assert err == 0
stream = torch.cuda.Stream()
s = stream.cuda_stream


def timed(label, fn, nbytes):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label:46s} {dt*1e3:8.2f} ms  {nbytes/dt/1e9:6.1f} GB/s")


for rows in (256, 1024, 2048, 4096, 10000, 20000):
    width = rows * 8
    h = min(n, dev.numel() // rows)
    def f():
        (e,) = cudart.cudaMemcpy2DAsync(host, n * 8, dev.data_ptr(), width, width, h, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost, s)
        assert e == 0
    timed(f"2D: segment {width//1024:4d} KB x {h} cols (dst pitch 156 KB)", f, width * h)
    timed(f"2D: segment {width//1024:4d} KB x {h} cols (dst pitch 156 KB)", f, width * h)

def g():
    (e,) = cudart.cudaMemcpyAsync(host, dev.data_ptr(), dev.numel() * 8, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost, s)
    assert e == 0
timed("contiguous 328 MB", g, dev.numel() * 8)
timed("contiguous 328 MB", g, dev.numel() * 8)
