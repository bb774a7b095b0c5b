"""PCIe copy rates of the shapes the e2e path uses (one GPU): contiguous H2D / D2H, and the strided D2H of a k-chunk's
columns into the (n_f, n_k, 3) complex64 result (psa_copy_rows = cudaMemcpy2DAsync).  Prints GB/s."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from psa_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
n_t, n_k, elem = 16384, 10000, 24
host = torch.empty(n_t * n_k * elem, dtype=torch.uint8, pin_memory=True)
host.zero_()
dbuf = torch.empty(n_t * n_k * elem, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


gb = n_t * n_k * elem / 1e9
print(f"contiguous H2D {gb / timed(lambda: dbuf.copy_(host, non_blocking=True)) * 1e3:.1f} GB/s")
print(f"contiguous D2H {gb / timed(lambda: host.copy_(dbuf, non_blocking=True)) * 1e3:.1f} GB/s")
for pieces in (16, 64):
    n = host.numel() // pieces
    print(f"H2D in {pieces} pieces {gb / timed(lambda: [dbuf[i * n:(i + 1) * n].copy_(host[i * n:(i + 1) * n], non_blocking=True) for i in range(pieces)]) * 1e3:.1f} GB/s")
for nk in (250, 500, 1000, 2000, 5000, 10000):
    def go():
        for k0 in range(0, n_k, nk):
            _lib.call("psa_copy_rows", host.data_ptr() + k0 * elem, n_k * elem, dbuf.data_ptr(), nk * elem, nk * elem, n_t,
                      stream.cuda_stream)
    print(f"strided D2H, {nk:5d} k per chunk ({nk * elem / 1024:.0f} KiB rows): {gb / timed(go) * 1e3:.1f} GB/s")
# both directions at once
side = torch.cuda.Stream(device=dev)
h2 = torch.empty(host.numel(), dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(host.numel(), dtype=torch.uint8, device=dev)


def duplex():
    side.wait_stream(stream)
    with torch.cuda.stream(side):
        d2.copy_(h2, non_blocking=True)
    host.copy_(dbuf, non_blocking=True)
    stream.wait_stream(side)


print(f"duplex (H2D + D2H at once): {2 * gb / timed(duplex) * 1e3:.1f} GB/s total")
