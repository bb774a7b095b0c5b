import sys, time, ctypes, os
import numpy as np
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
from psa_b200 import _lib
lib = _lib.load()
n_fr, n_at = 100, 64000
rng = np.random.default_rng(0)
frames = (rng.random((n_fr, n_at, 3)) * 108).astype(np.float32)
types = np.ones(n_at, np.int32)
box = np.diag([108.6, 108.6, 108.6]).astype(np.float32)
path = b'/dev/shm/dump_bench.dump'
for thr in (1, 2, 4, 8, 1):
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        rc = lib.psa_write_dump(path, frames.ctypes.data, types.ctypes.data, n_fr, n_at, box.ctypes.data, thr)
        best = min(best, time.perf_counter() - t)
    sz = os.path.getsize(path)
    print(f"threads {thr}: {best:.3f} s  {n_fr*n_at/best/1e6:.2f} M lines/s  {sz/best/1e6:.0f} MB/s  ({sz/1e6:.0f} MB)")
import hashlib
print(hashlib.md5(open(path,'rb').read()).hexdigest())
