import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from acrobotics_b200 import _native
lib = _native.load()
L, n = 10000, int(os.environ.get("DP_N", "842"))
rng = np.random.default_rng(0)
q = torch.from_numpy(rng.uniform(-3, 3, (L * n, 6))).cuda()
off = (np.arange(L + 1) * n).astype(np.int64)
M = L * n
value = torch.empty(M, dtype=torch.float64, device="cuda"); pred = torch.empty(M, dtype=torch.int32, device="cuda")
off_dev = torch.empty(L + 1, dtype=torch.int64, device="cuda"); rows = torch.empty(L, dtype=torch.int64, device="cuda"); length = torch.empty(1, dtype=torch.float64, device="cuda")
def call():
    _native.check(lib.acro_shortest_path_f64(q.data_ptr(), 6, off.ctypes.data, L, 2, None, None, value.data_ptr(), pred.data_ptr(), off_dev.data_ptr(), rows.data_ptr(), length.data_ptr(), None))
call(); torch.cuda.synchronize()
t0 = time.perf_counter(); call(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue %.1f ms, total %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
