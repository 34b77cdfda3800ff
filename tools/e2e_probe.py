import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridge_mc_b200 import bidding, _lib
from bridge_mc_b200.engine import Constraints, Engine
eng = Engine(0)
cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
L = _lib.load()
cap = 110_000_000
hp = L.bmc_host_alloc(cap * 16)
host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint64)), shape=(cap, 2))
host[:] = 0
dev = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
th = torch.from_numpy(host.view(np.int64))
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th.copy_(dev, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"plain D2H into bmc_host_alloc buffer: {cap*16/dt/1e9:.1f} GB/s")
tp = torch.empty((cap, 2), dtype=torch.int64, pin_memory=True)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tp.copy_(dev, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"plain D2H into torch pinned buffer: {cap*16/dt/1e9:.1f} GB/s")
for name, buf in (("bmc_host_alloc", host), ("torch pinned", tp.numpy().view(np.uint64))):
    for i in range(3):
        rec, tal, st = eng.sample(cons, n_accept=100_000_000, wave=1 << 28, first_index=(i + 1) << 40, cap=cap, records=buf)
        print(f"{name}: waves {st.waves} kernel_ms {st.kernel_ms:.1f} total_ms {st.total_ms:.1f}")
