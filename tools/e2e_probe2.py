import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridge_mc_b200 import bidding, _lib
from bridge_mc_b200.engine import Constraints, Engine
eng = Engine(0)
cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
cap = 110_000_000
dev = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
src = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
tal = torch.zeros(_lib.TALLY_WORDS, dtype=torch.int64, device="cuda")
tp = torch.empty((cap, 2), dtype=torch.int64, pin_memory=True)
s2 = torch.cuda.Stream()
def job():
    for i in range(3):
        st = eng.sample_dev(cons, dev.data_ptr(), cap, tal.data_ptr(), n_raw=10 << 28, first_index=(i + 1) << 40)
    print("sampler kernel_ms (3 x 10 x 2^28)", st.kernel_ms)
eng.sample_dev(cons, dev.data_ptr(), cap, tal.data_ptr(), n_raw=1 << 20)
torch.cuda.synchronize()
t = threading.Thread(target=job); t.start()
time.sleep(0.01)
SRC = dev if len(sys.argv) > 1 else src
with torch.cuda.stream(s2):
    t0 = time.perf_counter()
    tp.copy_(SRC, non_blocking=True); s2.synchronize()
    dt = time.perf_counter() - t0
print(f"1.76 GB D2H concurrent with the sampler: {cap*16/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms)")
t.join()
with torch.cuda.stream(s2):
    t0 = time.perf_counter()
    tp.copy_(src, non_blocking=True); s2.synchronize()
    dt = time.perf_counter() - t0
print(f"1.76 GB D2H alone: {cap*16/dt/1e9:.1f} GB/s")
