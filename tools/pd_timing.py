"""Diagnostic: per-deal device time of the trick estimator (BMC_PD_TIMING=1) on the N1 sample.

    python tools/pd_timing.py [n] ["<first-pass arena KB>[:0]" ...]      (":0" = submission order, no probe pass)

One JSON line (batch time, deals/s, sum / mean / p99 / max of the per-deal SM cycles; with a -DPD_PROFILE build named by
PD_TIMING_LIB also the cycles per phase) and gpurun_out/pd_timing_<arena>.npz (records + cycles: the training data of
tools/fit_playdeal_order.py)."""
import json
import os
import sys

os.environ["BMC_PD_TIMING"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bridge_mc_b200 import _lib, bidding  # noqa: E402
if os.environ.get("PD_TIMING_LIB"):                              # a -DPD_PROFILE build kept beside the product
    _lib.LIB_PATH = os.environ["PD_TIMING_LIB"]
from bridge_mc_b200.engine import Constraints, Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
eng = Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
rec, tal, stats = eng.sample(cons, n_raw=int(n / 0.038133), seed=11)
rec = np.ascontiguousarray(rec[np.lexsort((rec[:, 0], rec[:, 1]))])
n = rec.shape[0]
d_rec = torch.from_numpy(rec.view(np.int64).copy()).cuda()
d_out = torch.empty(n * 72, dtype=torch.uint8, device="cuda")
d_t = torch.tensor([-1], dtype=torch.int8, device="cuda")
d_d = torch.tensor([2], dtype=torch.uint8, device="cuda")
eng.play_deal_dev(d_rec.data_ptr(), min(n, 2048), d_t.data_ptr(), d_d.data_ptr(), False, d_out.data_ptr())
if hasattr(_lib.load(), "bmc_pd_debug_profile"):
    import ctypes
    _lib.load().bmc_pd_debug_profile((ctypes.c_ulonglong * 8)())          # clear what the warm-up added
res = {"deals": n}
for kb in (sys.argv[2:] or ["1024", "1024:0"]):                 # "<first-pass arena KB>[:0]", :0 = submission order (no probe pass)
    os.environ["BMC_PD_ARENA_KB"] = kb.split(":")[0]
    os.environ["BMC_PD_ORDER"] = "0" if kb.endswith(":0") else "1"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.play_deal_dev(d_rec.data_ptr(), n, d_t.data_ptr(), d_d.data_ptr(), False, d_out.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    out = d_out.cpu().numpy().view(_lib.PLAY_DTYPE)
    cyc = (out["reserved"].astype(np.uint32) * np.array([1, 256, 65536], dtype=np.uint32)).sum(axis=1).astype(np.float64) * 4096
    res[f"arena_{kb}kb"] = {"ms": e0.elapsed_time(e1), "deals_per_s": n / e0.elapsed_time(e1) * 1e3, "status": np.bincount(out["status"], minlength=5).tolist(),
                            "sum_cycles": float(cyc.sum()), "max_cycles": float(cyc.max()), "mean_cycles": float(cyc.mean()),
                            "p99_cycles": float(np.percentile(cyc, 99))}
    L = _lib.load()
    if hasattr(L, "bmc_pd_debug_profile"):                      # a -DPD_PROFILE build: cycles per phase
        import ctypes
        buf = (ctypes.c_ulonglong * 8)()
        L.bmc_pd_debug_profile(buf)
        res[f"arena_{kb}kb"]["phase_cycles"] = dict(zip(["setup", "play_strategy", "mk_route", "simple_trick", "after_tricks", "at_normalise", "at_first_pass", "at_rest"], [int(x) for x in buf[:8]]))
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"pd_timing_{kb.replace(':', '_')}.npz"), rec=rec, cyc=cyc, status=out["status"], tricks=out["score"][:, 1])
print(json.dumps(res))
