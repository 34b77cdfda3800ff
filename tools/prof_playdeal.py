"""Throughput of the trick estimator: n random full deals (sampled on the device), play-deal with no trump and
with a trump suit, device buffers, kernel time by CUDA events.  python tools/prof_playdeal.py [n] [arena_kb]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bridge_mc_b200 import _lib  # noqa: E402
from bridge_mc_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
arena_kb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
eng = Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
rec, _, _ = eng.sample(None, n_raw=n + n // 64, seed=7)
rec = rec[:n]
d_rec = torch.from_numpy(rec.view(np.int64).copy()).cuda()
d_out = torch.empty(n * 72, dtype=torch.uint8, device="cuda")
for trump in (-1, 3):
    d_t = torch.tensor([trump], dtype=torch.int8, device="cuda")
    d_d = torch.tensor([2], dtype=torch.uint8, device="cuda")
    eng.play_deal_dev(d_rec.data_ptr(), min(n, 4096), d_t.data_ptr(), d_d.data_ptr(), False, d_out.data_ptr(), arena_kb)   # warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.play_deal_dev(d_rec.data_ptr(), n, d_t.data_ptr(), d_d.data_ptr(), False, d_out.data_ptr(), arena_kb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = d_out.cpu().numpy().view(_lib.PLAY_DTYPE)
    t0 = time.perf_counter()
    host = eng.play_deal(rec, trump, 2, arena_kb)
    e2e = time.perf_counter() - t0
    assert host.tobytes() == out.tobytes()
    st = np.bincount(out["status"], minlength=5)
    print(json.dumps({"config": "play-deal", "trump": trump, "deals": n, "kernel_ms": ms, "deals_per_s": n / ms * 1e3,
                      "e2e_host_buffers_deals_per_s": n / e2e, "status_counts": st.tolist(),
                      "mean_declarer_tricks": float(out["score"][:, 1][out["status"] == 0].mean())}))
