"""N1 measurement: the trick estimator behind the reference's compscore measures (rest.cl:287-303), end to end.

    python tools/run_playdeal.py [n_deals] [--no-cpu]

Pipeline ("the only thing between sampled deals and real compscore tallies"):
  1. sample n deals in which South opens 1NT (configuration C2: `(test-bid '(1 NT) hand)`), records stay on the device;
  2. `play-deal` in no trump with South as declarer on every deal (bmc_play_deal_dev, one warp per deal);
  3. `base-score` of "1NTS" and "3NTS", not vulnerable and vulnerable, with the estimated tricks (bmc_score);
  4. tallies: trick histogram, mean scores, how often 3NT makes.
Prints one JSON object: kernel times by CUDA events, deals/s, and beside it the same algorithm compiled for the host
(tests/playdeal_host_selftest.cpp, one core, a bounded sample).  `python bench.py --workload n1` runs the same pipeline
and adds the CPU restatement of the reference's list algorithm as `cpu_baseline` (only bench.py may run that)."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bridge_mc_b200 import _lib, bidding, cards  # noqa: E402
from bridge_mc_b200.engine import Constraints, Engine  # noqa: E402

def run(n=32768, with_host_build=True):
    """the N1 pipeline; returns (result dict, the records, the estimated declarer tricks)"""
    eng = Engine(0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    # a fixed index range, every accepted deal of it, in a canonical order: the same deal set on every run
    rec, tal, stats = eng.sample(cons, n_raw=int(n / 0.038133), seed=11)
    rec = np.ascontiguousarray(rec[np.lexsort((rec[:, 0], rec[:, 1]))])
    n = rec.shape[0]
    d_rec = torch.from_numpy(rec.view(np.int64).copy()).cuda()
    d_out = torch.empty(n * 72, dtype=torch.uint8, device="cuda")
    d_t = torch.tensor([-1], dtype=torch.int8, device="cuda")
    d_d = torch.tensor([2], dtype=torch.uint8, device="cuda")
    eng.play_deal_dev(d_rec.data_ptr(), min(n, 2048), d_t.data_ptr(), d_d.data_ptr(), False, d_out.data_ptr())        # warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.play_deal_dev(d_rec.data_ptr(), n, d_t.data_ptr(), d_d.data_ptr(), False, d_out.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    kernel_ms = e0.elapsed_time(e1)
    out = d_out.cpu().numpy().view(_lib.PLAY_DTYPE)
    ok = out["status"] == 0
    tricks = out["score"][:, 1].astype(np.uint8)
    contracts = [_lib.Contract(1, 4, 2, 0), _lib.Contract(1, 4, 2, 0), _lib.Contract(3, 4, 2, 0), _lib.Contract(3, 4, 2, 0)]
    t0 = time.perf_counter()
    scores = eng.score(contracts, [0, 1, 0, 1], np.repeat(tricks[ok][:, None], 4, axis=1))
    score_ms = (time.perf_counter() - t0) * 1e3
    res = {"config": "N1: C2 sample -> play-deal (NT, South declares) -> base-score of 1NTS / 3NTS", "deals": n,
           "status_counts": np.bincount(out["status"], minlength=5).tolist(),
           "play_deal_kernel_ms": kernel_ms, "deals_per_s": n / kernel_ms * 1e3, "score_ms_host_buffers": score_ms,
           "tricks_hist": np.bincount(tricks[ok], minlength=14).tolist(), "mean_tricks": float(tricks[ok].mean()),
           "p_3nt_makes": float((tricks[ok] >= 9).mean()),
           "mean_score": {"1NTS": float(scores[:, 0].mean()), "1NTS_vul": float(scores[:, 1].mean()),
                          "3NTS": float(scores[:, 2].mean()), "3NTS_vul": float(scores[:, 3].mean())}}

    if with_host_build:
        # (a) the same algorithm on one host core
        exe = os.path.join(tempfile.mkdtemp(), "pd_host")
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "playdeal_host_selftest.cpp")], check=True)
        m = min(n, 1500)
        masks = cards.seat_masks(rec[:m])                                   # N E S W lane masks; declarer South: S W N E
        lines = []
        for i in range(m):
            hs = [int(masks[i, (2 + k) % 4]) for k in range(4)]
            lines.append("-1 " + " ".join(str((h >> (16 * s)) & 0x1FFF) for h in hs for s in range(4)))
        t0 = time.perf_counter()
        txt = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.strip().split("\n")
        host_s = time.perf_counter() - t0
        host_tricks = np.array([int(line.split()[2]) for line in txt], dtype=np.uint8)
        assert (host_tricks == tricks[:m]).all(), "the host build of playdeal.cuh disagrees with the GPU"
        res["cpu_same_algorithm_one_core"] = {"deals": m, "deals_per_s": m / host_s, "kind": "playdeal.cuh compiled by g++ -O2 (test harness)"}
    return res, rec, tricks


if __name__ == "__main__":
    n_arg = int(next((a for a in sys.argv[1:] if not a.startswith("--")), 32768))
    print(json.dumps(run(n_arg, "--no-cpu" not in sys.argv)[0]))
