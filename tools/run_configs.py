"""Run the five BASELINE.json configurations and print one JSON object per config.

    python tools/run_configs.py [c1 c2 c3 c4 c5] [--scale 1.0]
    python -m torch.distributed.run --nproc-per-node N ... tools/run_configs.py c3 c5

Under torchrun the raw index range of a config is split evenly over the ranks
(disjoint Philox ranges) and the tallies are all-reduced with NCCL; rank 0 prints.
Times are CUDA-event times of the sampling kernels, max over ranks.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from bridge_mc_b200 import _lib, bidding, compscore  # noqa: F401
from bridge_mc_b200 import dist as bdist
from bridge_mc_b200.engine import Constraints, Engine, tallies_from_words

SEED = 0x6272696467656D63


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scale = 1.0
    use_jit = "--jit" in sys.argv
    for i, a in enumerate(sys.argv):
        if a == "--scale":
            scale = float(sys.argv[i + 1])
    which = [a for a in args if a.startswith("c") or a == "ref"] or ["c1", "c2", "c3", "c4", "c5", "ref"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    tallies = torch.zeros(_lib.TALLY_WORDS, dtype=torch.int64, device=dev)

    def run(name, cons, n_raw, note, tally_mask=7, records=False):
        if use_jit and cons is not None:
            cons.set_jit(True)
        first, per = bdist.rank_range(0, n_raw, rank, world)
        cap = 0
        rec = None
        if records:
            cap = per
            rec = torch.empty((cap, 2), dtype=torch.int64, device=dev)
        tallies.zero_()
        eng.sample_dev(cons, rec.data_ptr() if records else 0, cap, tallies.data_ptr(), n_raw=min(per, 1 << 20),
                       first_index=1 << 50, seed=SEED, tally_mask=tally_mask)            # warm-up + calibration
        tallies.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = eng.sample_dev(cons, rec.data_ptr() if records else 0, cap, tallies.data_ptr(), n_raw=per, first_index=first,
                            seed=SEED, tally_mask=tally_mask)
        bdist.allreduce_tallies(tallies)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = torch.tensor([st.kernel_ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        t = tallies_from_words(tallies.cpu().numpy().view(np.uint64))
        if rank == 0:
            print(json.dumps({
                "config": name, "note": note, "n_gpus": world, "mode": (cons.mode if cons else 1),
                "jit_compile_ms": (cons.jit_ms if cons else 0.0),
                "raw": t["raw"], "voided": t["voided"], "accepted": t["accepted"], "acceptance": t["accepted"] / max(1, t["raw"]),
                "kernel_ms": float(ms[0]), "wall_ms_incl_allreduce": float(ms[1]),
                "raw_deals_per_s": t["raw"] / (float(ms[0]) * 1e-3), "accepted_per_s": t["accepted"] / (float(ms[0]) * 1e-3),
                "hcp_seat2_15_17": [int(x) for x in t["hcp"][2][15:18]],
                "primary_seat_exact_acceptance": (cons.primary_acceptance[0] / 635013559600) if (cons is not None and cons.mode == 4) else None,
                "tally_checksum": int((t["hcp"] * np.arange(1, 153).reshape(4, 38)).sum() + 3 * t["len"].sum() + 7 * t["shape"].sum()) % (1 << 61),
            }))
        return t

    p = bidding.openings.passes()
    if "c1" in which:
        run("C1", None, 1_000_000, "1M unconstrained deals, HCP + suit-length histograms, records stored", tally_mask=3,
            records=True)
        run("C1-1e9", None, int(1e9 * scale), "unconstrained, all tallies, no records")
    if "c2" in which:
        cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
        run("C2", cons, int(2_684_354_560 * scale), "South test-bid (1 NT); 10 x 2^28 raw = 1.02e8 accepted; records stored",
            records=True)
    if "c3" in which:
        nt = bidding.nt_responses(1)
        cons = Constraints(None, [f"(and {p} {nt.chooses((3, 'NT'))})", p, bidding.openings.chooses((1, "NT")), p])
        run("C3", cons, int(1e10 * scale), "N E W pass, S chooses 1NT, N answers 3NT over (nt-responses 1)")
    if "c5" in which:
        nt2 = bidding.nt_responses(2)
        cons = Constraints(None, [nt2.chooses((6, "NT")), None, bidding.openings.chooses((2, "NT")), None])
        run("C5", cons, int(1e11 * scale), "S chooses 2NT, N answers 6NT over (nt-responses 2)")
    if "ref" in which:
        # the REST shape (rest.cl:264-277): one fixed hand + three ranged seats, sampled in the reference's own
        # seat-by-seat order (bmc_constraints_set_reference_order)
        from bridge_mc_b200.cards import str2hand
        from bridge_mc_b200.engine import seat_spec
        south = str2hand("♠ AKJ5 ♥ K1064 ♦ Q87 ♣ A6")
        specs = [seat_spec(fixed=south), seat_spec(hcp=(10, 12), s=(4, None)), seat_spec(), seat_spec()]
        cons = Constraints(specs).set_reference_order(3)
        run("REF-order", cons, int(2e8 * scale), "fixed hand + 3 ranged seats in `build`'s sequential order (mode 3)")
        cons2 = Constraints([seat_spec(hcp=(15, 17), s=(5, None)), seat_spec(hcp=(10, 11), h=(4, None))]).set_reference_order(2)
        run("REF-order-2", cons2, int(2e8 * scale), "two ranged seats (15-17 with 5+ spades, then 10-11 with 4+ hearts), mode 3")
    if "c4" in which:
        # compscore tallies over CONSTRAINED samples: C2's sampler scores every accepted deal in the launch that draws it
        # (BMC_TALLY_SCORE): 8 contract x vulnerability rows + 6 IMP pairs; the index range is fixed (8 x 10 x 2^28 raw,
        # ~8.2e8 accepted) and split over the ranks, so every tally is identical for every number of GPUs
        from bridge_mc_b200._lib import TALLY_ALL, TALLY_SCORE
        cs = [compscore.Contract(t).c_struct() for t in ("1NTS", "3NTS", "4HS", "4SS")] * 2
        cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
        cons.set_scoring(cs, [0, 0, 0, 0, 1, 1, 1, 1], [0, 1, 1, 2, 2, 3, 4, 5, 5, 6, 6, 7])
        t = run("C4", cons, int(8 * 10 * (1 << 28) * scale), "South test-bid (1 NT) sample scored in stage B: tricks x base-score per "
                "contract, match-points per pair (stand-in trick source keyed on the deal; real tricks: tools/run_playdeal.py)",
                tally_mask=TALLY_ALL | TALLY_SCORE)
        if rank == 0:
            print(json.dumps({"config": "C4-tallies", "n_gpus": world, "accepted": t["accepted"],
                              "tricks_row0": [int(x) for x in t["tricks"][0]], "imps_pair0_nonzero": int((t["imps"][0] > 0).sum()),
                              "imp_dropped": [int(x) for x in t["imp_dropped"]],
                              "score_checksum": int((t["tricks"] * np.arange(1, 113).reshape(8, 14)).sum() + 3 * (t["imps"] * np.arange(1, 393).reshape(8, 49)).sum()) % (1 << 61)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
