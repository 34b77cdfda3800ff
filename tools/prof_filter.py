"""K2 timings: filter_kernel with features only, with choose-bid interpreted, and with the table compiled (NVRTC).
Device buffers (bmc_filter_dev), CUDA-event times, 2^22 deals.  python tools/prof_filter.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bridge_mc_b200 import _lib, bidding
from bridge_mc_b200.engine import Engine, Scheme

n = 1 << 22
eng = Engine(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
rec = torch.empty((n, 2), dtype=torch.int64, device="cuda")
tal = torch.zeros(_lib.TALLY_WORDS, dtype=torch.int64, device="cuda")
eng.sample_dev(None, rec.data_ptr(), n, tal.data_ptr(), n_raw=n + n // 256, tally_mask=0)
feats = torch.empty((n, 48), dtype=torch.uint8, device="cuda")
acc = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, scheme in (("features only", None), ("choose-bid openings, interpreted", Scheme(bidding.openings.sexpr()).set_jit(0)),
                     ("choose-bid openings, compiled", Scheme(bidding.openings.sexpr()).set_jit(1)),
                     ("choose-bid (nt-responses 1), compiled", Scheme(bidding.nt_responses(1).sexpr()).set_jit(1))):
    for _ in range(2):
        eng.filter_dev(rec.data_ptr(), n, None, scheme, acc.data_ptr(), feats.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        eng.filter_dev(rec.data_ptr(), n, None, scheme, acc.data_ptr(), feats.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    gbs = n * 65 / us / 1e3
    print(f"{name:45s} {us:9.1f} us  {gbs:8.1f} GB/s algorithmic (16 B in + 49 B out per deal)  {100 * gbs / 6553.9:5.1f} % of measured HBM copy")
