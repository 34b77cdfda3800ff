"""The N>1 path on CPU: world_size-2 gloo processes exercise the product's partition + merge
plumbing (bridge_mc_b200/dist.py).  The per-rank sampler is stood in for by the oracle's CPU
mirror of the GPU sampler (there is no GPU here), which is exactly what makes the check
meaningful: the merged tallies must equal the single-rank tallies bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bridge_mc_b200 import dist as bdist
from bridge_mc_b200._lib import TALLY_WORDS

SEED, FIRST, N = 0x6272696467656D63, 12345, 20011          # N odd on purpose: uneven shares


def cpu_tallies(first, n):
    """tally words (raw, voided, accepted, stored, hcp[4][40], len[4][4][14], shape..) from the CPU mirror"""
    from oracle import oracle as O
    from bridge_mc_b200.cards import seat_masks
    rec, void = O.gpu_deal_batch(SEED, first, n)
    rec = rec[void == 0]
    t = np.zeros(TALLY_WORDS, dtype=np.int64)
    t[0], t[1], t[2] = n, int(void.sum()), rec.shape[0]
    f = O.features_batch(seat_masks(rec).reshape(-1)).reshape(-1, 4, 11)
    for k in range(4):
        t[4 + 40 * k: 4 + 40 * k + 38] = np.bincount(f[:, k, 8], minlength=38)[:38]
        for s in range(4):
            off = 4 + 160 + (k * 4 + s) * 14
            t[off: off + 14] = np.bincount(f[:, k, s], minlength=14)
    return t


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    bdist.init_process_group("gloo")
    first, count = bdist.rank_range(FIRST, N, rank, world)
    t = torch.from_numpy(cpu_tallies(first, count))
    bdist.allreduce_tallies(t)
    ms = bdist.max_over_ranks(torch.tensor([float(rank + 1)], dtype=torch.float64))
    if rank == 0:
        np.save(os.path.join(out_dir, "merged.npy"), t.numpy())
        np.save(os.path.join(out_dir, "ms.npy"), ms.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_rank_ranges_tile_exactly():
    for n in (0, 1, 7, 1 << 20, 10 ** 11 + 3):
        for world in (1, 2, 3, 4, 8):
            shares = [bdist.rank_range(100, n, r, world) for r in range(world)]
            assert shares[0][0] == 100 and sum(c for _, c in shares) == n
            for (a, ca), (b, _) in zip(shares, shares[1:]):
                assert a + ca == b
    firsts = {bdist.step_first_index(s, r, 8) for s in range(50) for r in range(8)}
    assert len(firsts) == 400 and min(np.diff(sorted(firsts))) == 1 << 40


def test_two_rank_gloo_merge_is_bit_identical(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    merged = np.load(tmp_path / "merged.npy")
    assert np.array_equal(merged, cpu_tallies(FIRST, N))
    assert np.load(tmp_path / "ms.npy")[0] == 2.0            # timings are the max over ranks
