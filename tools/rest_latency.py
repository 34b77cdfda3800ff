"""Latency of the REST-sized jobs the reference actually runs (rest.cl:264-277: 500 deals per /api/mc job)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bridge_mc_b200.cards import str2hand
from bridge_mc_b200.deal import Deal, McCase, mc_deal
from bridge_mc_b200.engine import Engine

t0 = time.perf_counter()
eng = Engine(0)
print(f"bmc_init: {1e3 * (time.perf_counter() - t0):.1f} ms")
south = str2hand("S: ♠ AKJ5 ♥ K1064 ♦ Q87 ♣ A6")


def ns_spades(trump, n, e, s, w):
    return len(n.suits[3]) + len(s.suits[3])


for rep in range(3):
    t0 = time.perf_counter()
    mc, tree = mc_deal([{"hcp": [10, 12], "s": [4, None]}, {}, south, {}], [ns_spades], runs=500, engine=eng, seed=rep)
    print(f"mc_deal, 500 deals, fixed South + ranged North, histogram tree: {1e3 * (time.perf_counter() - t0):.1f} ms "
          f"({len(tree)} top-level entries)")
d = Deal(engine=eng, seed=1)                      # training.cl:1 full-random
t0 = time.perf_counter()
d.build()
print(f"first (build full-random): {1e3 * (time.perf_counter() - t0):.2f} ms (fills a 2048-deal buffer)")
t0 = time.perf_counter()
for _ in range(1000):
    d.build()
print(f"next 1000 builds: {1e3 * (time.perf_counter() - t0):.1f} ms")
d2 = Deal(defs=[{"hcp": [15, 17], "c": [2, 5], "d": [2, 5], "h": [2, 4], "s": [2, 4]}], engine=eng, seed=2)
t0 = time.perf_counter()
recs = d2.build_records(100_000)
print(f"100 000 box-1NT deals as records: {1e3 * (time.perf_counter() - t0):.1f} ms")
