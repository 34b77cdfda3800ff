import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bridge_mc_b200.engine import Engine
eng = Engine(0)
n = 1 << 30
for mask in (0, 1, 2, 4, 3, 7):
    for i in range(2):
        rec, tal, st = eng.sample(None, n_raw=n, first_index=i << 40, cap=0, tally_mask=mask)
    print("tally_mask", mask, "ms", round(st.kernel_ms, 2), "deals/s %.3e" % (tal["raw"] / st.kernel_ms * 1e3))
