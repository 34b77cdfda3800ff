#!/usr/bin/env python
"""Extract the reference's own golden vectors into small committed fixtures.

Run in the build container (it reads /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

Outputs (all under tests/golden/):
  kats.json          every top-level `(test FORM EXPECTED PRED)` form of
                     bridge.cl / bidding.cl / compscore.cl / rest.cl whose head
                     function is on the hot path, as prin1 text with file:line
  distgen_golden.json  backend/test/distgen-test.dat (12 012 rows) re-encoded as
                     [weight, pattern16] pairs; pattern16 bit 15 = club J ...
                     bit 0 = spade A (the order of the 16 0/1 flags in each row)

  playdeal_kats.json  every `(test ...)` form inside the trick estimator
                     (bridge.cl:1497-2984): play-low ... play-deal

No reference *source* is copied: only the known-answer data of its test forms.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from bridge_mc_b200 import sexp  # noqa: E402

REF = "/root/reference/backend"

# head symbols (of the tested form, possibly nested one level) we keep
KEEP = {
    "SUIT-LOOSERS", "ASSESS-HAND", "TEST-BID", "CHOOSE-BID", "BASE-SCORE", "MATCH-POINTS",
    "HIST-BASE", "HIST-BREAKDOWN", "HIST-PERCENT", "HISTOGRAM",
    "INFER-HCP-RANGE", "INFER-SUIT-RANGES", "INFER-DISTGEN-PARAMS",
    "HC-PERMUTS", "POSSIBLE-LC-LENS", "PERMUT-WEIGHT", "COMBINATIONS", "TAKE", "REMOVE-CARDS",
    "SUITS", "HANDID", "MAPCAR", "LENGTH", "SCANSTR", "FORMAT", "INVERT-ID", "BID-JUMP",
    "COMBINE-SHAPES", "CARD", "UNHAND", "REORDER",
    # the small list utilities the deal path is written with (bridge_mc_b200/lists.py)
    "CARDS-BY", "VALID-HCP-SUMS", "PUSH-POS", "ROLL", "ZIP-DEALS", "EVEN-SUBLISTS",
}
# comparisons of a tested call against a constant: (< (suit<> 'c 'd) 0)
KEEP_NESTED = {"SUIT<>", "CARD<>"}


def head_of(form):
    return form[0] if isinstance(form, list) and form else None


def main():
    kats = []
    for fname in ("bridge.cl", "bidding.cl", "compscore.cl", "rest.cl"):
        text = open(os.path.join(REF, fname), encoding="utf-8").read()
        for line, form in sexp.read_all_with_lines(text):
            if head_of(form) != "TEST" or len(form) != 4:
                continue
            tested = form[1]
            nested = head_of(tested[1]) if isinstance(tested, list) and len(tested) > 1 else None
            if head_of(tested) not in KEEP and not (head_of(tested) in ("<", ">", "=") and nested in KEEP_NESTED):
                continue
            kats.append({
                "file": fname, "line": line,
                "form": sexp.dumps(tested),
                "expected": sexp.dumps(form[2]),
                "pred": sexp.dumps(form[3]),
            })
    with open(os.path.join(HERE, "kats.json"), "w", encoding="utf-8") as f:
        json.dump(kats, f, ensure_ascii=False, indent=0)
    print(f"kats.json: {len(kats)} known-answer tests")

    # the trick estimator's inline tests (replayed by tests/test_playdeal_oracle.py through a small form evaluator)
    pk = []
    text = open(os.path.join(REF, "bridge.cl"), encoding="utf-8").read()
    for line, form in sexp.read_all_with_lines(text):
        if 1497 <= line <= 2984 and head_of(form) == "TEST" and len(form) == 4:
            pk.append({"file": "bridge.cl", "line": line, "form": sexp.dumps(form[1]),
                       "expected": sexp.dumps(form[2]), "pred": sexp.dumps(form[3])})
    with open(os.path.join(HERE, "playdeal_kats.json"), "w", encoding="utf-8") as f:
        json.dump(pk, f, ensure_ascii=False, indent=0)
    print(f"playdeal_kats.json: {len(pk)} known-answer tests")

    rows = sexp.read_all(open(os.path.join(REF, "test", "distgen-test.dat"), encoding="utf-8").read())
    out = []
    for weight, pattern in rows:
        bits = [b for suit in pattern for b in suit]
        assert len(bits) == 16
        p16 = 0
        for b in bits:
            p16 = (p16 << 1) | b
        out.append([weight, p16])
    with open(os.path.join(HERE, "distgen_golden.json"), "w") as f:
        json.dump({"source": "backend/test/distgen-test.dat, bridge.cl:1047-1051",
                   "spec": ":c (5 nil) :d (0 4) :h (0 1) :hcp (10 15)",
                   "rows": out}, f, separators=(",", ":"))
    print(f"distgen_golden.json: {len(out)} rows, sum={sum(r[0] for r in out)}")


if __name__ == "__main__":
    main()
