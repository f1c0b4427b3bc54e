"""Pick case seeds whose ReLU gate decisions are unambiguous (TEST INFRASTRUCTURE).

For every (case, algo) try seeds in order and keep the first one for which the numpy oracle's
smallest kept |pre-activation| over all 12 steps, for both optimizers, exceeds GUARD.  The chosen
seeds are pasted into oracle/cases.py; oracle/make_golden.py re-measures the margin on the real
reference (forward hooks on its Linear layers) and stores it in the fixture.

    python -m oracle.find_seeds
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import cases as C                    # noqa: E402
from oracle import recnn_oracle as O             # noqa: E402

GUARD = C.GATE_GUARD


def margin(case, algo, seed):
    import tests._golden as G
    spec = dict(C.CASES[case], seeds={algo: seed})
    C.CASES["_probe"] = spec
    worst = float("inf")
    golden = None
    if algo == "td3":                      # same CPU-generator draws as oracle/make_golden.py
        import torch
        golden = {}
        for step in range(spec["steps"]):
            torch.manual_seed(9000 + step)
            golden["noise.%d" % step] = torch.normal(torch.zeros(spec["n_rows"], spec["dim"]),
                                                     C.TD3_PARAMS["noise_std"]).numpy()
    for opt in ("adam", "sgd"):
        O.reset_gate_margin()
        G.run_oracle_case("_probe", algo, opt, golden=golden)
        worst = min(worst, O.GATE_MARGIN["min"])
    del C.CASES["_probe"]
    return worst


def main():
    for case in ("tiny", "canon"):
        for algo in ("ddpg", "td3"):
            for seed in range(1, 400):
                m = margin(case, algo, seed)
                if m > GUARD:
                    print("%s %s: seed %d margin %.3g" % (case, algo, seed, m))
                    break
            else:
                print(case, algo, "no seed found")


if __name__ == "__main__":
    main()
