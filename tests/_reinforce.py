"""Shared loader for the REINFORCE golden cases (tests/golden/reinforce_*.npz, made by oracle/make_reinforce_golden.py)."""
from __future__ import annotations

import glob
import os

import numpy as np

from oracle import reinforce_oracle as RO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLDEN, "reinforce_*.npz")))
IDS = [os.path.basename(f)[len("reinforce_"):-4] for f in FILES]


def load(path):
    g = dict(np.load(path))
    S, H, I, N, T, K = (int(x) for x in g["dims"])
    name = os.path.basename(path)[len("reinforce_"):-4]
    method_name = name.split("_", 1)[1]
    case = {
        "g": g, "S": S, "H": H, "I": I, "N": N, "T": T, "K": K, "method_name": method_name,
        "method": RO.METHODS[method_name],
        "params": {"w1": g["param.linear1.weight"], "b1": g["param.linear1.bias"],
                   "w2": g["param.linear2.weight"], "b2": g["param.linear2.bias"]},
    }
    # models.py:137-140: which draw plays the role of the policy's / the behaviour policy's action
    pi_from_beta = bool(g["source_pi_is_beta"]) and case["method"] != RO.BASIC
    beta_from_pi = bool(g["source_beta_is_pi"])
    case["pi_action"] = g["beta_draws"] if pi_from_beta else g["pi_draws"]
    case["beta_action"] = g["pi_draws"] if beta_from_pi else g["beta_draws"]
    # the saved rows, concatenated in env-step order
    case["rows_state"] = g["states"].reshape(T * N, S)
    case["rows_action"] = case["pi_action"].reshape(T * N).astype(np.int64)
    case["rows_step"] = np.repeat(np.arange(T), N)
    if case["method"] != RO.BASIC:
        blp = np.stack([RO.categorical_log_prob(g["beta_probs"][t], case["beta_action"][t]) for t in range(T)])
        case["rows_beta_logp"] = blp.reshape(T * N).astype(np.float32)
    else:
        case["rows_beta_logp"] = None
    return case
