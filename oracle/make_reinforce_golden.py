"""Generate tests/golden/reinforce_*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_reinforce_golden

Pinned (the reference ships no vectors for this path): ``recnn.nn.models.DiscreteActor`` --
``forward``, ``_select_action``, ``_select_action_with_correction``, ``_select_action_with_TopK_correction`` over T env
steps with both ``action_source`` settings -- and ``recnn.nn.update.reinforce.ChooseREINFORCE`` with its three methods
(``learn=True`` with torch.optim.SGD so that the gradient is observable both as ``.grad`` and as the stepped
parameters).  The Categorical draws are made observable by replacing the name ``Categorical`` inside the imported
``recnn.nn.models`` module with a recording subclass (an attribute of the imported module; no reference file is
modified) -- the recorded draws are what the CUDA path and the oracle are then told to replay.

Stored per case: parameters, the T state batches, the behaviour policy's probabilities, every draw, the log-probs,
corrections and lambda_K the reference appended to its lists, the T rewards, the normalised returns it computed, the
policy loss, every parameter's .grad and the parameters after the SGD step.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (state_dim, hidden, num_items, rows per env step, env steps, K, action_source)
    "tiny": (13, 16, 37, 6, 4, 3, {"pi": "pi", "beta": "beta"}),
    "small": (52, 64, 300, 10, 11, 10, {"pi": "beta", "beta": "beta"}),
}


class _Writer:
    def add_histogram(self, *a, **k):
        pass


def run_case(name, method_name):
    recnn = import_reference()
    models = recnn.nn.models
    from recnn.nn.update.reinforce import ChooseREINFORCE

    S, H, I, N, T, K, source = CASES[name]
    seed = 1234 + sum(map(ord, name + method_name))
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)

    draws = []

    class RecordingCategorical(torch.distributions.Categorical):
        def sample(self, *a, **k):
            out = super().sample(*a, **k)
            draws.append(out.clone())
            return out

    saved = models.Categorical
    models.Categorical = RecordingCategorical
    try:
        policy = models.DiscreteActor(S, I, H)
        policy.action_source = dict(source)
        params0 = {k: v.detach().clone().numpy() for k, v in policy.state_dict().items()}
        beta_w = torch.tensor(rng.normal(0, 0.3, (I, S)).astype(np.float32))

        def beta(state, action=None):
            return torch.softmax(state @ beta_w.T, dim=1)

        states, beta_probs, rewards, fwd_probs = [], [], [], []
        pi_draws, beta_draws = [], []
        for t in range(T):
            state = torch.tensor(rng.normal(0, 1, (N, S)).astype(np.float32))
            states.append(state.numpy().copy())
            n0 = len(draws)
            if method_name == "basic_reinforce":
                probs = policy._select_action(state)
                pi_draws.append(draws[n0].numpy().copy())
                beta_draws.append(np.zeros(N, np.int64))
                beta_probs.append(np.zeros((N, I), np.float32))
            else:
                beta_probs.append(beta(state).numpy().copy())
                if method_name == "reinforce_with_correction":
                    probs = policy._select_action_with_correction(state, beta, None, _Writer(), t)
                else:
                    probs = policy._select_action_with_TopK_correction(state, beta, None, K, _Writer(), t)
                # models.py:133-136: the pi draw is made first, then the beta draw
                pi_draws.append(draws[n0].numpy().copy())
                beta_draws.append(draws[n0 + 1].numpy().copy())
            fwd_probs.append(probs.detach().numpy().copy())
            r = torch.tensor(np.float32(rng.normal(0, 1)))
            policy.rewards.append(r)
            rewards.append(float(r))
        out = {
            "dims": np.asarray([S, H, I, N, T, K], np.int64),
            "source_pi_is_beta": np.asarray(int(source["pi"] == "beta")),
            "source_beta_is_pi": np.asarray(int(source["beta"] == "pi")),
            "states": np.stack(states), "beta_probs": np.stack(beta_probs), "probs": np.stack(fwd_probs),
            "pi_draws": np.stack(pi_draws), "beta_draws": np.stack(beta_draws),
            "rewards": np.asarray(rewards, np.float32),
            "saved_log_probs": np.stack([x.detach().numpy() for x in policy.saved_log_probs]),
        }
        if policy.correction:
            out["correction"] = np.stack([x.detach().numpy() for x in policy.correction])
        if policy.lambda_k:
            out["lambda_k"] = np.stack([x.detach().numpy() for x in policy.lambda_k])
        # the returns exactly as reinforce.py:44-52 forms them
        R = 0
        rets = []
        for r in policy.rewards[::-1]:
            R = r + 0.99 * R
            rets.insert(0, R)
        rets = torch.tensor(rets)
        out["returns"] = ((rets - rets.mean()) / (rets.std() + 0.0001)).numpy()
        lr = 0.05
        opt = torch.optim.SGD(policy.parameters(), lr=lr)
        loss = ChooseREINFORCE(getattr(ChooseREINFORCE, method_name))(policy, opt, learn=True)
        out["lr"] = np.asarray(lr)
        out["loss"] = np.asarray(float(loss))
        for k, v in policy.named_parameters():
            out["grad." + k] = v.grad.detach().numpy().copy()
            out["after." + k] = v.detach().numpy().copy()
        for k, v in params0.items():
            out["param." + k] = v
        assert len(policy.saved_log_probs) == 0 and len(policy.rewards) == 0      # gc() ran (reinforce.py:62)
    finally:
        models.Categorical = saved
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in CASES:
        for method in ("basic_reinforce", "reinforce_with_correction", "reinforce_with_TopK_correction"):
            out = run_case(name, method)
            path = os.path.join(GOLDEN_DIR, "reinforce_%s_%s.npz" % (name, method))
            np.savez_compressed(path, **out)
            print(path, "loss", float(out["loss"]), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
