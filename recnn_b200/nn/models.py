"""Actor / Critic / DiscreteActor with the reference's constructor, attributes and state_dict
layout (recnn/nn/models.py:41-73, :76-184, :187-213), evaluated by the sm_100a kernels.

``forward`` is the inference / evaluation entry (no autograd graph): training
goes through recnn_b200.nn.update.*, which runs forward+backward+optimizer as
one fused device step.  There is no CPU implementation behind ``forward`` --
calling it on CPU tensors raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from .arena import param_arena, discrete_dims


def _dims(state_dim, action_dim, hidden):
    return _lib.Dims(int(state_dim), int(action_dim), int(hidden), 0)


def _device_check(module, *tensors):
    dev = module.linear1.weight.device
    if dev.type != "cuda":
        raise _lib.RecnnError("recnn_b200 nets run on CUDA only (module is on %s); call .cuda() first" % dev)
    out = []
    for t in tensors:
        out.append(t.detach().to(device=dev, dtype=torch.float32).contiguous())
    return dev, out


def _train_masks(module, n_rows, hidden, device):
    """Dropout(p=0.5) keep-masks for the two hidden layers (train mode only)."""
    if not module.training:
        return None, None
    keep = torch.rand(2, n_rows, hidden, device=device) >= 0.5
    keep = keep.to(torch.uint8)
    return keep[0], keep[1]


class Actor(nn.Module):
    """Vanilla actor: state -> action.  Same signature as recnn.nn.Actor."""

    def __init__(self, input_dim, action_dim, hidden_size, init_w=2e-1):
        super().__init__()
        self.drop_layer = nn.Dropout(p=0.5)       # kept for attribute parity; p is fixed at 0.5 in the kernels
        self.linear1 = nn.Linear(input_dim, hidden_size)
        self.linear2 = nn.Linear(hidden_size, hidden_size)
        self.linear3 = nn.Linear(hidden_size, action_dim)
        self.linear3.weight.data.uniform_(-init_w, init_w)
        self.linear3.bias.data.uniform_(-init_w, init_w)

    @property
    def dims(self):
        return _dims(self.linear1.in_features, self.linear3.out_features, self.linear1.out_features)

    def forward(self, state, tanh=False, masks=None):
        dev, (state,) = _device_check(self, state)
        n = state.shape[0]
        d = self.dims
        flat = param_arena(self)
        m1, m2 = masks if masks is not None else _train_masks(self, n, d.hidden, dev)
        out = torch.empty(n, d.action_dim, device=dev, dtype=torch.float32)
        scratch = torch.empty(_lib.lib().recnn_forward_scratch_floats(d, n, 0), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().recnn_actor_forward(
                d, flat.data_ptr(), state.data_ptr(), n, _lib.ptr(m1), _lib.ptr(m2), int(bool(tanh)),
                out.data_ptr(), scratch.data_ptr(), _lib.stream_ptr(dev)))
        return out


class Critic(nn.Module):
    """Vanilla critic: (state, action) -> value [N,1].  Same signature as recnn.nn.Critic."""

    def __init__(self, input_dim, action_dim, hidden_size, init_w=3e-5):
        super().__init__()
        self.drop_layer = nn.Dropout(p=0.5)
        self.linear1 = nn.Linear(input_dim + action_dim, hidden_size)
        self.linear2 = nn.Linear(hidden_size, hidden_size)
        self.linear3 = nn.Linear(hidden_size, 1)
        self.linear3.weight.data.uniform_(-init_w, init_w)
        self.linear3.bias.data.uniform_(-init_w, init_w)
        self._action_dim = int(action_dim)

    @property
    def dims(self):
        a = self._action_dim
        return _dims(self.linear1.in_features - a, a, self.linear1.out_features)

    def forward(self, state, action, masks=None):
        dev, (state, action) = _device_check(self, state, action)
        n = state.shape[0]
        d = self.dims
        flat = param_arena(self)
        m1, m2 = masks if masks is not None else _train_masks(self, n, d.hidden, dev)
        out = torch.empty(n, 1, device=dev, dtype=torch.float32)
        scratch = torch.empty(_lib.lib().recnn_forward_scratch_floats(d, n, 1), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().recnn_critic_forward(
                d, flat.data_ptr(), state.data_ptr(), action.data_ptr(), n, _lib.ptr(m1), _lib.ptr(m2),
                out.data_ptr(), scratch.data_ptr(), _lib.stream_ptr(dev)))
        return out


class DiscreteActor(nn.Module):
    """REINFORCE policy over a discrete item set: state -> probabilities [N, action_dim]
    (recnn/nn/models.py:76-184; same constructor, attributes and methods).

    Differences from the reference, all forced by running without an autograd graph:

    * ``saved_log_probs`` / ``correction`` / ``lambda_k`` hold the same VALUES (device tensors) but carry no graph; the
      policy update recomputes the forward over the rows saved here (``_saved``: state, drawn action, beta log-prob per
      env step) inside one fused backward (recnn_reinforce_policy_grad) -- valid because the policy's weights do not
      change between two policy updates.  ``gc()`` drops both.
    * Categorical draws are inverse-CDF draws on a counter-based Philox stream keyed by ``torch.initial_seed()`` (or on
      ``uniform_source(n_rows) -> tensor[n_rows]`` when set: replayable draws for tests), not torch's global generator.
    """

    def __init__(self, input_dim, action_dim, hidden_size, init_w=0):
        super().__init__()
        self.linear1 = nn.Linear(input_dim, hidden_size)
        self.linear2 = nn.Linear(hidden_size, action_dim)
        self.saved_log_probs = []
        self.rewards = []
        self.correction = []
        self.lambda_k = []
        # {pi: pi, beta: beta} by default; {pi: beta, beta: beta} is the variant of awarebayes/RecNN issue 7
        self.action_source = {"pi": "pi", "beta": "beta"}
        self.select_action = self._select_action
        self.uniform_source = None
        self._saved = []
        self._draws = 0

    @property
    def dims(self):
        return discrete_dims(self)

    def forward(self, inputs):
        dev, (state,) = _device_check(self, inputs)
        n = state.shape[0]
        d = self.dims
        flat = param_arena(self)
        out = torch.empty(n, d.num_items, device=dev, dtype=torch.float32)
        if n == 0:
            return out
        L = _lib.lib()
        scratch = torch.empty(L.recnn_discrete_scratch_floats(d, n, 0), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(L.recnn_discrete_forward(d, flat.data_ptr(), state.data_ptr(), n, out.data_ptr(),
                                                scratch.data_ptr(), _lib.stream_ptr(dev)))
        return out

    def gc(self):
        del self.rewards[:]
        del self.saved_log_probs[:]
        del self.correction[:]
        del self.lambda_k[:]
        del self._saved[:]

    # -- Categorical(probs).sample() / .log_prob() on the device ----------------------------------------------------
    def _sample(self, probs):
        """(action int64 [N], log_prob fp32 [N]) of one draw per row."""
        n, items = probs.shape
        dev = probs.device
        action = torch.empty(n, dtype=torch.int64, device=dev)
        logp = torch.empty(n, dtype=torch.float32, device=dev)
        u = None
        if self.uniform_source is not None:
            u = torch.as_tensor(self.uniform_source(n)).to(device=dev, dtype=torch.float32).contiguous()
            if u.shape != (n,):
                raise ValueError("uniform_source must return %d values" % n)
        self._draws += 1
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().recnn_categorical_sample(
                probs.data_ptr(), n, items, probs.stride(0), _lib.ptr(u), int(torch.initial_seed()) & (2 ** 64 - 1),
                self._draws, action.data_ptr(), logp.data_ptr(), _lib.stream_ptr(dev)))
        return action, logp

    @staticmethod
    def _log_prob(probs, action):
        n, items = probs.shape
        dev = probs.device
        logp = torch.empty(n, dtype=torch.float32, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        action = action.to(device=dev, dtype=torch.int64).contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().recnn_categorical_log_prob(probs.data_ptr(), n, items, probs.stride(0),
                                                             action.data_ptr(), logp.data_ptr(), flag.data_ptr(),
                                                             _lib.stream_ptr(dev)))
        if int(flag.item()) != 0:
            raise IndexError("action index out of range for the policy's output layer")
        return logp

    @staticmethod
    def _as_probs(t, dev):
        t = t.detach().to(device=dev, dtype=torch.float32)
        return t if t.stride(1) == 1 else t.contiguous()

    def _select_action(self, state, **kwargs):
        # REINFORCE without correction: only pi is available, the action source is ignored (models.py:102-111)
        dev, (state,) = _device_check(self, state)
        pi_probs = self.forward(state)
        pi_action, pi_log_prob = self._sample(pi_probs)
        self.saved_log_probs.append(pi_log_prob)
        self._saved.append({"state": state.clone(), "action": pi_action, "beta_log_prob": None})
        return pi_probs

    def pi_beta_sample(self, state, beta, action, **kwargs):
        """models.py:113-145.  ``beta`` is any callable (state, action=...) -> probabilities [N, action_dim]."""
        dev, (state,) = _device_check(self, state)
        beta_probs = self._as_probs(beta(state.detach(), action=action), dev)
        pi_probs = self.forward(state)
        # the pi draw is made first, then the beta draw (models.py:133-136)
        pi_draw = self._sample(pi_probs)
        beta_draw = self._sample(beta_probs)
        available = {"pi": (pi_draw, pi_probs), "beta": (beta_draw, beta_probs)}
        (pi_action, pi_lp), src_pi = available[self.action_source["pi"]]
        (beta_action, beta_lp), src_beta = available[self.action_source["beta"]]
        pi_log_prob = pi_lp if src_pi is pi_probs else self._log_prob(pi_probs, pi_action)
        beta_log_prob = beta_lp if src_beta is beta_probs else self._log_prob(beta_probs, beta_action)
        self._last_sample = {"state": state, "action": pi_action, "beta_log_prob": beta_log_prob}
        return pi_log_prob, beta_log_prob, pi_probs

    def _select_action_with_correction(self, state, beta, action, writer, step, **kwargs):
        pi_log_prob, beta_log_prob, pi_probs = self.pi_beta_sample(state, beta, action)
        corr = torch.exp(pi_log_prob) / torch.exp(beta_log_prob)
        writer.add_histogram("correction", corr, step)
        writer.add_histogram("pi_log_prob", pi_log_prob, step)
        writer.add_histogram("beta_log_prob", beta_log_prob, step)
        self.correction.append(corr)
        self.saved_log_probs.append(pi_log_prob)
        rec = self._last_sample
        self._saved.append({"state": rec["state"].clone(), "action": rec["action"], "beta_log_prob": rec["beta_log_prob"]})
        return pi_probs

    def _select_action_with_TopK_correction(self, state, beta, action, K, writer, step, **kwargs):
        pi_log_prob, beta_log_prob, pi_probs = self.pi_beta_sample(state, beta, action)
        corr = torch.exp(pi_log_prob) / torch.exp(beta_log_prob)
        l_k = K * (1 - torch.exp(pi_log_prob)) ** (K - 1)
        writer.add_histogram("correction", corr, step)
        writer.add_histogram("l_k", l_k, step)
        writer.add_histogram("pi_log_prob", pi_log_prob, step)
        writer.add_histogram("beta_log_prob", beta_log_prob, step)
        self.correction.append(corr)
        self.lambda_k.append(l_k)
        self.saved_log_probs.append(pi_log_prob)
        rec = self._last_sample
        self._saved.append({"state": rec["state"].clone(), "action": rec["action"], "beta_log_prob": rec["beta_log_prob"],
                            "K": int(K)})
        return pi_probs
