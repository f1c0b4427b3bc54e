"""Actor / Critic with the reference's constructor, attributes and state_dict
layout (recnn/nn/models.py:41-73, :187-213), evaluated by the sm_100a kernels.

``forward`` is the inference / evaluation entry (no autograd graph): training
goes through recnn_b200.nn.update.*, which runs forward+backward+optimizer as
one fused device step.  There is no CPU implementation behind ``forward`` --
calling it on CPU tensors raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from .arena import param_arena


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
