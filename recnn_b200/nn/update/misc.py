"""recnn.nn.update.misc equivalents (recnn/nn/update/misc.py:6-55)."""
from __future__ import annotations

import torch

from ... import _lib
from ... import utils
from ._engine import get_engine


def temporal_difference(reward, done, gamma, target):
    """reward + (1 - done) * gamma * target   (misc.py:6-7).  Plain tensor algebra kept for
    API parity; inside the update step it is fused into the critic-head kernel."""
    return reward + (1.0 - done) * gamma * target


def value_update(batch, params, nets, optimizer, device=torch.device("cpu"), debug=None,
                 writer=utils.DummyWriter(), learn=False, step=-1):
    """DDPG critic step on its own (misc.py:10-55).  Returns the value loss as a 0-dim tensor
    (the reference returns the loss tensor, not a float)."""
    eng = get_engine(_lib.ALGO_DDPG, nets, device)
    vals = eng.value_only(batch, params, nets, optimizer, learn, debug)
    return torch.tensor(vals[0])
