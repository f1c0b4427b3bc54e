"""Algo / DDPG / TD3 / Reinforce wrappers with the reference's wiring (recnn/nn/algo.py:15-233).

These are the dispatchers the update functions plug into (``Algo.algorithm``);
they only hold state.  The reference builds ``torch_optimizer.Ranger(lr=1e-5,
weight_decay=1e-2)`` optimizers (algo.py:84-89), a third-party package that is not
available here; the default is recnn_b200.optim.Ranger -- the
same algorithm (RAdam + Lookahead) with the package's defaults, fused into the device
step, parity with the package unpinned -- and any optimizer can be supplied
through ``self.optimizers`` as in the reference.
"""
from __future__ import annotations

import copy

import torch

from .. import optim, utils
from . import update


def _loss_layout(*keys):
    """{"test": {key: []...}, "train": {...}} with a trailing "step" series, as the plotting helpers expect."""
    return {kind: {k: [] for k in keys + ("step",)} for kind in ("test", "train")}


def _default_optimizer(net):
    # the reference's Ranger(lr=1e-5, weight_decay=1e-2) for every net (algo.py:84-89, :139-147, :201-206)
    return optim.Ranger(net.parameters(), lr=1e-5, weight_decay=1e-2)


def _target_of(net):
    """deepcopy + eval == the reference's copy, .eval() and soft_update(tau=1.0) (algo.py:73-81)."""
    target = copy.deepcopy(net)
    target.__dict__.pop("_recnn_engines", None)
    target.eval()
    return target


def _with_targets(**online):
    nets = {}
    for name, net in online.items():
        nets[name] = net
        nets["target_" + name] = _target_of(net)
    return nets


class Algo:
    """State holder + dispatcher (algo.py:15-62): ``update`` forwards to ``self.algorithm``."""

    def __init__(self):
        self.nets = {"value_net": None, "policy_net": None}
        self.optimizers = {"policy_optimizer": None, "value_optimizer": None}
        self.params = {"Some parameters here": None}
        self._step = 0
        self.debug = {}
        self.writer = utils.misc.DummyWriter()
        self.device = torch.device("cpu")
        self.loss_layout = _loss_layout("value", "policy")
        self.algorithm = None

    def update(self, batch, learn=True):
        return self.algorithm(batch, self.params, self.nets, self.optimizers, device=self.device,
                              debug=self.debug, writer=self.writer, learn=learn, step=self._step)

    def to(self, device):
        self.nets = {k: v.to(device) for k, v in self.nets.items()}
        self.device = device
        return self

    def step(self):
        self._step += 1


class DDPG(Algo):
    """algo.py:65-114."""

    def __init__(self, policy_net, value_net):
        super().__init__()
        self.algorithm = update.ddpg_update
        self.nets = _with_targets(value_net=value_net, policy_net=policy_net)
        self.optimizers = {"policy_optimizer": _default_optimizer(policy_net),
                           "value_optimizer": _default_optimizer(value_net)}
        self.params = dict(gamma=0.99, min_value=-10, max_value=10, policy_step=10, soft_tau=0.001)
        self.loss_layout = _loss_layout("value", "policy")


class TD3(Algo):
    """algo.py:117-179."""

    def __init__(self, policy_net, value_net1, value_net2):
        super().__init__()
        self.algorithm = update.td3_update
        self.nets = _with_targets(value_net1=value_net1, value_net2=value_net2, policy_net=policy_net)
        self.optimizers = {"policy_optimizer": _default_optimizer(policy_net),
                           "value_optimizer1": _default_optimizer(value_net1),
                           "value_optimizer2": _default_optimizer(value_net2)}
        self.params = dict(gamma=0.99, noise_std=0.5, noise_clip=3, soft_tau=0.001, policy_update=10,
                           policy_lr=1e-5, value_lr=1e-5, actor_weight_init=25e-2, critic_weight_init=6e-1)
        self.loss_layout = _loss_layout("value1", "value2", "policy")


class Reinforce(Algo):
    """algo.py:182-233."""

    def __init__(self, policy_net, value_net):
        super().__init__()
        self.algorithm = update.reinforce_update
        self.nets = _with_targets(value_net=value_net, policy_net=policy_net)
        self.optimizers = {"policy_optimizer": _default_optimizer(policy_net),
                           "value_optimizer": _default_optimizer(value_net)}
        self.params = dict(reinforce=update.ChooseREINFORCE(update.ChooseREINFORCE.basic_reinforce), K=10, gamma=0.99,
                           min_value=-10, max_value=10, policy_step=10, soft_tau=0.001)
        self.loss_layout = _loss_layout("value", "policy")
