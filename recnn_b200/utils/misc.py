"""recnn.utils.misc equivalents (recnn/utils/misc.py:1-35)."""
from __future__ import annotations

import torch

from .. import _lib


def soft_update(net, target_net, soft_tau=1e-2):
    """target <- target*(1-tau) + net*tau over the whole net in one launch
    (reference: six out-of-place tensor expressions per net, utils/misc.py:1-5)."""
    from ..nn.arena import param_arena
    src = param_arena(net)
    dst = param_arena(target_net)
    if src.device.type != "cuda" or dst.device != src.device:
        raise _lib.RecnnError("soft_update needs both nets on the same CUDA device")
    if src.numel() != dst.numel():
        raise ValueError("nets have different parameter counts")
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().recnn_polyak_update(dst.data_ptr(), src.data_ptr(), src.numel(), float(soft_tau),
                                                  _lib.stream_ptr(src.device)))


def write_losses(writer, loss_dict, kind="train"):
    step = loss_dict["step"]
    for key, value in loss_dict.items():
        if key == "step":
            continue
        writer.add_scalar(kind + "/" + key, value, global_step=step)
    writer.close()


class DummyWriter:
    """Stand-in for torch.utils.tensorboard.SummaryWriter that drops everything (utils/misc.py:15-35): every
    ``add_*`` call and ``close`` is accepted and ignored."""

    @staticmethod
    def _ignore(*args, **kwargs):
        return None

    add_figure = add_histogram = add_scalar = add_scalars = close = _ignore
