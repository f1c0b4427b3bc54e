"""Flat parameter / gradient arenas behind Actor and Critic.

The CUDA kernels address a net as ONE contiguous fp32 buffer in
``nn.Module.parameters()`` order (include/recnn_b200.h, "net").  The modules keep
ordinary ``nn.Parameter`` objects -- ``state_dict`` keys linear{1,2,3}.{weight,bias}
stay loadable (examples/streamlit_demo.py:152-160 in the reference) and external
torch optimizers keep working -- but their storage is a view into the arena.
``module.to(device)``, ``load_state_dict`` into fresh tensors or
``zero_grad(set_to_none=True)`` can break that aliasing, so every kernel entry
re-validates it (a few pointer compares) and rebuilds the arena if needed.
"""
from __future__ import annotations

import torch


def _params(module):
    return [module.linear1.weight, module.linear1.bias, module.linear2.weight, module.linear2.bias,
            module.linear3.weight, module.linear3.bias]


def _is_view_of(flat, tensors):
    if flat is None:
        return False
    off = 0
    base = flat.data_ptr()
    for t in tensors:
        if t is None or t.device != flat.device or t.dtype != torch.float32 or not t.is_contiguous():
            return False
        if t.data_ptr() != base + 4 * off:
            return False
        off += t.numel()
    return off == flat.numel()


def param_arena(module) -> torch.Tensor:
    """Flat fp32 view of all parameters of ``module`` (rebuilt if aliasing broke)."""
    ps = _params(module)
    flat = getattr(module, "_recnn_flat", None)
    if _is_view_of(flat, [p.data for p in ps]):
        return flat
    dev = ps[0].device
    count = sum(p.numel() for p in ps)
    flat = torch.empty(count, dtype=torch.float32, device=dev)
    off = 0
    with torch.no_grad():
        for p in ps:
            n = p.numel()
            view = flat[off:off + n].view(p.shape)
            view.copy_(p.data.to(device=dev, dtype=torch.float32))
            p.data = view
            off += n
    object.__setattr__(module, "_recnn_flat", flat)
    object.__setattr__(module, "_recnn_flat_grad", None)
    return flat


def grad_arena(module) -> torch.Tensor:
    """Flat fp32 gradient buffer; ``p.grad`` of every parameter is a view into it."""
    ps = _params(module)
    flat = param_arena(module)
    g = getattr(module, "_recnn_flat_grad", None)
    if g is None or g.device != flat.device or g.numel() != flat.numel():
        g = torch.zeros_like(flat)
        object.__setattr__(module, "_recnn_flat_grad", g)
    if not _is_view_of(g, [p.grad for p in ps]):
        off = 0
        for p in ps:
            n = p.numel()
            p.grad = g[off:off + n].view(p.shape)
            off += n
    return g
