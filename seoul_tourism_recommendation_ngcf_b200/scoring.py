"""Fused scoring + top-k: the ``torch.mm(u, all_i_emb.T)`` + ``torch.topk`` pair of demo.py:234-235 and
experiment.py:93,104,109 without materialising the score matrix."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


@_lib.on_device
def score_topk(u_embeds: torch.Tensor, item_embeds: torch.Tensor, k: int):
    """Returns (values [U,k] fp32, indices [U,k] int64), descending, like ``torch.topk(u @ items.T, k)``."""
    if u_embeds.device.type != "cuda" or item_embeds.device.type != "cuda":
        raise RuntimeError("score_topk (B200) runs on CUDA tensors only; there is no CPU fallback")
    lib = _lib.load()
    u = u_embeds.detach().to(torch.float32).contiguous()
    it = item_embeds.detach().to(torch.float32).contiguous()
    if u.dim() != 2 or it.dim() != 2 or u.shape[1] != it.shape[1]:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(u.shape)} and {tuple(it.T.shape)})")
    U, D = u.shape
    n_items = it.shape[0]
    if k > n_items:
        raise RuntimeError("selected index k out of range")          # what torch.topk raises
    need = C.c_size_t(0)
    _lib.check(lib.ngcf_score_topk_workspace(U, n_items, D, k, C.byref(need)), "score_topk_workspace")
    ws = torch.empty(need.value, dtype=torch.uint8, device=u.device)
    val = torch.empty(U, k, dtype=torch.float32, device=u.device)
    idx = torch.empty(U, k, dtype=torch.int64, device=u.device)
    _lib.check(lib.ngcf_score_topk(u.data_ptr(), U, it.data_ptr(), n_items, D, k, val.data_ptr(), idx.data_ptr(),
                                   ws.data_ptr(), need.value, _lib.current_stream()), "score_topk")
    return val, idx
