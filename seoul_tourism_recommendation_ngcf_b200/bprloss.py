"""Drop-in ``BPR`` loss (reference: model/bprloss.py:9-22) as one fused forward+gradient kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class _BprFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, p, n, weight_decay, batch_size, wu, wp, wn):
        lib = _lib.load()
        B, D = u.shape
        need_grad = any(ctx.needs_input_grad[:3])
        loss = torch.empty((), dtype=torch.float32, device=u.device)
        if need_grad:
            g = torch.empty(3, B, D, dtype=torch.float32, device=u.device)
            gu, gp, gn = g[0], g[1], g[2]
        else:
            gu = gp = gn = None
        _lib.check(lib.ngcf_bpr_fwd_bwd(u.data_ptr(), p.data_ptr(), n.data_ptr(), B, D, float(weight_decay),
                                        float(batch_size), wu, wp, wn, loss.data_ptr(), _lib.ptr(gu), _lib.ptr(gp),
                                        _lib.ptr(gn), _lib.current_stream()), "bpr_fwd_bwd")
        ctx.grads = (gu, gp, gn)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        gu, gp, gn = ctx.grads
        return gu * gloss, gp * gloss, gn * gloss, None, None, None, None, None


class BPR(nn.Module):
    def __init__(self, weight_decay, batch_size):
        super().__init__()
        self.weight_decay = weight_decay
        self.batch_size = batch_size          # the divisor is this ctor value, not the row count (bprloss.py:22)

    @_lib.on_device
    def forward(self, u_idx, pos_idx, neg_idx):
        if u_idx.device.type != "cuda":
            raise RuntimeError("BPR (B200) runs on CUDA tensors only; there is no CPU fallback")
        ts = [t.to(dtype=torch.float32) for t in (u_idx, pos_idx, neg_idx)]
        if any(t.dim() != 2 for t in ts):
            raise ValueError("BPR expects three [rows, D] tensors (bprloss.py:16-17)")
        B = max(t.shape[0] for t in ts)
        D = ts[0].shape[1]
        ws = []
        for i, t in enumerate(ts):
            if t.shape[1] != D or t.shape[0] not in (1, B):
                raise RuntimeError(f"The size of tensor a ({tuple(ts[0].shape)}) must match the size of tensor b "
                                   f"({tuple(t.shape)}) (bprloss.py:16-17)")
            if t.shape[0] != B:
                # torch.mul broadcasts a single row (experiment.py:96-100 passes pos_i_embeds[:1]); its norm is
                # regularised once (bprloss.py:20-21), hence the 1/B weight on the expanded rows
                ts[i] = t.expand(B, D)
                ws.append(1.0 / B)
            else:
                ws.append(1.0)
        ts = [t.contiguous() for t in ts]
        return _BprFn.apply(ts[0], ts[1], ts[2], self.weight_decay, self.batch_size, ws[0], ws[1], ws[2])
