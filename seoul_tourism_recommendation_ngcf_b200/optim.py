"""Drop-in for the reference's optimizer: ``optim.Adam(model.parameters(), lr=args.lr)`` (main.py:74), stepped once per
batch (experiment.py:55,58).  Same constructor arguments, same ``state_dict`` layout (``step`` / ``exp_avg`` /
``exp_avg_sq`` per parameter) and the same arithmetic as torch's non-amsgrad Adam, but ONE launch of
``ngcf_adam_step`` for all parameters instead of a dozen foreach kernels; parameters whose gradient is ``None`` (the
feature tables, NGCF.py:115) are skipped like torch does.  No CPU fallback."""
from __future__ import annotations

import torch

from . import _lib


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step_dev = None        # device step counter shared by all parameters (CUDA-graph replay safe)

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = torch.zeros((), dtype=torch.float32)          # torch.optim.Adam's layout
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def prepare(self, params):
        """Allocates the moment buffers and the device step counter of ``params`` now (GraphedStep calls this before it
        captures: a capture must not allocate optimizer state or read a counter back)."""
        params = list(params)
        for p in params:
            self._state(p)
        if params and (self._step_dev is None or self._step_dev.device != params[0].device):
            self._step_dev = torch.full((1,), int(self.state[params[0]]["step"]), dtype=torch.int64,
                                        device=params[0].device)

    @torch.no_grad()
    def step(self, closure=None, zero_grads: bool = False):
        """One Adam update of every parameter that has a gradient; ``zero_grads=True`` also clears the gradients in the
        same pass (the ``optimizer.zero_grad()`` of experiment.py:55 for the next batch)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        capturing = torch.cuda.is_current_stream_capturing() if torch.cuda.is_available() else False
        groups = [(g, [p for p in g["params"] if p.grad is not None]) for g in self.param_groups]
        live = [p for _, ps in groups for p in ps]
        if not live:
            return loss
        dev = live[0].device
        if dev.type != "cuda":
            raise RuntimeError("ngcf_b200 Adam runs on CUDA parameters only (no CPU fallback)")
        # ONE step counter for the whole optimizer (it lives on the device so that a captured CUDA graph advances it on
        # every replay): torch.optim.Adam keeps `step` per parameter, which only differs from a shared counter when a
        # parameter receives its first gradient later than the others — refused here rather than silently mis-corrected
        done = {int(self.state[p]["step"]) if self.state[p] else 0 for p in live}
        if len(done) > 1 and not capturing:
            raise RuntimeError("ngcf_b200 Adam keeps one step count for all parameters, but the parameters that carry "
                               f"gradients have been stepped {sorted(done)} times (a parameter whose first gradient "
                               "arrives late needs torch.optim.Adam)")
        if self._step_dev is None or self._step_dev.device != dev:
            self._step_dev = torch.full((1,), max(done), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            self._step_dev.add_(1)                                    # t of this update: once per step(), not per group
            for group, ps in groups:
                for i in range(0, len(ps), 32):
                    chunk = ps[i:i + 32]
                    sts = [self._state(p) for p in chunk]
                    for p in chunk:
                        if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                            raise RuntimeError("ngcf_b200 Adam needs contiguous fp32 parameters and gradients")
                        if p.device != dev:
                            raise RuntimeError("ngcf_b200 Adam: all parameters must live on one CUDA device")
                    _lib.check(lib.ngcf_adam_step(_lib.ptr_array(chunk), _lib.ptr_array([p.grad for p in chunk]),
                                                  _lib.ptr_array([s["exp_avg"] for s in sts]),
                                                  _lib.ptr_array([s["exp_avg_sq"] for s in sts]),
                                                  _lib.i64_array([p.numel() for p in chunk]), len(chunk),
                                                  float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]),
                                                  float(group["eps"]), float(group["weight_decay"]), 0,
                                                  self._step_dev.data_ptr(), int(zero_grads),
                                                  _lib.current_stream()), "adam_step")
        _lib.bump_param_epoch()                                       # forwards still waiting for their backward are stale
        if not capturing:
            for p in live:
                self.state[p]["step"] += 1
        return loss

    def state_dict(self):
        # under CUDA-graph replay the per-parameter host counters are not advanced: refresh them from the device
        if self._step_dev is not None:
            t = float(self._step_dev.item())
            for st in self.state.values():
                if st:
                    st["step"] = torch.tensor(t, dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._step_dev = None
