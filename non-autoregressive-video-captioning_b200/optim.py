"""Fused optimizer step with the reference's optimizer API (contract: reference misc/optim.py:3-68 and
the clip + step pair of misc/run.py:260-261).

``get_optimizer(opt, model)`` returns a ``ScheduledOptim`` with the reference's methods (``step``,
``zero_grad``, ``epoch_update_learning_rate``, ``step_update_learning_rate``, ``get_lr``).  Its
``step()`` is ONE kernel launch (``navc_clip_adam``) over flat fp32 buffers: clip_grad_value_,
L2 weight decay, Adam moments, bias correction and the parameter update.  Parameters and gradients
are re-homed as views of two flat buffers; with data parallelism the gradient buffer is the very
buffer the single NCCL all-reduce runs on (``parallel.GradientAllReduce``).
"""
from __future__ import annotations

import torch

from . import _lib as L
from .parallel import GradientAllReduce


class FusedClipAdam:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_clip=0.0,
                 grads: GradientAllReduce = None):
        self.model = model
        self.grads = grads if grads is not None else GradientAllReduce(model, broadcast=False)
        self.params = self.grads.params
        dev = self.params[0].device
        L.ensure_init(dev)
        n = self.grads.numel
        # parameters become views of one flat buffer (same values)
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)   # same (256-byte aligned) layout as the gradients
        with torch.no_grad():
            for p, off in zip(self.params, self.grads.offsets):
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.lr, self.betas, self.eps, self.weight_decay, self.grad_clip = lr, betas, eps, weight_decay, grad_clip
        self.t = 0
        self.param_groups = [{"lr": lr, "params": self.params}]  # what misc/optim.py:45-46 writes to

    def zero_grad(self, set_to_none=False):
        self.grads.zero_grad()

    def step(self):
        self.grads.attach()
        self.t += 1
        lr = float(self.param_groups[0]["lr"])
        L.call("navc_clip_adam", L.ptr(self.flat_p), L.ptr(self.grads.flat), L.ptr(self.m), L.ptr(self.v), self.grads.numel,
               float(self.grad_clip or 0.0), lr, float(self.betas[0]), float(self.betas[1]), float(self.eps),
               float(self.weight_decay), self.t, L.stream())
        eng = getattr(self.model, "engine", None)
        if eng is not None:
            eng.invalidate()  # the kernel wrote the parameters behind autograd's version counters

    def state_dict(self):
        return {"t": self.t, "m": self.m, "v": self.v, "lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, sd):
        self.t = int(sd["t"])
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.param_groups[0]["lr"] = sd["lr"]


class ScheduledOptim:
    """reference misc/optim.py:3-49 (per-step warm-up ratio, per-epoch decay with a floor)."""

    def __init__(self, optimizer, learning_rate, minimum_learning_rate, epoch_decay_rate, grad_clip=2, n_warmup_steps=0,
                 summarywriter=None):
        self._optimizer = optimizer
        self.n_current_steps = 0
        self.lr = learning_rate
        self.mlr = minimum_learning_rate
        self.decay = epoch_decay_rate
        self.grad_clip = grad_clip
        self.n_warmup_steps = n_warmup_steps
        self.summarywriter = summarywriter

    def step(self):
        self.step_update_learning_rate()
        self._optimizer.step()

    def zero_grad(self):
        self._optimizer.zero_grad()

    def epoch_update_learning_rate(self):
        if self.n_current_steps > self.n_warmup_steps:
            self.lr = max(self.mlr, self.decay * self.lr)

    def step_update_learning_rate(self):
        self.n_current_steps += 1
        ratio = min(self.n_current_steps / (self.n_warmup_steps + 1.0), 1)
        learning_rate = self.lr * ratio
        if self.summarywriter is not None:
            self.summarywriter.add_scalar("learning_rate", learning_rate, global_step=self.n_current_steps)
        for group in self._optimizer.param_groups:
            group["lr"] = learning_rate

    def get_lr(self):
        return self.lr


def get_optimizer(opt, model, summarywriter=None, grads: GradientAllReduce = None):
    """reference misc/optim.py:51-68 (``optim: adam``); clipping (opt['grad_clip'], misc/run.py:260) is fused in."""
    if opt.get("optim", "adam").lower() != "adam":
        raise NotImplementedError("fused optimizer implements Adam (the method presets' optimizer)")
    inner = FusedClipAdam(model, lr=opt["learning_rate"], weight_decay=opt.get("weight_decay", 0.0),
                          grad_clip=opt.get("grad_clip", 0.0), grads=grads)
    return ScheduledOptim(inner, learning_rate=opt["learning_rate"], minimum_learning_rate=opt["minimum_learning_rate"],
                          epoch_decay_rate=opt["decay"], grad_clip=opt.get("grad_clip", 2),
                          n_warmup_steps=opt.get("n_warmup_steps", 0), summarywriter=summarywriter)
