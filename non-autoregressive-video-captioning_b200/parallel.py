"""Data parallelism for the hot path: one process per GPU, videos sharded across ranks.

The reference is single-process / single-device.  The path shards on
independent units (videos), so

* inference: every rank decodes its own slice of the batch -- NO data-path collective
  (``shard`` / ``shard_batch``; an optional ``gather_hypotheses`` brings the ids to rank 0);
* training: every rank runs forward/backward on its slice (BatchNorm statistics and dropout
  streams stay rank-local exactly as in plain DDP) and the gradients are averaged
  with ONE all-reduce over a flat fp32 buffer per step (``GradientAllReduce``).  The loss of the
  reference is a token sum divided by the LOCAL batch size (misc/crit.py:40-46), so the mean of
  the per-rank gradients over EQUAL shards is the global-batch gradient (``shard_bounds`` gives the
  first ``n % world`` ranks one extra unit when the batch does not divide: pass
  ``GradientAllReduce.allreduce(weight=local_batch / global_batch * world)`` to weight the ranks).

``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


ALIGN = 64  # floats


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_bounds(n: int, rank_: Optional[int] = None, world_: Optional[int] = None):
    """Contiguous, balanced split of n units: the first n % world ranks get one extra."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    base, extra = divmod(n, w)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def shard(t, rank_: Optional[int] = None, world_: Optional[int] = None):
    """Slice dim 0 (videos) of a tensor, or of every tensor in a list / tuple / dict."""
    if isinstance(t, dict):
        return {k: shard(v, rank_, world_) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        if t and isinstance(t[0], str):  # e.g. video_ids
            lo, hi = shard_bounds(len(t), rank_, world_)
            return type(t)(t[lo:hi])
        return type(t)(shard(v, rank_, world_) for v in t)
    if torch.is_tensor(t) and t.dim() > 0:
        lo, hi = shard_bounds(t.shape[0], rank_, world_)
        return t[lo:hi]
    return t


def broadcast_model(model: torch.nn.Module, src: int = 0):
    """Make every rank start from rank `src`'s parameters and buffers (done once)."""
    if world() == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


def gather_hypotheses(hyp: torch.Tensor, pad: int = 0) -> Optional[List[torch.Tensor]]:
    """Optional: collect the per-rank [B_r, S_r] id tensors on rank 0 (shapes may differ per rank)."""
    if world() == 1:
        return [hyp]
    out = [None] * world() if rank() == 0 else None
    dist.gather_object(hyp.cpu(), out, dst=0)
    return out


class GradientAllReduce:
    """Flat fp32 gradient buffer + ONE all-reduce per step.

    ``p.grad`` of every trainable parameter is a view into ``self.flat`` so autograd accumulates
    straight into the buffer the collective runs on (no pack / unpack copies)::

        dp = GradientAllReduce(model)          # broadcasts weights from rank 0
        for batch in loader:                    # each rank: its own shard
            dp.zero_grad()
            loss = crit(model(**shard(batch)))
            loss.backward()
            dp.allreduce()                      # mean over ranks, one collective
            clip_grad_value_(model.parameters(), 5); optimizer.step()
    """

    def __init__(self, model: torch.nn.Module, broadcast: bool = True):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("model has no trainable parameters")
        dev = self.params[0].device
        # every tensor starts on a 256-byte boundary of the flat buffer (the vectorised / TMA kernels
        # need 16-byte aligned operands); the padding stays zero and rides along in the collective.
        # Parameters that the engine packs into ONE GEMM operand (query | key | value of a layer; the text -> video
        # key | value of all layers) are laid out next to each other in that operand's row order, so that the weight /
        # bias gradient of the packed GEMM is ONE view of the buffer and the kernels accumulate into it directly
        # (otherwise: a zero-filled temporary + one AccumulateGrad add per parameter, ~150 tiny launches per step)
        groups = self._packed_groups(model, dev)
        member = {id(q): g for g in groups for q in g}
        placed, layout = set(), []
        for p in self.params:
            if id(p) in placed:
                continue
            for q in member.get(id(p), [p]):
                layout.append(q)
                placed.add(id(q))
        off_of, off = {}, 0
        for p in layout:
            off_of[id(p)] = off
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.offsets = [off_of[id(p)] for p in self.params]
        self.numel = off
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        self._group_views = {}
        for g in groups:
            o0 = off_of[id(g[0])]
            n = sum(q.numel() for q in g)
            tail = g[0].shape[1:]
            self._group_views[tuple(id(q) for q in g)] = self.flat[o0:o0 + n].view((-1,) + tuple(tail))
        if broadcast:
            broadcast_model(model)
        self.attach()
        self._view_of = {id(p): v for p, v in zip(self.params, self.views)}
        eng = getattr(model, "engine", None)
        if eng is not None:
            eng.grad_sink = self.sink
            eng.grad_sink_group = self.sink_group
            eng.grads_final_hook = self.begin_early   # called by the encoder's backward before it launches anything
        # all-reduce the finished part of the buffer under the encoder's backward ($NAVC_DP_OVERLAP=0: one collective)
        self.overlap = os.environ.get("NAVC_DP_OVERLAP", "1") not in ("0", "no", "off")
        self.weight = 1.0         # this rank's weight (unequal shards), set BEFORE backward when overlap is on
        self._early = None        # (split, work) of the collective in flight
        self._split_of = {}

    def _packed_groups(self, model, dev):
        """Parameter groups behind the engine's concatenated GEMM operands (PackedLinear.src with several entries), each as
        two lists -- the weights and the biases -- in row order.  Only groups whose members are all trainable, distinct,
        of ALIGN-multiple size (no padding between them) and in no other group qualify."""
        eng = getattr(model, "engine", None)
        if eng is None or dev.type != "cuda":
            return []
        from .engine import PackedLinear
        eng.sync_weights()
        named = eng.named_params()
        trainable = set(id(p) for p in self.params)
        found, seen = [], set()

        def walk(o):
            if isinstance(o, PackedLinear):
                if len(o.src) > 1:
                    for col in (0, 1):
                        keys = [s_[col] for s_ in o.src]
                        ps = [named.get(k) if k is not None else None for k in keys]
                        ids = [id(q) for q in ps]
                        if any(q is None for q in ps) or len(set(ids)) != len(ids) or any(i in seen or i not in trainable for i in ids):
                            continue
                        if any(q.numel() % ALIGN for q in ps) or any(q.shape[1:] != ps[0].shape[1:] for q in ps):
                            continue
                        rows = [(s_[2], s_[3]) for s_ in o.src]
                        if rows[0][0] != 0 or any(a[1] != b[0] for a, b in zip(rows, rows[1:])) or \
                                any(q.shape[0] != r1 - r0 for q, (r0, r1) in zip(ps, rows)):
                            continue
                        seen.update(ids)
                        found.append(ps)
            elif isinstance(o, dict):
                for v in o.values():
                    walk(v)
            elif isinstance(o, (list, tuple)):
                for v in o:
                    walk(v)

        walk(eng.P)
        return found

    def sink_group(self, ps):
        """The ONE flat-buffer view behind the gradients of parameters ``ps`` (the members of a packed GEMM operand, in row
        order), while every p.grad still is its own view of the buffer; else None."""
        v = self._group_views.get(tuple(id(q) for q in ps))
        if v is None or any(q.grad is not self._view_of.get(id(q)) for q in ps):
            return None
        return v

    def sink(self, p):
        """The flat-buffer view a backward kernel may accumulate into directly: only while p.grad IS that view
        (zero_grad() / attach() keep it so; an optimizer's zero_grad(set_to_none=True) detaches it)."""
        v = self._view_of.get(id(p))
        return v if (v is not None and p.grad is v) else None

    def attach(self):
        """(Re)point p.grad at the flat buffer (optimizers' zero_grad(set_to_none=True) detaches it)."""
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                if p.grad is None:
                    v.zero_()   # detached by zero_grad(set_to_none=True): no gradient this step, not last step's
                elif p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad)
                p.grad = v

    def zero_grad(self):
        if self._early is not None:   # a step was abandoned between backward() and allreduce()
            self._early[1].wait()
            self._early = None
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def _reduce(self, buf, async_op=False):
        if dist.get_backend() == "nccl":
            return dist.all_reduce(buf, op=dist.ReduceOp.AVG, async_op=async_op)  # sum and 1/world inside the collective
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=async_op)

    def begin_early(self, outstanding: Sequence[torch.nn.Parameter]):
        """Overlap: called (by the encoder's backward, the last autograd node of a step that produces parameter
        gradients) when only the gradients of ``outstanding`` are still to come.  Everything behind the last of
        those parameters in the flat buffer is final, so its all-reduce is launched now (asynchronously, on the
        process group's stream, ordered after the kernels already queued) and runs under the rest of the backward
        pass; ``allreduce()`` then reduces the head of the buffer and joins both."""
        if not self.overlap or world() == 1 or self._early is not None:
            return
        key = tuple(id(p) for p in outstanding)
        split = self._split_of.get(key)
        if split is None:
            ids = set(key)
            split = 0
            for p, o in zip(self.params, self.offsets):
                if id(p) in ids:
                    split = max(split, o + (p.numel() + ALIGN - 1) // ALIGN * ALIGN)
            self._split_of[key] = split
        if split >= self.numel or any(p.grad is not v for p, v in zip(self.params, self.views)):
            return   # nothing behind the outstanding parameters, or some p.grad is not (yet) a view of the buffer
        tail = self.flat[split:]
        if self.weight != 1.0:
            tail.mul_(float(self.weight))
        self._early = (split, self._reduce(tail, async_op=True))

    def allreduce(self, weight: Optional[float] = None):
        """Average the gradients over the ranks: one collective on the whole buffer (two when the tail of the buffer
        went out early, ``begin_early``).  ``weight`` scales this rank's gradients first (unequal shards:
        local_batch * world / global_batch); with overlap on, set ``self.weight`` before the backward pass instead."""
        self.attach()
        w = world()
        weight = self.weight if weight is None else float(weight)
        early, self._early = self._early, None
        if early is not None and weight != self.weight:
            raise RuntimeError("GradientAllReduce: part of the buffer was reduced early with weight %g; set .weight before "
                               "backward() (or .overlap = False) instead of passing weight=%g here" % (self.weight, weight))
        head = self.flat if early is None else self.flat[:early[0]]
        if weight != 1.0:
            head.mul_(weight)
        if w == 1:
            return self.flat
        self._reduce(head)
        if early is not None:
            early[1].wait()
        if dist.get_backend() != "nccl":
            self.flat.div_(w)
        return self.flat

    @property
    def nbytes(self) -> int:
        return self.numel * 4
