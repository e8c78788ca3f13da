"""Training-mode (autograd) execution of the hot path: forward in train mode and a hand-written
backward, both through the C ABI of ``include/navc.h``.

The reference trains with ``loss.backward()`` through eager PyTorch modules (misc/run.py:254-261).
Here the three boundary calls of ``Seq2Seq`` -- ``encode``, ``decoder`` and ``tgt_word_prj`` (+
``log_softmax``) -- are each ONE ``torch.autograd.Function`` whose forward and backward are
sequences of navc kernels; autograd only stitches the three together, accumulates parameter
gradients into ``.grad`` and runs the caller's loss (misc/crit.py) on the returned log-probs.

* Dropout (p=0.5 at every site of the reference, incl. the double dropout of BertOutput,
  models/bert.py:240-247) uses a counter-based hash: masks are regenerated in the backward pass
  from per-site seeds drawn from torch's CPU generator (``torch.manual_seed`` reproducible), never
  stored.  torch's own Philox stream cannot be reproduced bit-exactly, so parity
  with the reference is tested with dropout disabled and statistically otherwise.
* BatchNorm uses batch statistics and updates ``running_mean/var/num_batches_tracked`` like
  ``nn.BatchNorm1d`` in train mode (models/joint_representation.py:43-45).
* Weight / input gradients are tensor-core GEMMs on transposed operand copies
  (``navc_transpose_pack``): dX = dY W, dW = dY^T X (split-K, atomic accumulate), db = column sums.
"""
from __future__ import annotations

import math
import os

from typing import Dict, List, Optional

import torch

from . import _lib as L
from .config import Constants
from .engine import Act, Engine, PackedLinear

_GOLD = 0x9E3779B97F4A7C15
_MASK64 = (1 << 64) - 1


def _up(x, m):
    return (x + m - 1) // m * m


class Seeds:
    """Per-site dropout seeds of one forward call (base from torch's CPU generator)."""

    def __init__(self):
        self.base = int(torch.randint(0, 2 ** 62, (1,)).item())
        self.n = 0

    def next(self) -> int:
        self.n += 1
        return (self.base + self.n * _GOLD) & _MASK64


# --------------------------------------------------------------------------------------------------
# GEMM helpers
# --------------------------------------------------------------------------------------------------
class Operand:
    """A row-major [rows, K] GEMM operand in the engine's format: fp32 (fp32 mode) or bf16 hi(/lo)."""
    __slots__ = ("f32", "hi", "lo", "rows", "K", "ld")

    def __init__(self, rows, K, ld, f32=None, hi=None, lo=None):
        self.rows, self.K, self.ld, self.f32, self.hi, self.lo = rows, K, ld, f32, hi, lo


def _alloc_operand(eng: Engine, rows, ld) -> Operand:
    dev = eng.device
    if eng.tc:
        hi = torch.empty((rows, ld), dtype=torch.bfloat16, device=dev)
        lo = torch.empty((rows, ld), dtype=torch.bfloat16, device=dev) if eng.split else None
        return Operand(rows, ld, ld, hi=hi, lo=lo)
    return Operand(rows, ld, ld, f32=torch.empty((rows, ld), dtype=torch.float32, device=dev))


def gemm(eng: Engine, x: Operand, w: Operand, M, N, K, out: torch.Tensor, ld_out, bias=None, residual=None,
         ld_res=0, accumulate=False, split_k=1):
    """out[M,N] (+)= x[M,K] w[N,K]^T (+bias +residual) in the engine's precision mode."""
    ep = L.Epilogue(L.ptr(bias), L.ptr(residual), None, 0, ld_res, L.ptr(out), None, None, ld_out, 0,
                    split_k if eng.tc else 1, 1 if accumulate else 0)
    if eng.tc:
        L.call("navc_linear_tc", eng.tc_mode, L.ptr(x.hi), L.ptr(x.lo), x.ld, L.ptr(w.hi), L.ptr(w.lo), w.ld, M, N, K, ep,
               L.stream())
    else:
        L.call("navc_linear_f32", L.ptr(x.f32), x.ld, L.ptr(w.f32), w.ld, M, N, K, ep, L.stream())


def transpose_pack(eng: Engine, x: torch.Tensor, M, N, ld, straight=False, transposed=False, colsum=None):
    """x [M,N] fp32 -> (straight operand [M, N] ld=ld, transposed operand [N, Mp]) in the engine's format."""
    s = t = None
    Mp = _up(M, 64)
    if straight:
        assert ld % 8 == 0
        s = Operand(M, N, ld, f32=x) if not eng.tc else _alloc_operand(eng, M, ld)
        if eng.tc:
            s.K = N
    if transposed:
        t = _alloc_operand(eng, N, Mp)
    need_kernel = (eng.tc and straight) or transposed or colsum is not None
    if need_kernel:
        L.call("navc_transpose_pack", L.ptr(x), M, N, ld,
               L.ptr(s.hi) if (s is not None and eng.tc) else None, L.ptr(s.lo) if (s is not None and eng.tc) else None, ld,
               L.ptr(t.f32) if t is not None else None, L.ptr(t.hi) if t is not None else None,
               L.ptr(t.lo) if t is not None else None, Mp, L.ptr(colsum), L.stream())
    return s, t


def weight_T(eng: Engine, lin: PackedLinear) -> Operand:
    """W^T [K, Np] (Np = N rounded up to 64, zero padded) for dX = dY W, cached per packed weight."""
    if lin.T is None:
        _, lin.T = transpose_pack(eng, lin.w, lin.N, lin.K, lin.K, transposed=True)
    return lin.T


# dX = dY W on the row-major weights (navc_dgrad_tc); NAVC_DGRAD_T=1 restores the transposed-weight GEMM
DGRAD_MN = os.environ.get("NAVC_DGRAD_T", "0") in ("0", "", "false")


def _split_k(out_rows, out_cols, k):
    tiles = ((out_rows + 127) // 128) * ((out_cols + 255) // 256)
    kb = (k + 63) // 64
    return max(1, min(kb, (2 * 148) // max(tiles, 1)))


def _check_pack(eng, st):
    """The backward kernels read the packed weights (eng.P) of the forward pass: a repack in between (an optimizer
    step or load_state_dict between forward and backward) would silently mix weight versions."""
    if st.get("pack_id", eng.pack_id) != eng.pack_id:
        raise RuntimeError("navc: the model's weights changed between forward and backward (packed weights %d -> %d)"
                           % (st["pack_id"], eng.pack_id))


class Grads:
    """Gradient accumulator keyed by state_dict name."""

    def __init__(self, eng=None):
        self.g: Dict[str, torch.Tensor] = {}
        # direct accumulation (parallel.GradientAllReduce owns p.grad as views of one flat buffer): kernels that
        # accumulate anyway (split-K weight gradients, bias column sums) add straight into p.grad and autograd gets
        # None for that parameter -- no zero-filled temporary, no AccumulateGrad add
        self.sink = getattr(eng, "grad_sink", None) if eng is not None else None
        self.sink_group = getattr(eng, "grad_sink_group", None) if eng is not None else None
        self.named = eng.named_params() if self.sink is not None else None

    def target(self, key):
        if self.sink is None or key is None:
            return None
        p = self.named.get(key)
        return None if p is None else self.sink(p)

    def target_group(self, keys):
        """One gradient view for the parameters behind a concatenated GEMM operand (parallel.GradientAllReduce lays them
        out adjacently), or None."""
        if self.sink_group is None or any(k is None for k in keys):
            return None
        ps = [self.named.get(k) for k in keys]
        return None if any(q is None for q in ps) else self.sink_group(ps)

    def add(self, key, t):
        if key is None:
            return
        self.g[key] = t if key not in self.g else self.g[key] + t

    def add_packed(self, lin: PackedLinear, dW, db):
        for wk, bk, r0, r1 in lin.src:
            if dW is not None:
                self.add(wk, dW[r0:r1] if (r0, r1) != (0, dW.shape[0]) else dW)
            if bk is not None and db is not None:
                self.add(bk, db[r0:r1] if (r0, r1) != (0, db.shape[0]) else db)


def lin_bwd(eng: Engine, x, lin: PackedLinear, dY: torch.Tensor, grads: Grads, need_dx=True,
            dx_residual: Optional[torch.Tensor] = None, ld_dy=None, dy_split: Optional[Operand] = None, db_pre=None):
    """Backward of y = x W^T + b.  x: the forward input, an fp32 tensor [M,K] or the Act the forward GEMM
    consumed (its bf16 hi/lo copies are reused); dY [M,N] fp32 (leading dimension ld_dy, pad columns zero).
    Accumulates dW / db into ``grads``; returns dX [M,K] (+ dx_residual) or None.

    Tensor-core modes: dY is split once (hi/lo, + column sums = db); dW = dY^T X runs straight on the
    row-major operands (``navc_wgrad_tc``, MN-major tcgen05 operands, split-K) and dX = dY W on the
    pre-transposed weights -- no transposed activation copies."""
    x_act = x if isinstance(x, Act) else None
    x32 = x.f32 if x_act is not None else x
    M, K = (x_act.M, x_act.N) if x_act is not None else x.shape
    N = lin.N
    dev = eng.device
    # dy_split: the producer already wrote dY as the bf16 hi (/ lo) operand (tensor-core path only) and accumulated its
    # column sums into db_pre = (bias-gradient tensor, whether that tensor IS the parameters' gradient view)
    ld = dy_split.ld if dy_split is not None else (ld_dy or dY.stride(0))
    if dy_split is None and (ld % 16 != 0 or (need_dx and ld < _up(N, 64)) or (eng.tc and N % 8 != 0 and ld < _up(N, 64))):
        # operand alignment: re-lay dY with a zero-padded leading dimension
        ldp = _up(N, 64)
        pad = torch.zeros((M, ldp), dtype=torch.float32, device=dev)
        pad[:, :N].copy_(dY.view(M, -1)[:, :N] if ld == dY.stride(0) else dY.as_strided((M, N), (ld, 1)))
        dY, ld = pad, ldp
    tw = tb = None
    if eng.tc and K % 8 == 0 and N % 8 == 0 and len(lin.src) == 1:  # one parameter behind this GEMM: accumulate in place
        tw = grads.target(lin.src[0][0])
        tb = grads.target(lin.src[0][1]) if lin.b is not None else None
    elif eng.tc and K % 8 == 0 and N % 8 == 0:  # several parameters, adjacent in the flat gradient buffer: one view
        tw = grads.target_group([s_[0] for s_ in lin.src])
        tb = grads.target_group([s_[1] for s_ in lin.src]) if lin.b is not None else None
        if tw is not None and tuple(tw.shape) != (N, K):
            tw = None
        if tb is not None and tb.numel() != N:
            tb = None
    if db_pre is not None:
        db, tb = db_pre[0], (db_pre[0] if db_pre[1] else None)
    else:
        db = tb if tb is not None else (torch.zeros((N,), dtype=torch.float32, device=dev) if lin.b is not None else None)
    if eng.tc and K % 8 == 0:
        if dy_split is not None:
            s = dy_split
        else:
            s, _ = transpose_pack(eng, dY, M, N, ld, straight=True, colsum=db)
        if x_act is not None and x_act.hi is not None and (x_act.lo is not None or not eng.split):
            x_hi, x_lo = x_act.hi, x_act.lo
        else:
            xs, _ = transpose_pack(eng, x32, M, K, x32.stride(0), straight=True)
            x_hi, x_lo = xs.hi, xs.lo
        n_eff = N if N % 8 == 0 else ld       # pad columns of dY are zero -> zero rows of dW, sliced off below
        dW = tw if tw is not None else torch.zeros((n_eff, K), dtype=torch.float32, device=dev)
        ep = L.Epilogue(None, None, None, 0, 0, L.ptr(dW), None, None, K, 0, _split_k(n_eff, K, M), 1)
        L.call("navc_wgrad_tc", eng.tc_mode, L.ptr(s.hi), L.ptr(s.lo), ld, L.ptr(x_hi), L.ptr(x_lo), K, M, n_eff, K, ep, L.stream())
        # hand over only what did not go straight into p.grad
        grads.add_packed(lin, None if tw is not None else (dW[:N] if n_eff != N else dW), None if tb is not None else db)
    else:
        s, t = transpose_pack(eng, dY, M, N, ld, straight=need_dx, transposed=True, colsum=db)
        _, xt = transpose_pack(eng, x32, M, K, x32.stride(0), transposed=True)
        Mp = _up(M, 64)
        dW = torch.zeros((N, K), dtype=torch.float32, device=dev)
        gemm(eng, t, xt, N, K, Mp, dW, K, accumulate=True, split_k=_split_k(N, K, Mp))
        grads.add_packed(lin, dW, db)
    if not need_dx:
        return None
    dX = torch.empty((M, K), dtype=torch.float32, device=dev)
    if eng.tc and K % 8 == 0 and lin.w_hi is not None and (lin.w_lo is not None or not eng.split) and DGRAD_MN:
        # dX = dY W straight from the forward's row-major bf16 weight copies (B operand consumed MN-major)
        ep = L.Epilogue(None, L.ptr(dx_residual), None, 0, K if dx_residual is not None else 0, L.ptr(dX), None, None, K, 0, 1, 0)
        L.call("navc_dgrad_tc", eng.tc_mode, L.ptr(s.hi), L.ptr(s.lo), ld, L.ptr(lin.w_hi), L.ptr(lin.w_lo), lin.K, M, N, K, ep,
               L.stream())
        return dX
    wt = weight_T(eng, lin)  # [K, Np]
    gemm(eng, s, wt, M, K, min(ld, wt.ld), dX, K, residual=dx_residual, ld_res=K if dx_residual is not None else 0)
    return dX


# --------------------------------------------------------------------------------------------------
# elementwise wrappers
# --------------------------------------------------------------------------------------------------
def drop_add(eng, y, res, s1, p1, s2, p2, row_tokens, f32=True, bf=True) -> Act:
    M, D = y.shape
    out = eng._new(M, D, f32 or not eng.tc, bf)
    L.call("navc_drop_add", L.ptr(y), L.ptr(res), s1, p1, s2, p2, L.ptr(row_tokens), M, D, L.ptr(out.f32), L.ptr(out.hi),
           L.ptr(out.lo), L.stream())
    return out


def drop_add_bwd(dout, s1, p1, s2, p2, row_tokens, want_res=True):
    M, D = dout.shape
    d_y = torch.empty_like(dout)
    d_res = torch.empty_like(dout) if (want_res and p1 > 0) else None
    L.call("navc_drop_add_bwd", L.ptr(dout), s1, p1, s2, p2, L.ptr(row_tokens), M, D, L.ptr(d_y), L.ptr(d_res), L.stream())
    if want_res and d_res is None:
        d_res = d_y  # no first dropout: the two gradients coincide
    return d_y, d_res


FOLD_SPLIT = os.environ.get("NAVC_FOLD_SPLIT", "1") not in ("0", "no", "off")


def _dy_target(eng, lin: PackedLinear, grads: Grads, M):
    """Buffers for a producer that writes the dY of ``lin`` directly in operand form (bf16 hi / lo + column sums):
    (Operand, bias-gradient tensor, whether that tensor is the parameter's gradient view), or None if the tensor-core
    gradient path / the alignment rules do not apply."""
    N = lin.N
    if not (FOLD_SPLIT and eng.tc and lin.K % 8 == 0 and N % 64 == 0 and lin.b is not None):
        return None
    if len(lin.src) == 1:
        tb = grads.target(lin.src[0][1])
    else:
        tb = grads.target_group([s_[1] for s_ in lin.src])
        if tb is not None and tb.numel() != N:
            tb = None
    db = tb if tb is not None else torch.zeros((N,), dtype=torch.float32, device=eng.device)
    op = _alloc_operand(eng, M, N)
    op.K = N
    return op, db, tb is not None


def _post_bwd_split(eng, g, s1, p1, s2, p2, tok_flat, sp):
    """_post_bwd without LayerNorm whose d_dense_out leaves as the operand sp (navc_drop_add_bwd_split); returns d_residual."""
    M, D = g.shape
    d_res = torch.empty_like(g)
    op, db, _ = sp
    L.call("navc_drop_add_bwd_split", L.ptr(g), s1, p1, s2, p2, L.ptr(tok_flat), M, D, L.ptr(d_res), L.ptr(op.hi), L.ptr(op.lo), op.ld,
           L.ptr(db), L.stream())
    return d_res


def _post(eng, y, res, ln, s1, p1, s2, p2, tok_flat, saved):
    """dense output -> dropout -> +residual -> [LayerNorm] -> [dropout] -> * non_pad_mask
    (models/bert.py:193-200 with p2 == 0; bert.py:241-247 with the second dropout)."""
    if ln is None:
        return drop_add(eng, y, res, s1, p1, s2, p2, tok_flat)
    t = drop_add(eng, y, res, s1, p1, 0, 0.0, None, f32=True, bf=False)
    saved["t"] = t.f32
    if p2 > 0:
        z = eng.layernorm(t, ln, None, f32=True, bf=False)
        return drop_add(eng, z.f32, None, s2, p2, 0, 0.0, tok_flat)
    return eng.layernorm(t, ln, tok_flat)


def _post_bwd(eng, g, ln, ln_key, s1, p1, s2, p2, tok_flat, saved, grads: Grads):
    """Returns (d_dense_out, d_residual)."""
    if ln is None:
        return drop_add_bwd(g, s1, p1, s2, p2, tok_flat)
    M, D = g.shape
    if p2 > 0:
        g, _ = drop_add_bwd(g, s2, p2, 0, 0.0, tok_flat, want_res=False)
        tok_for_ln = None
    else:
        tok_for_ln = tok_flat
    dx = torch.empty_like(g)
    dw = torch.zeros((D,), dtype=torch.float32, device=g.device)
    db = torch.zeros((D,), dtype=torch.float32, device=g.device)
    L.call("navc_layernorm_bwd", L.ptr(g), L.ptr(saved["t"]), L.ptr(ln[0]), eng.eps, L.ptr(tok_for_ln), M, D, L.ptr(dx),
           L.ptr(dw), L.ptr(db), L.stream())
    grads.add(ln_key + ".weight", dw)
    grads.add(ln_key + ".bias", db)
    d_y, d_res = drop_add_bwd(dx, s1, p1, 0, 0.0, None)
    return d_y, d_res


def _lin_f32(eng: Engine, w, b, src) -> PackedLinear:
    """Small head GEMMs (length predictor) always run on the fp32 CUDA-core path."""
    return PackedLinear(w, b, src)


def _gemm_f32(x, w, b, M, N, K, out):
    ep = L.Epilogue(L.ptr(b), None, None, 0, 0, L.ptr(out), None, None, N, 0, 1, 0)
    L.call("navc_linear_f32", L.ptr(x), K, L.ptr(w), K, M, N, K, ep, L.stream())


class _F32Engine:
    """View of an engine that forces the fp32 CUDA-core GEMM path (tiny head GEMMs)."""

    def __init__(self, eng):
        self.device, self.tc, self.precision, self.tc_mode = eng.device, False, "fp32", 0
        self.split = self.tf32 = False


# --------------------------------------------------------------------------------------------------
# encoder  (models/seq2seq.py:35-63 in train mode)
# --------------------------------------------------------------------------------------------------
def _param_list(model, keys):
    sd = model.engine.named_params()
    return [sd[k] for k in keys]


def _encode_param_keys(model):
    return [k for k in model.engine.named_params()
            if k.startswith(("encoder.", "joint_representation_learner.", "auxiliary_task_predictor."))]


class EncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, keys, n_feats, *tensors):
        eng: Engine = model.engine
        eng.sync_weights()
        opt, P, D = eng.opt, eng.P, eng.D
        feats = [f.detach() for f in tensors[:n_feats]]
        dev = eng.device
        if opt.get("fusion", "temporal_concat") not in ("temporal_concat", "none"):
            raise NotImplementedError("fusion=%r (broken in the reference, joint_representation.py:41)" % opt["fusion"])
        B, F_ = feats[0].shape[0], feats[0].shape[1]
        if any(f.shape[1] != F_ for f in feats):
            raise NotImplementedError("modalities with different frame counts")
        E = F_ * len(feats)
        BF = B * F_
        seeds = Seeds()
        p_enc = float(opt.get("encoder_dropout", 0.5))
        p_hid = float(opt["hidden_dropout_prob"])
        no_norm = opt.get("fusion", "temporal_concat") == "none" or opt.get("no_encoder_bn", False)
        enc = Act(B * E, D, f32=torch.empty((B, E, D), dtype=torch.float32, device=dev))
        enc_hidden = torch.empty((B, D), dtype=torch.float32, device=dev)
        streams = []
        jr = model.joint_representation_learner
        for i, (f, st) in enumerate(zip(feats, P["streams"])):
            x_in = eng.from_f32(f.reshape(BF, f.shape[2]).to(dev))
            x = eng.linear(x_in, st["l0"], f32=True, bf=True)
            yg = eng.linear(x, st["l12"], f32=True, bf=False)
            o = torch.empty((BF, D), dtype=torch.float32, device=dev)
            seed = seeds.next()
            L.call("navc_highway_fwd_train", L.ptr(x.f32), L.ptr(yg.f32), st["gate"], BF, D, seed, p_enc, L.ptr(o), L.stream())
            norm = None if no_norm else P["norms"][i]
            mean = var = bw = bb = None
            if norm is not None:
                if norm[0] == "ln":
                    raise NotImplementedError("norm_type='ln' encoder norm")
                bn = getattr(jr, "bn%d" % i)
                bw, bb = norm[3], norm[4]
                mean = torch.empty((D,), dtype=torch.float32, device=dev)
                var = torch.empty((D,), dtype=torch.float32, device=dev)
                L.call("navc_bn_stats", L.ptr(o), BF, D, L.ptr(mean), L.ptr(var), float(bn.momentum if bn.momentum is not None else 0.1),
                       L.ptr(bn.running_mean), L.ptr(bn.running_var), L.stream())
                bn.num_batches_tracked += 1
            L.call("navc_bn_apply_concat", L.ptr(o), L.ptr(mean), L.ptr(var), L.ptr(bw), L.ptr(bb), 1e-5, B, F_, D, E, i,
                   len(feats), int(i > 0), L.ptr(enc_hidden), L.ptr(enc.f32), None, None, L.stream())
            streams.append(dict(x_in=x_in, x=x, yg=yg.f32, o=o, mean=mean, var=var, bw=bw, seed=seed))
        outs = [enc.f32, enc_hidden]
        head = None
        if P["len_head"] is not None:
            w1, b1, w2, b2 = P["len_head"]
            max_len = w2.shape[0]
            m = torch.empty((B, D), dtype=torch.float32, device=dev)
            L.call("navc_length_head", L.ptr(enc.f32), B, E, D, None, None, None, None, 0, L.ptr(m), None, L.stream())
            h_pre = torch.empty((B, D), dtype=torch.float32, device=dev)
            _gemm_f32(m, w1, b1, B, D, D, h_pre)
            h = torch.empty_like(h_pre)
            hseed = seeds.next()
            L.call("navc_act_drop", L.ptr(h_pre), L.ACT["relu"], hseed, p_hid, h_pre.numel(), L.ptr(h), None, None, L.stream())
            logits = torch.empty((B, max_len), dtype=torch.float32, device=dev)
            _gemm_f32(h, w2, b2, B, max_len, D, logits)
            pred = torch.empty_like(logits)
            L.call("navc_log_softmax", L.ptr(logits), L.ptr(pred), B, max_len, max_len, L.stream())
            # pred is an OUTPUT: keeping that very object on ctx would close a reference cycle
            # (pred -> grad_fn -> ctx.state -> pred) and park every saved activation until the cyclic GC runs
            head = dict(m=m, h_pre=h_pre, h=h, pred=pred.detach(), seed=hseed, p=p_hid)
            outs.append(pred)
        ctx.model, ctx.keys, ctx.n_feats = model, keys, n_feats
        ctx.state = dict(streams=streams, head=head, B=B, F=F_, E=E, p_enc=p_enc, pack_id=eng.pack_id)
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_enc, d_hidden, *rest):
        model, st = ctx.model, ctx.state
        eng: Engine = model.engine
        _check_pack(eng, st)
        P, D = eng.P, eng.D
        B, F_, E = st["B"], st["F"], st["E"]
        dev = eng.device
        BF = B * F_
        grads = Grads(eng)
        # every consumer of the encoder's outputs has run its backward: the gradients of all other parameters are
        # final, and a data-parallel owner of p.grad may start reducing them under this node (parallel.begin_early)
        hook = getattr(eng, "grads_final_hook", None)
        if hook is not None:
            named = eng.named_params()
            hook([named[k] for k in ctx.keys if k is not None and k in named])
        feng = _F32Engine(eng)
        d_enc = torch.zeros((B, E, D), dtype=torch.float32, device=dev) if d_enc is None else d_enc.contiguous().clone()
        d_hidden = None if d_hidden is None else d_hidden.contiguous()
        head = st["head"]
        d_pred = rest[0] if rest else None
        if head is not None and d_pred is not None:
            w1, b1, w2, b2 = P["len_head"]
            k1w, k1b, k2w, k2b = P["len_head_keys"]
            max_len = w2.shape[0]
            Lp = _up(max_len, 64)
            dlog = torch.empty((B, Lp), dtype=torch.float32, device=dev)
            L.call("navc_log_softmax_bwd", L.ptr(d_pred.contiguous()), L.ptr(head["pred"]), B, max_len, max_len, L.ptr(dlog),
                   Lp, L.stream())
            lin2 = PackedLinear(w2, b2, [(k2w, k2b, 0, max_len)])
            dh = lin_bwd(feng, head["h"], lin2, dlog, grads, ld_dy=Lp)
            dh_pre = torch.empty_like(dh)
            L.call("navc_act_drop_bwd", L.ptr(dh), L.ptr(head["h_pre"]), L.ACT["relu"], head["seed"], head["p"], dh.numel(),
                   L.ptr(dh_pre), L.stream())
            lin1 = PackedLinear(w1, b1, [(k1w, k1b, 0, D)])
            dm = lin_bwd(feng, head["m"], lin1, dh_pre, grads)
            L.call("navc_mean_bwd", L.ptr(dm), B, E, D, L.ptr(d_enc), L.stream())
        n_mod = len(st["streams"])
        for i, (s, pw) in enumerate(zip(st["streams"], P["streams"])):
            d_o = torch.empty((BF, D), dtype=torch.float32, device=dev)
            d_w = d_b = None
            if s["mean"] is not None:
                d_w = torch.empty((D,), dtype=torch.float32, device=dev)
                d_b = torch.empty((D,), dtype=torch.float32, device=dev)
            L.call("navc_bn_bwd", L.ptr(d_enc), L.ptr(d_hidden), L.ptr(s["o"]), L.ptr(s["mean"]), L.ptr(s["var"]), L.ptr(s["bw"]),
                   1e-5, B, F_, D, E, i, n_mod, L.ptr(d_w), L.ptr(d_b), L.ptr(d_o), L.stream())
            if d_w is not None:
                grads.add(P["norm_keys"][i] + ".weight", d_w)
                grads.add(P["norm_keys"][i] + ".bias", d_b)
            gate = pw["gate"]
            d_x = torch.empty((BF, D), dtype=torch.float32, device=dev)
            d_yg = torch.empty((BF, (2 if gate else 1) * D), dtype=torch.float32, device=dev)
            L.call("navc_highway_bwd", L.ptr(d_o), L.ptr(s["x"].f32), L.ptr(s["yg"]), gate, BF, D, s["seed"], st["p_enc"],
                   L.ptr(d_x), L.ptr(d_yg), L.stream())
            d_x = lin_bwd(eng, s["x"], pw["l12"], d_yg, grads, dx_residual=d_x)
            lin_bwd(eng, s["x_in"], pw["l0"], d_x, grads, need_dx=False)
        return (None, None, None) + (None,) * ctx.n_feats + tuple(grads.g.get(k) for k in ctx.keys)


def encode_train(model, feats):
    keys = _encode_param_keys(model)
    feats = list(feats)
    outs = EncodeFn.apply(model, keys, len(feats), *feats, *_param_list(model, keys))
    results = {}
    if len(outs) == 3:
        results[Constants.mapping["length"][0]] = outs[2]
    results["enc_output"] = outs[0]
    results["enc_hidden"] = outs[1]
    return results


# --------------------------------------------------------------------------------------------------
# decoder  (models/Decoder.py:96-178, models/bert.py:262-303 in train mode)
# --------------------------------------------------------------------------------------------------
def train_packed_enabled(opt) -> bool:
    """Packed rows on the training path: opt['navc_train_packed'] or $NAVC_TRAIN_PACKED (default on)."""
    v = opt.get("navc_train_packed", None)
    if v is None:
        v = os.environ.get("NAVC_TRAIN_PACKED", "1")
    return str(v).lower() not in ("0", "false", "no", "off")


def _plan_key(t: torch.Tensor):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t._version)


def plan_packing(eng: Engine, token_sets, E: int, between=None):
    """Packed rows for the token tensors of one training batch: {key: plan or None}, ONE small host read for
    all of them.  PAD positions carry no signal on this path: their keys are masked out of every softmax and
    their rows are zeroed by ``* non_pad_mask`` after every sub-layer (models/bert.py:271-299), so only the
    sum(len) real positions are rows of the decoder / vocabulary GEMMs.  Needs the tcgen05 attention cores
    and PAD only as a suffix (what the reference's loaders produce).  ``between`` runs after the length
    kernels are queued and before the host read (host work that does not depend on the answer)."""
    plans, todo = {}, []
    for t in token_sets:
        N, S = t.shape
        if eng.tc_attention_ok(S, E) and train_packed_enabled(eng.opt):
            nonpad = t.ne(Constants.PAD)
            lens = nonpad.sum(1, dtype=torch.int32)
            suffix = (nonpad == (torch.arange(S, device=t.device).unsqueeze(0) < lens.unsqueeze(1))).all()
            todo.append((t, nonpad, lens, torch.stack([lens.sum().to(torch.int32), suffix.to(torch.int32), lens.min()])))
        else:
            plans[_plan_key(t)] = None
    dev_stats = torch.stack([x[3] for x in todo]) if todo else None
    stats = dev_stats.to("cpu", non_blocking=True) if todo else None
    if between is not None:
        between()
    if not todo:
        return plans
    if dev_stats.is_cuda:
        torch.cuda.current_stream().synchronize()
    for (t, nonpad, lens, _), (total, ok, shortest) in zip(todo, stats.tolist()):  # the host read
        N, S = t.shape
        pk = None
        if ok and shortest >= 1 and total <= 0.92 * N * S:
            pk = eng.pack_rows(lens, S)
            pk["rows"] = int(total)
            pk["rowmap"] = pk["rowmap"][:total]
            # padded row ids that are not packed (size known on the host: no second read)
            pk["pad_rows"] = torch.nonzero_static(~nonpad.reshape(-1), size=N * S - int(total)).to(torch.int32).reshape(-1)
        plans[_plan_key(t)] = pk
    return plans


def plan_packing_ahead(model, token_sets, E: int):
    """Called at the top of the training forward, BEFORE the encoder is launched: the host read then waits
    only for the previous step's tail instead of stalling the launch stream between encoder and decoder, and
    the per-step weight repack is queued while that tail drains."""
    eng: Engine = model.engine
    eng.train_plans = plan_packing(eng, [t for t in token_sets if torch.is_tensor(t) and t.dim() == 2], E,
                                   between=eng.sync_weights)


def _take_plan(eng: Engine, tokens: torch.Tensor, E: int):
    plans = getattr(eng, "train_plans", None) or {}
    key = _plan_key(tokens)
    if key in plans:
        return plans.pop(key)
    return plan_packing(eng, [tokens], E)[key]  # decoder called on its own: plan now (one host read here)


def _decoder_param_keys(model):
    return [k for k in model.engine.named_params() if k.startswith("decoder.")]


class DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, keys, tokens, category, decoding_type, enc_output, *params):
        eng: Engine = model.engine
        eng.sync_weights()
        opt, P, D, H = eng.opt, eng.P, eng.D, eng.H
        dev = eng.device
        raw_tokens = tokens
        tokens = tokens.contiguous()
        N, S = tokens.shape
        R = N * S
        Bv, E, _ = enc_output.shape
        if N != Bv:
            raise NotImplementedError("training decoder expects one token row per video (N=%d, videos=%d)" % (N, Bv))
        tok_flat = tokens.reshape(R)
        cat = category.contiguous() if category is not None else None
        seeds = Seeds()
        p = float(opt["hidden_dropout_prob"])
        if float(opt.get("attention_probs_dropout_prob", 0.0) or 0.0) > 0.0:
            raise NotImplementedError("attention_probs_dropout_prob > 0 in training (models/bert.py:135,169; the method "
                                      "presets use 0.0): the attention cores take no dropout seed")
        enc = eng.from_f32(enc_output.detach().reshape(Bv * E, D))
        extra = None
        if decoding_type == "NARFormer":
            ei = opt.get("enhance_input", 2)
            if ei == 2:
                extra = torch.empty((Bv, D), dtype=torch.float32, device=dev)
                L.call("navc_length_head", L.ptr(enc.f32), Bv, E, D, None, None, None, None, 0, L.ptr(extra), None, L.stream())
            elif ei != 0:
                raise NotImplementedError("enhance_input=1 fails in the reference itself (SURVEY 8c)")
        tc_attn = eng.tc_attention_ok(S, E)
        pk = _take_plan(eng, raw_tokens, E)
        kv = eng.linear(enc, P["kv_all"], f32=True, bf=tc_attn)
        emb = P["emb"]
        if pk is not None:  # rows = real positions only; every row is non-PAD, so `* non_pad_mask` is the identity
            R = pk["rows"]
            tok_flat = None
            x_ln = torch.empty((R, D), dtype=torch.float32, device=dev)
            L.call("navc_embed_ln_packed", L.ptr(tokens), L.ptr(cat), L.ptr(emb["word"]), L.ptr(emb["pos"]), L.ptr(emb["cat"]),
                   L.ptr(extra), 1, L.ptr(emb["ln_w"]), L.ptr(emb["ln_b"]), eng.eps, N, S, D, L.ptr(pk["seq_off"]),
                   L.ptr(pk["rowmap"]), None, L.ptr(x_ln), None, None, L.stream())
        else:
            x_ln = torch.empty((R, D), dtype=torch.float32, device=dev)
            L.call("navc_embed_ln", L.ptr(tokens), L.ptr(cat), L.ptr(emb["word"]), L.ptr(emb["pos"]), L.ptr(emb["cat"]),
                   L.ptr(extra), 1, L.ptr(emb["ln_w"]), L.ptr(emb["ln_b"]), eng.eps, N, S, D, L.ptr(x_ln), None, None, L.stream())
        seed_e = seeds.next()
        x = drop_add(eng, x_ln, None, seed_e, p, 0, 0.0, None)
        mask_kind = L.MASK_KIND[decoding_type]
        watch = int(opt.get("watch", 0))
        layers = []
        for l, lw in enumerate(P["layers"]):
            sv = dict(x=x)
            qkv = eng.linear(x, lw["qkv"], f32=True, bf=tc_attn)
            ctx1 = eng._new(R, D, True, True)
            if pk is not None:
                L.call("navc_self_attention_tc_rows", eng.tc_mode, L.ptr(qkv.hi), L.ptr(qkv.lo), 3 * D, L.ptr(tokens),
                       L.ptr(pk["seq_off"]), R, N, S, D, H, mask_kind, watch, L.ptr(ctx1.f32), L.ptr(ctx1.hi), L.ptr(ctx1.lo),
                       L.stream())
            elif tc_attn:
                L.call("navc_self_attention_tc", eng.tc_mode, L.ptr(qkv.hi), L.ptr(qkv.lo), 3 * D, L.ptr(tokens), N, S,
                       D, H, mask_kind, watch, L.ptr(ctx1.f32), L.ptr(ctx1.hi), L.ptr(ctx1.lo), L.stream())
            else:
                L.call("navc_self_attention", L.ptr(qkv.f32), 3 * D, L.ptr(tokens), N, S, D, H, mask_kind, watch,
                       L.ptr(ctx1.f32), L.ptr(ctx1.hi), L.ptr(ctx1.lo), None, L.stream())
            so = eng.linear(ctx1, lw["so"], f32=True, bf=False)
            sv["s_so"] = seeds.next()
            sv["so"] = {}
            a = _post(eng, so.f32, x.f32, lw["so_ln"], sv["s_so"], p, 0, 0.0, tok_flat, sv["so"])
            q = eng.linear(a, lw["cq"], f32=True, bf=tc_attn)
            ctx2 = eng._new(R, D, True, True)
            if pk is not None:
                off = l * 2 * D
                L.call("navc_cross_attention_tc_rows", eng.tc_mode, L.ptr(q.hi), L.ptr(q.lo), D, kv.hi[:, off:].data_ptr(),
                       kv.lo[:, off:].data_ptr() if kv.lo is not None else None, kv.N, L.ptr(pk["seq_off"]), R, N, S, E, D, H, 1,
                       L.ptr(ctx2.f32), L.ptr(ctx2.hi), L.ptr(ctx2.lo), L.stream())
            elif tc_attn:
                off = l * 2 * D
                L.call("navc_cross_attention_tc", eng.tc_mode, L.ptr(q.hi), L.ptr(q.lo), D, kv.hi[:, off:].data_ptr(),
                       kv.lo[:, off:].data_ptr() if kv.lo is not None else None, kv.N, N, S, E, D, H, 1,
                       L.ptr(ctx2.f32), L.ptr(ctx2.hi), L.ptr(ctx2.lo), L.stream())
            else:
                L.call("navc_cross_attention", L.ptr(q.f32), D, kv.f32[:, l * 2 * D:].data_ptr(), kv.N, N, S, E, D, H, 1,
                       L.ptr(ctx2.f32), L.ptr(ctx2.hi), L.ptr(ctx2.lo), None, L.stream())
            co = eng.linear(ctx2, lw["co"], f32=True, bf=False)
            sv["s_co"] = seeds.next()
            sv["co"] = {}
            c = _post(eng, co.f32, a.f32, lw["co_ln"], sv["s_co"], p, 0, 0.0, tok_flat, sv["co"])
            u = eng.linear(c, lw["f1"], f32=True, bf=False)
            h = eng._new(R, lw["f1"].N, True, True)
            L.call("navc_act_drop", L.ptr(u.f32), eng.act, 0, 0.0, u.f32.numel(), L.ptr(h.f32), L.ptr(h.hi), L.ptr(h.lo), L.stream())
            f2 = eng.linear(h, lw["f2"], f32=True, bf=False)
            sv["s_f1"], sv["s_f2"] = seeds.next(), seeds.next()
            sv["f2"] = {}
            xn = _post(eng, f2.f32, c.f32, lw["f2_ln"], sv["s_f1"], p, sv["s_f2"], p, tok_flat, sv["f2"])
            sv.update(qkv=qkv.f32, ctx1=ctx1, a=a, q=q.f32, ctx2=ctx2, c=c, u=u.f32, h=h)
            layers.append(sv)
            x = xn
        ctx.model, ctx.keys = model, keys
        ctx.state = dict(pack_id=eng.pack_id, tokens=tokens, tok_flat=tok_flat, cat=cat, enc=enc, extra=extra, kv=kv.f32, layers=layers,
                         seed_e=seed_e, p=p, N=N, S=S, E=E, Bv=Bv, mask_kind=mask_kind, watch=watch, decoding_type=decoding_type,
                         pk=pk)
        ctx.pk = pk
        eng.last_train_rows = (R, N * S)  # rows that went through the decoder GEMMs / padded rows (stats only)
        if pk is None:
            return x.f32.view(N, S, D)
        hidden = torch.zeros((N, S, D), dtype=torch.float32, device=dev)  # PAD rows are exactly zero, as in the reference
        L.call("navc_rows_f32", L.ptr(x.f32), L.ptr(hidden), D, L.ptr(pk["rowmap"]), R, 1, L.stream())
        return hidden

    @staticmethod
    def backward(ctx, d_hidden):
        model, st = ctx.model, ctx.state
        eng: Engine = model.engine
        _check_pack(eng, st)
        P, D, H = eng.P, eng.D, eng.H
        dev = eng.device
        N, S, E, Bv = st["N"], st["S"], st["E"], st["Bv"]
        R = N * S
        p = st["p"]
        tok_flat = st["tok_flat"]
        pk = st["pk"]
        grads = Grads(eng)
        g = d_hidden.contiguous().view(R, D).float()
        if pk is not None:  # gradient rows of the real positions (PAD rows get none: `* non_pad_mask`)
            R = pk["rows"]
            gp = torch.empty((R, D), dtype=torch.float32, device=dev)
            L.call("navc_rows_f32", L.ptr(g), L.ptr(gp), D, L.ptr(pk["rowmap"]), R, 0, L.stream())
            g = gp
        nl = len(P["layers"])
        d_kv = None   # fp32 gradient of the all-layer K|V projection output (allocated below unless it leaves split)
        # attention gradients on the tensor cores (csrc/attention_bwd_tc.cu: dk == 64, S <= 32, E <= 128, packed rows).
        # opt['navc_attn_bwd_tc'] / $NAVC_ATTN_BWD_TC: 1 (default) text -> video attention only -- 535 vs 669 us per launch
        # at 1024 videos; the token self-attention (<= 30 keys: one 32-key chunk of a 128-wide tile) measured 302 vs 288 us
        # against the fp32 CUDA-core kernel and stays there; 2 = both on the tensor cores; 0 = neither
        want = str(eng.opt.get("navc_attn_bwd_tc", os.environ.get("NAVC_ATTN_BWD_TC", "1"))).lower()
        tc_ok = pk is not None and eng.tc and eng.tc_attention_ok(S, E)
        tc_bwd = tc_ok and want not in ("0", "false", "no", "off")
        tc_bwd_self = tc_ok and want in ("2", "all", "both")
        # with the tensor-core kernel the K|V gradient leaves directly as the bf16 hi / lo operand of the projection's
        # gradient GEMMs, its column sums (bias gradient) accumulated by the same kernel: no fp32 d_kv (3 GB at 1024 videos)
        kv_lin = P["kv_all"]
        kv_split = None
        if tc_bwd and kv_lin.b is not None and kv_lin.K % 8 == 0 and os.environ.get("NAVC_KV_SPLIT", "1") not in ("0", "no", "off"):
            kv_db = grads.target_group([s_[1] for s_ in kv_lin.src])
            kv_db_direct = kv_db is not None and kv_db.numel() == kv_lin.N
            if not kv_db_direct:
                kv_db = torch.zeros((kv_lin.N,), dtype=torch.float32, device=dev)
            kv_split = Operand(Bv * E, kv_lin.N, kv_lin.N,
                               hi=torch.empty((Bv * E, kv_lin.N), dtype=torch.bfloat16, device=dev),
                               lo=torch.empty((Bv * E, kv_lin.N), dtype=torch.bfloat16, device=dev) if eng.split else None)
        else:
            d_kv = torch.empty((Bv * E, nl * 2 * D), dtype=torch.float32, device=dev)
        for l in range(nl - 1, -1, -1):
            lw, sv = P["layers"][l], st["layers"][l]
            # (the dY of every GEMM below leaves its producer directly as the bf16 hi / lo operand + bias column sums where
            # the layer has no LayerNorm in between: no fp32 dY, no separate split pass)
            sp = _dy_target(eng, lw["f2"], grads, g.shape[0]) if lw["f2_ln"] is None else None
            if sp is not None:
                d_c_res = _post_bwd_split(eng, g, sv["s_f1"], p, sv["s_f2"], p, tok_flat, sp)
                d_h = lin_bwd(eng, sv["h"], lw["f2"], None, grads, dy_split=sp[0], db_pre=(sp[1], sp[2]))
            else:
                d_f2, d_c_res = _post_bwd(eng, g, lw["f2_ln"], lw["f2_ln_key"], sv["s_f1"], p, sv["s_f2"], p, tok_flat, sv["f2"], grads)
                d_h = lin_bwd(eng, sv["h"], lw["f2"], d_f2, grads)
            sp = _dy_target(eng, lw["f1"], grads, d_h.shape[0])
            if sp is not None:
                L.call("navc_act_drop_bwd_split", L.ptr(d_h), L.ptr(sv["u"]), eng.act, 0, 0.0, d_h.shape[0], d_h.shape[1],
                       L.ptr(sp[0].hi), L.ptr(sp[0].lo), sp[0].ld, L.ptr(sp[1]), L.stream())
                d_c = lin_bwd(eng, sv["c"], lw["f1"], None, grads, dx_residual=d_c_res, dy_split=sp[0], db_pre=(sp[1], sp[2]))
            else:
                d_u = torch.empty_like(d_h)
                L.call("navc_act_drop_bwd", L.ptr(d_h), L.ptr(sv["u"]), eng.act, 0, 0.0, d_h.numel(), L.ptr(d_u), L.stream())
                d_c = lin_bwd(eng, sv["c"], lw["f1"], d_u, grads, dx_residual=d_c_res)
            sp = _dy_target(eng, lw["co"], grads, d_c.shape[0]) if lw["co_ln"] is None else None
            if sp is not None:
                d_a_res = _post_bwd_split(eng, d_c, sv["s_co"], p, 0, 0.0, tok_flat, sp)
                d_ctx2 = lin_bwd(eng, sv["ctx2"], lw["co"], None, grads, dy_split=sp[0], db_pre=(sp[1], sp[2]))
            else:
                d_co, d_a_res = _post_bwd(eng, d_c, lw["co_ln"], lw["co_ln_key"], sv["s_co"], p, 0, 0.0, tok_flat, sv["co"], grads)
                d_ctx2 = lin_bwd(eng, sv["ctx2"], lw["co"], d_co, grads)
            d_q = torch.empty((R, D), dtype=torch.float32, device=dev)
            off = l * 2 * D
            if pk is not None and kv_split is not None:
                L.call("navc_cross_attention_bwd_tc_split", eng.tc_mode, L.ptr(sv["q"]), D, st["kv"][:, off:].data_ptr(),
                       st["kv"].shape[1], L.ptr(pk["seq_off"]), N, S, E, D, H, L.ptr(d_ctx2), L.ptr(sv["ctx2"].f32), L.ptr(d_q), D,
                       kv_split.hi[:, off:].data_ptr(), kv_split.lo[:, off:].data_ptr() if kv_split.lo is not None else None,
                       kv_lin.N, kv_db[off:].data_ptr(), L.stream())
            elif pk is not None and tc_bwd:
                L.call("navc_cross_attention_bwd_tc", eng.tc_mode, L.ptr(sv["q"]), D, st["kv"][:, off:].data_ptr(), st["kv"].shape[1],
                       L.ptr(pk["seq_off"]), N, S, E, D, H, L.ptr(d_ctx2), L.ptr(sv["ctx2"].f32), L.ptr(d_q), D,
                       d_kv[:, off:].data_ptr(), d_kv.shape[1], L.stream())
            elif pk is not None:
                L.call("navc_cross_attention_bwd_packed", L.ptr(sv["q"]), D, st["kv"][:, off:].data_ptr(), st["kv"].shape[1],
                       L.ptr(pk["seq_off"]), N, S, E, D, H, L.ptr(d_ctx2), L.ptr(d_q), D, d_kv[:, off:].data_ptr(),
                       d_kv.shape[1], L.stream())
            else:
                L.call("navc_cross_attention_bwd", L.ptr(sv["q"]), D, st["kv"][:, off:].data_ptr(), st["kv"].shape[1], N, S, E, D, H, 1,
                       L.ptr(d_ctx2), L.ptr(d_q), D, d_kv[:, off:].data_ptr(), d_kv.shape[1], L.stream())
            d_a = lin_bwd(eng, sv["a"], lw["cq"], d_q, grads, dx_residual=d_a_res)
            sp = _dy_target(eng, lw["so"], grads, d_a.shape[0]) if lw["so_ln"] is None else None
            if sp is not None:
                d_x_res = _post_bwd_split(eng, d_a, sv["s_so"], p, 0, 0.0, tok_flat, sp)
                d_ctx1 = lin_bwd(eng, sv["ctx1"], lw["so"], None, grads, dy_split=sp[0], db_pre=(sp[1], sp[2]))
            else:
                d_so, d_x_res = _post_bwd(eng, d_a, lw["so_ln"], lw["so_ln_key"], sv["s_so"], p, 0, 0.0, tok_flat, sv["so"], grads)
                d_ctx1 = lin_bwd(eng, sv["ctx1"], lw["so"], d_so, grads)
            d_qkv = torch.empty((R, 3 * D), dtype=torch.float32, device=dev)
            if pk is not None and tc_bwd_self:
                L.call("navc_self_attention_bwd_tc", eng.tc_mode, L.ptr(sv["qkv"]), 3 * D, L.ptr(pk["seq_off"]), N, S, D, H,
                       st["mask_kind"], st["watch"], L.ptr(d_ctx1), L.ptr(sv["ctx1"].f32), L.ptr(d_qkv), L.stream())
            elif pk is not None:
                L.call("navc_self_attention_bwd_packed", L.ptr(sv["qkv"]), 3 * D, L.ptr(st["tokens"]), L.ptr(pk["seq_off"]), N, S,
                       D, H, st["mask_kind"], st["watch"], L.ptr(d_ctx1), L.ptr(d_qkv), L.stream())
            else:
                L.call("navc_self_attention_bwd", L.ptr(sv["qkv"]), 3 * D, L.ptr(st["tokens"]), N, S, D, H, st["mask_kind"],
                       st["watch"], L.ptr(d_ctx1), L.ptr(d_qkv), L.stream())
            g = lin_bwd(eng, sv["x"], lw["qkv"], d_qkv, grads, dx_residual=d_x_res)
        # embeddings
        g_ln, _ = drop_add_bwd(g, st["seed_e"], p, 0, 0.0, None, want_res=False)
        emb, ep = P["emb"], P["emb_prefix"]
        d_word = torch.zeros_like(emb["word"])
        d_pos = torch.zeros_like(emb["pos"])
        d_cat = torch.zeros_like(emb["cat"]) if emb["cat"] is not None else None
        d_extra = torch.zeros_like(st["extra"]) if st["extra"] is not None else None
        d_lw, d_lb = torch.zeros_like(emb["ln_w"]), torch.zeros_like(emb["ln_b"])
        if pk is not None:
            L.call("navc_embed_ln_bwd_packed", L.ptr(g_ln), L.ptr(st["tokens"]), L.ptr(st["cat"]), L.ptr(emb["word"]),
                   L.ptr(emb["pos"]), L.ptr(emb["cat"]), L.ptr(st["extra"]), 1, L.ptr(emb["ln_w"]), eng.eps, S, D,
                   L.ptr(pk["rowmap"]), R, L.ptr(d_word), L.ptr(d_pos), L.ptr(d_cat), L.ptr(d_extra), L.ptr(d_lw), L.ptr(d_lb),
                   L.stream())
        else:
            L.call("navc_embed_ln_bwd", L.ptr(g_ln), L.ptr(st["tokens"]), L.ptr(st["cat"]), L.ptr(emb["word"]), L.ptr(emb["pos"]),
                   L.ptr(emb["cat"]), L.ptr(st["extra"]), 1, L.ptr(emb["ln_w"]), L.ptr(emb["ln_b"]), eng.eps, N, S, D,
                   L.ptr(d_word), L.ptr(d_pos), L.ptr(d_cat), L.ptr(d_extra), L.ptr(d_lw), L.ptr(d_lb), L.stream())
        grads.add(ep + "word_embeddings.weight", d_word)
        grads.add(ep + "position_embeddings.weight", d_pos)
        if d_cat is not None:
            grads.add(ep + "category_embeddings.weight", d_cat)
        grads.add(ep + "LayerNorm.weight", d_lw)
        grads.add(ep + "LayerNorm.bias", d_lb)
        # encoder memory: through the K|V projection of every layer and the enhance_input mean
        if kv_split is not None:
            d_enc = lin_bwd(eng, st["enc"], kv_lin, None, grads, dy_split=kv_split, db_pre=(kv_db, kv_db_direct))
        else:
            d_enc = lin_bwd(eng, st["enc"], kv_lin, d_kv, grads)
        if d_extra is not None:
            L.call("navc_mean_bwd", L.ptr(d_extra), Bv, E, D, L.ptr(d_enc), L.stream())
        return (None, None, None, None, None, d_enc.view(Bv, E, D)) + tuple(grads.g.get(k) for k in ctx.keys)


def decoder_forward_train(decoder, eng, tgt_seq, enc_output, category, decoding_type):
    model = decoder._engine_ref()
    keys = _decoder_param_keys(model)
    hidden = DecoderFn.apply(model, keys, tgt_seq, category, decoding_type, enc_output, *_param_list(model, keys))
    pk = getattr(hidden.grad_fn, "pk", None)
    if pk is not None:  # the vocabulary projection of exactly this tensor may skip the PAD rows too (vocab_logprobs_train)
        hidden._navc_pack = pk
    with torch.no_grad():  # returned, unused downstream (models/bert.py:301)
        non_pad = tgt_seq.ne(Constants.PAD).float().unsqueeze(-1)
        embs = hidden.detach().sum(1) / non_pad.sum(1)
    return ([hidden], embs)


# --------------------------------------------------------------------------------------------------
# vocabulary projection (+ log-softmax)  (models/seq2seq.py:102-103)
# --------------------------------------------------------------------------------------------------
class VocabFn(torch.autograd.Function):
    """hidden [.., D] -> log_softmax(tgt_word_prj(hidden)) [.., V] (log_probs=True) or the logits."""

    @staticmethod
    def forward(ctx, model, log_probs, pk, hidden, *params):
        eng: Engine = model.engine
        eng.sync_weights()
        lin = eng.P["vocab"]
        dev = eng.device
        shape = hidden.shape
        D, V = lin.K, lin.N
        h2 = hidden.detach().reshape(-1, D)
        R = R_all = h2.shape[0]
        # packed rows (pk): `hidden` is the untouched tensor DecoderFn produced (PAD rows exactly zero) -> only the real
        # positions go through the projection; the PAD rows of the output are log_softmax(bias), a constant row
        if pk is not None:
            R = pk["rows"]
            hp = torch.empty((R, D), dtype=torch.float32, device=dev)
            L.call("navc_rows_f32", L.ptr(h2), L.ptr(hp), D, L.ptr(pk["rowmap"]), R, 0, L.stream())
            h2 = hp
        h = eng.from_f32(h2)
        Vp = _up(V, 64)
        logits = torch.empty((R, Vp), dtype=torch.float32, device=dev)
        if lin.pad is None:  # weight rows / bias zero-padded to Vp so the GEMM epilogue stays vectorised
            w_pad = torch.zeros((Vp, D), dtype=torch.float32, device=dev)
            w_pad[:V].copy_(lin.w)
            b_pad = None
            if lin.b is not None:
                b_pad = torch.zeros((Vp,), dtype=torch.float32, device=dev)
                b_pad[:V].copy_(lin.b)
            wop, _ = transpose_pack(eng, w_pad, Vp, D, D, straight=True)
            lin.pad = (wop, b_pad)
        wop, b_pad = lin.pad
        xop = Operand(R, D, D, hi=h.hi, lo=h.lo) if eng.tc else Operand(R, D, D, f32=h.f32)
        gemm(eng, xop, wop, R, Vp, D, logits, Vp, bias=b_pad)
        const_lp = None
        if pk is not None:
            if lin.b is not None:
                const_lp = torch.empty((1, V), dtype=torch.float32, device=dev)
                L.call("navc_log_softmax_ld", L.ptr(b_pad), Vp, L.ptr(const_lp), V, 1, V, L.stream())
            else:
                const_lp = torch.full((1, V), -math.log(V), dtype=torch.float32, device=dev)
            out = const_lp.expand(R_all, V).contiguous()
            L.call("navc_log_softmax_rows", L.ptr(logits), Vp, L.ptr(out), V, L.ptr(pk["rowmap"]), R, V, L.stream())
        else:
            out = torch.empty((R, V), dtype=torch.float32, device=dev)
            if log_probs:
                L.call("navc_log_softmax_ld", L.ptr(logits), Vp, L.ptr(out), V, R, V, L.stream())
            else:
                out.copy_(logits[:, :V])
        ctx.model, ctx.log_probs = model, log_probs
        ctx.state = dict(pack_id=eng.pack_id, h=h, out=out if log_probs else None, R=R, V=V, Vp=Vp, shape=shape, n_params=len(params), pk=pk,
                         R_all=R_all, const_lp=const_lp)
        return out.view(*shape[:-1], V)

    @staticmethod
    def backward(ctx, d_out):
        model, st = ctx.model, ctx.state
        eng: Engine = model.engine
        _check_pack(eng, st)
        lin = eng.P["vocab"]
        dev = eng.device
        R, V, Vp, pk = st["R"], st["V"], st["Vp"], st["pk"]
        grads = Grads(eng)
        d_out = d_out.contiguous().view(st["R_all"], V)
        dlog = torch.empty((R, Vp), dtype=torch.float32, device=dev)
        if pk is not None:
            L.call("navc_log_softmax_bwd_rows", L.ptr(d_out), L.ptr(st["out"]), L.ptr(pk["rowmap"]), R, V, V, L.ptr(dlog), Vp,
                   L.stream())
        elif ctx.log_probs:
            L.call("navc_log_softmax_bwd", L.ptr(d_out), L.ptr(st["out"]), R, V, V, L.ptr(dlog), Vp, L.stream())
        else:
            dlog.zero_()
            dlog[:, :V].copy_(d_out)
        d_h = lin_bwd(eng, st["h"], lin, dlog, grads, ld_dy=Vp)
        gw = grads.g.get("tgt_word_prj.weight")
        gb = grads.g.get("tgt_word_prj.bias")
        if pk is not None:
            if gb is not None and st["R_all"] > R:  # the skipped PAD rows only reach the bias (zero for PAD-ignoring losses)
                L.call("navc_log_softmax_bwd_padrows", L.ptr(d_out), V, L.ptr(pk["pad_rows"]), st["R_all"] - R,
                       L.ptr(st["const_lp"]), V, L.ptr(gb), L.stream())
            dp = torch.zeros((st["R_all"], d_h.shape[1]), dtype=torch.float32, device=dev)
            L.call("navc_rows_f32", L.ptr(d_h), L.ptr(dp), d_h.shape[1], L.ptr(pk["rowmap"]), R, 1, L.stream())
            d_h = dp
        return (None, None, None, d_h.view(st["shape"])) + ((gw,) if st["n_params"] == 1 else (gw, gb))


def _vocab_params(prj):
    return [prj.weight] + ([prj.bias] if prj.bias is not None else [])


def vocab_forward_train(engine, prj, hidden):
    """model.tgt_word_prj(hidden) under autograd -> logits."""
    return VocabFn.apply(prj._owner(), False, None, hidden, *_vocab_params(prj))


def vocab_logprobs_train(model, hidden):
    """log_softmax(tgt_word_prj(hidden)) as one autograd node (projection + log-softmax kernels)."""
    pk = getattr(hidden, "_navc_pack", None)
    if pk is not None and not (hidden._version == 0 and hidden.is_contiguous() and hidden.dim() == 3
                               and hidden.shape[0] * hidden.shape[1] == pk["N"] * pk["S"]):
        pk = None  # modified since the decoder produced it: PAD rows may no longer be zero
    return VocabFn.apply(model, True, pk, hidden, *_vocab_params(model.tgt_word_prj))


# --------------------------------------------------------------------------------------------------
# fused cross-entropy: projection + log-softmax + masked NLL without the [rows, V] log-prob tensor
# (reference seq2seq.py:102-103 + misc/crit.py:62-84)
# --------------------------------------------------------------------------------------------------
class LazyLogProbs:
    """Stands in for one ``tgt_word_logprobs`` entry when ``opt['navc_fused_ce']`` is set: the hidden
    states plus the model, consumed by ``navc_b200.misc.crit.LanguageGeneration`` (fused loss) --
    ``materialize()`` gives the reference's [B, S, V] log-prob tensor for any other consumer."""

    def __init__(self, model, hidden):
        self.model, self.hidden = model, hidden
        self.stats = None  # (nll [R], argmax [R]) of the last fused loss evaluation

    def size(self, dim=None):
        shape = tuple(self.hidden.shape[:-1]) + (self.model.engine.P["vocab"].N if self.model.engine.P else self.model.opt["vocab_size"],)
        return shape if dim is None else shape[dim]

    def materialize(self):
        return vocab_logprobs_train(self.model, self.hidden)

    def nll_sum(self, labels):
        """sum over non-PAD positions of -log p(label) as ONE autograd node."""
        out = FusedCEFn.apply(self, labels, self.hidden, *_vocab_params(self.model.tgt_word_prj))
        return out


class FusedCEFn(torch.autograd.Function):
    CHUNK = 2048  # rows per backward chunk: a [2048, V] fp32 logits slab (86 MB at V = 10547) stays L2 resident

    @staticmethod
    def forward(ctx, lazy, labels, hidden, *params):
        model = lazy.model
        eng: Engine = model.engine
        eng.sync_weights()
        lin = eng.P["vocab"]
        dev = eng.device
        D = lin.K
        h = eng.from_f32(hidden.detach().reshape(-1, D))
        R = h.M
        lab = labels.contiguous().view(-1)
        pm, ps, pi, nt, tl = eng.vocab_partials(h, target=lab)
        lse = torch.empty((R,), dtype=torch.float32, device=dev)
        nll = torch.empty((R,), dtype=torch.float32, device=dev)
        arg = torch.empty((R,), dtype=torch.int32, device=dev)
        L.call("navc_ce_stats", L.ptr(pm), L.ptr(ps), L.ptr(pi), nt, L.ptr(tl), L.ptr(lab), R, L.ptr(lse), L.ptr(nll), L.ptr(arg), L.stream())
        lazy.stats = (nll.view(labels.shape), arg.view(labels.shape))
        ctx.lazy, ctx.n_params = lazy, len(params)
        ctx.state = dict(pack_id=eng.pack_id, h=h, lse=lse, lab=lab, shape=hidden.shape)
        return nll.sum()

    @staticmethod
    def backward(ctx, g):
        model = ctx.lazy.model
        eng: Engine = model.engine
        st = ctx.state
        _check_pack(eng, st)
        lin = eng.P["vocab"]
        dev = eng.device
        h, lse, lab = st["h"], st["lse"], st["lab"]
        R, D, V = h.M, lin.K, lin.N
        Vp = _up(V, 64)
        if lin.pad is None:
            w_pad = torch.zeros((Vp, D), dtype=torch.float32, device=dev)
            w_pad[:V].copy_(lin.w)
            b_pad = None
            if lin.b is not None:
                b_pad = torch.zeros((Vp,), dtype=torch.float32, device=dev)
                b_pad[:V].copy_(lin.b)
            wop, _ = transpose_pack(eng, w_pad, Vp, D, D, straight=True)
            lin.pad = (wop, b_pad)
        wop, b_pad = lin.pad
        scale = g.detach().reshape(1).float().contiguous()
        grads = Grads(eng)
        d_h = torch.empty((R, D), dtype=torch.float32, device=dev)
        C = FusedCEFn.CHUNK
        slab = torch.empty((min(C, R), Vp), dtype=torch.float32, device=dev)
        for r0 in range(0, R, C):
            n = min(C, R - r0)
            xop = Operand(n, D, D, hi=h.hi[r0:r0 + n], lo=None if h.lo is None else h.lo[r0:r0 + n]) if eng.tc \
                else Operand(n, D, D, f32=h.f32[r0:r0 + n])
            gemm(eng, xop, wop, n, Vp, D, slab, Vp, bias=b_pad)                      # logits of this row chunk
            L.call("navc_ce_grad", L.ptr(slab), lse[r0:].data_ptr(), lab[r0:].data_ptr(), L.ptr(scale), n, V, Vp, L.stream())
            xa = Act(n, D, f32=h.f32[r0:r0 + n], hi=None if h.hi is None else h.hi[r0:r0 + n], lo=None if h.lo is None else h.lo[r0:r0 + n])
            d_h[r0:r0 + n] = lin_bwd(eng, xa, lin, slab[:n], grads, ld_dy=Vp)
        gw = grads.g.get("tgt_word_prj.weight")
        gb = grads.g.get("tgt_word_prj.bias")
        return (None, None, d_h.view(st["shape"])) + ((gw,) if ctx.n_params == 1 else (gw, gb))
