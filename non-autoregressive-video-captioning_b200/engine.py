"""Host-side orchestration of the sm_100a kernels behind the reference's model API.

``Engine`` owns the packed weights of one ``Seq2Seq`` model (concatenated QKV / cross K|V of all
layers, bf16 hi/lo copies for the tensor-core modes) and issues the C-ABI calls of
``include/navc.h`` on torch-owned device buffers.  PyTorch is used for memory, streams and module
plumbing only; every arithmetic step of the hot path is one of our kernels.

Precision modes (``opt['navc_precision']`` or ``$NAVC_PRECISION``):
  ``fp32``    CUDA-core fp32 GEMMs: reference-exact mode used for bit-exact token parity.
  ``bf16x3``  tcgen05 GEMMs on split-bf16 operands (3 MMAs / product): ~fp32 accuracy.
  ``tf32``    tcgen05 ``kind::tf32`` GEMMs (2 bf16-MMA time units / product, 10-bit mantissa: ~5e-4 relative on the
              logits) wherever the A operand exists in fp32 -- QKV / query / FFN / K|V / encoder projections; the
              out-projections (their input comes from the attention cores as a bf16 hi/lo pair), the attention cores and
              the vocabulary projection stay split-bf16.
  ``bf16``    tcgen05 GEMMs on bf16 operands: fastest, ~3e-3 relative logit error.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from . import _lib as L
from .config import Constants

PRECISIONS = ("fp32", "bf16x3", "bf16", "tf32")


def default_precision(opt=None) -> str:
    p = (opt or {}).get("navc_precision") or os.environ.get("NAVC_PRECISION") or "bf16x3"
    if p not in PRECISIONS:
        raise ValueError("navc precision must be one of %s, got %r" % (PRECISIONS, p))
    return p


def _f32(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous and 16-byte aligned view / copy of a parameter (kernels use 128-bit loads)."""
    t = t.detach().contiguous().float()
    return t if t.data_ptr() % 16 == 0 else t.clone()


class Act:
    """An activation matrix [M, N]: fp32 copy and/or bf16 hi(/lo) copies."""
    __slots__ = ("f32", "hi", "lo", "M", "N")

    def __init__(self, M, N, f32=None, hi=None, lo=None):
        self.M, self.N, self.f32, self.hi, self.lo = M, N, f32, hi, lo


class PackedLinear:
    """One (possibly concatenated) nn.Linear in kernel operand formats.  ``src`` lists the
    state_dict entries the rows came from: (weight key, bias key or None, row0, row1) -- the training
    path uses it to hand gradients of the packed matrix back to the individual parameters."""
    __slots__ = ("w", "b", "w_hi", "w_lo", "N", "K", "src", "T", "pad")

    def __init__(self, w, b, src=()):
        self.w = _f32(w)
        self.b = None if b is None else _f32(b)
        self.N, self.K = self.w.shape
        self.w_hi = self.w_lo = None
        self.src = tuple(src)
        self.T = None  # transposed operand copy for dgrad GEMMs (training.py), built on first use
        self.pad = None  # row-padded operand copy (vocabulary projection in training), built on first use


class Engine:
    def __init__(self, model, precision: Optional[str] = None):
        self.model = model
        self.opt = model.opt
        self.precision = precision or default_precision(self.opt)
        self.tc = self.precision != "fp32"
        self.tc_mode = {"bf16": L.TC_BF16, "bf16x3": L.TC_BF16X3, "tf32": L.TC_BF16X3}.get(self.precision, 0)
        self.split = self.precision in ("bf16x3", "tf32")   # bf16 operands carry a lo part
        self.tf32 = self.precision == "tf32"
        self._sig = None
        self._members = None
        self._named = None
        self.grad_sink = None   # set by parallel.GradientAllReduce: parameter -> its p.grad view, for in-place accumulation
        self.device = None
        self.P: Dict[str, object] = {}
        # optional live timing of the launch classes (bench.py roofline): when `prof` is a dict, CUDA events are
        # recorded on the launching stream around every tagged launch: prof[tag] = [(start, stop), ...]
        self.prof = None
        self.pack_id = 0
        self._lins: List[PackedLinear] = []
        self._copies = []
        self._refresh_table = None
        self._ptr_sig = None
        self.graphs: Dict[tuple, object] = {}  # CUDA graphs of the decode loop (decoding/na_generate.py)

    # ------------------------------------------------------------------------------------------
    # weight packing
    # ------------------------------------------------------------------------------------------
    def members(self):
        """(named parameters, tracked buffers) of the model, walked once: nn.Module.named_parameters() visits
        every sub-module and was ~30 % of the host time of a training step.  Parameter OBJECTS are stable under
        load_state_dict / .to() / optimizer steps; after replacing one (``m.weight = nn.Parameter(..)``) call
        ``invalidate(structure=True)``."""
        if self._members is None:
            # remove_duplicate=False: with opt['tie_weights'] the vocabulary projection shares the word-embedding
            # Parameter (seq2seq.py:30-33) and both state_dict names must resolve
            self._members = (list(self.model.named_parameters(remove_duplicate=False)),
                             [(n, b) for n, b in self.model.named_buffers() if not n.endswith("num_batches_tracked")])
        return self._members

    def named_params(self) -> Dict[str, torch.Tensor]:
        if self._named is None:
            self._named = dict(self.members()[0])
        return self._named

    def _signature(self):
        # num_batches_tracked is bookkeeping only (BatchNorm momentum is fixed); the running statistics
        # are referenced in place, so train-mode updates need no repack
        params, bufs = self.members()
        return tuple((p.data_ptr(), p._version) for _, p in params) + tuple((b.data_ptr(), b._version) for _, b in bufs)

    def sync_weights(self):
        """(Re)pack weights if any parameter changed (optimizer step, load_state_dict, .to())."""
        dev = self.members()[0][0][1].device
        L.ensure_init(dev)
        if dev.type == "cuda" and dev.index is not None and torch.cuda.current_device() != dev.index:
            # every launch goes to the CURRENT device's current stream: a model on cuda:1 driven while cuda:0 is
            # current would launch on device 0 against device-1 pointers
            torch.cuda.set_device(dev)
        sig = self._signature()
        if sig == self._sig and dev == self.device:
            return
        ptrs = tuple(a[0] for a in sig)
        same_storage = dev == self.device and self._refresh_table is not None and ptrs == self._ptr_sig
        self.device = dev
        if same_storage and os.environ.get("NAVC_REFRESH", "1") not in ("0", "no", "off"):
            self._refresh()   # an optimizer step: same tensors, new values -- one launch instead of a full repack
        else:
            self._pack()
        self._sig = sig
        self._ptr_sig = ptrs
        self.pack_id += 1
        self.graphs.clear()  # captured graphs hold pointers into the previous packed weights

    def invalidate(self, structure: bool = False):
        """Force a repack at the next use (for writers that bypass torch's version counters);
        structure=True also drops the cached parameter list (a parameter / sub-module OBJECT was replaced)."""
        self._sig = None
        if structure:
            self._members = None
            self._named = None
            self._ptr_sig = None

    def _lin(self, w, b=None, src=()) -> PackedLinear:
        pl = PackedLinear(w, b, src)
        if self.tc:
            pl.w_hi = torch.empty(pl.w.shape, dtype=torch.bfloat16, device=pl.w.device)
            pl.w_lo = torch.empty_like(pl.w_hi) if self.split else None
            L.call("navc_split_bf16", L.ptr(pl.w), L.ptr(pl.w_hi), L.ptr(pl.w_lo), pl.w.numel(), L.stream())
        self._lins.append(pl)
        return pl

    def _live(self, t: torch.Tensor) -> torch.Tensor:
        """_f32 of a parameter / buffer that the kernels read in place; a copy (dtype / alignment) is registered for refresh."""
        v = _f32(t)
        if v.data_ptr() != t.data_ptr():
            self._copies.append((t, v))
        return v

    def _build_refresh_table(self):
        """Device table for navc_refresh_pack: every source parameter -> its rows of the packed fp32 operand (if that is a
        copy) and of the bf16 hi / lo copies."""
        named = dict(self.members()[0])
        rows = []
        for pl in self._lins:
            for wk, bk, r0, r1 in pl.src:
                w = named.get(wk)
                if w is None or w.dtype != torch.float32 or not w.is_contiguous():
                    return None
                dst = pl.w[r0:r1]
                rows.append((w.data_ptr(), 0 if dst.data_ptr() == w.data_ptr() else dst.data_ptr(),
                             pl.w_hi[r0:r1].data_ptr() if pl.w_hi is not None else 0,
                             pl.w_lo[r0:r1].data_ptr() if pl.w_lo is not None else 0, w.numel()))
                if bk is not None and pl.b is not None:
                    b = named.get(bk)
                    if b is None or b.dtype != torch.float32 or not b.is_contiguous():
                        return None
                    dstb = pl.b[r0:r1]
                    if dstb.data_ptr() != b.data_ptr():
                        rows.append((b.data_ptr(), dstb.data_ptr(), 0, 0, b.numel()))
        for t, v in self._copies:
            if t.dtype != torch.float32 or not t.is_contiguous():
                return None
            rows.append((t.data_ptr(), v.data_ptr(), 0, 0, t.numel()))
        rows = [r for r in rows if r[1] or r[2]]
        if not rows:
            return None
        return torch.tensor(rows, dtype=torch.int64).to(self.device)

    def _refresh(self):
        L.call("navc_refresh_pack", L.ptr(self._refresh_table), int(self._refresh_table.shape[0]), L.stream())
        for pl in self._lins:
            pl.T = None
            pl.pad = None

    def _pack(self):
        m, opt = self.model, self.opt
        self._lins, self._copies, self._refresh_table = [], [], None
        params, bufs = self.members()  # what state_dict() holds (minus num_batches_tracked), without the module walk
        sd = {k: v.detach() for k, v in params}
        sd.update(bufs)
        P = {}
        with torch.no_grad():
            # encoder streams (models/Encoder.py)
            P["streams"] = []
            for ch in opt["modality"].lower():
                pre = "encoder.Encoder_%s" % ch.upper()
                gate = (pre + ".1.w2.weight") in sd
                w12 = torch.cat([sd[pre + ".1.w1.weight"]] + ([sd[pre + ".1.w2.weight"]] if gate else []), 0)
                b12 = torch.cat([sd[pre + ".1.w1.bias"]] + ([sd[pre + ".1.w2.bias"]] if gate else []), 0)
                D_ = sd[pre + ".0.weight"].shape[0]
                src12 = [(pre + ".1.w1.weight", pre + ".1.w1.bias", 0, D_)] + \
                    ([(pre + ".1.w2.weight", pre + ".1.w2.bias", D_, 2 * D_)] if gate else [])
                P["streams"].append(dict(l0=self._lin(sd[pre + ".0.weight"], sd[pre + ".0.bias"],
                                                      [(pre + ".0.weight", pre + ".0.bias", 0, D_)]),
                                         l12=self._lin(w12, b12, src12), gate=int(gate)))
            # joint representation norms
            P["norms"] = []
            jr = "joint_representation_learner."
            for i in range(len(opt["modality"])):
                if (jr + "bn%d.weight" % i) in sd:
                    P["norms"].append(("bn", self._live(sd[jr + "bn%d.running_mean" % i]), self._live(sd[jr + "bn%d.running_var" % i]),
                                       self._live(sd[jr + "bn%d.weight" % i]), self._live(sd[jr + "bn%d.bias" % i])))
                elif (jr + "ln%d.weight" % i) in sd:
                    P["norms"].append(("ln", self._live(sd[jr + "ln%d.weight" % i]), self._live(sd[jr + "ln%d.bias" % i])))
                else:
                    P["norms"].append(None)
            # length head
            ap = "auxiliary_task_predictor.layers.0.net."
            P["len_head"] = None
            P["norm_keys"] = [jr + ("bn%d" if (jr + "bn%d.weight" % i) in sd else "ln%d") % i for i in range(len(opt["modality"]))]
            if (ap + "0.weight") in sd:
                P["len_head"] = tuple(self._live(sd[ap + k]) for k in ("0.weight", "0.bias", "3.weight", "3.bias"))
                P["len_head_keys"] = tuple(ap + k for k in ("0.weight", "0.bias", "3.weight", "3.bias"))
            # decoder
            dp = "decoder.bert." if ("decoder.bert.embedding.LayerNorm.weight" in sd) else "decoder."
            e = dp + "embedding."
            P["emb_prefix"] = e
            P["emb"] = dict(word=self._live(sd[e + "word_embeddings.weight"]),
                            pos=self._live(sd[e + "position_embeddings.weight"]),
                            cat=self._live(sd[e + "category_embeddings.weight"]) if (e + "category_embeddings.weight") in sd else None,
                            ln_w=self._live(sd[e + "LayerNorm.weight"]), ln_b=self._live(sd[e + "LayerNorm.bias"]))
            layers, kv_w, kv_b, kv_src = [], [], [], []
            D_ = opt["dim_hidden"]
            for l in range(opt["num_hidden_layers_decoder"]):
                lp = "%slayer.%d." % (dp, l)
                sa, ca = lp + "attention.", lp + "attend_to_enc_output."
                qkv_w = torch.cat([sd[sa + "self.%s.weight" % n] for n in ("query", "key", "value")], 0)
                qkv_b = torch.cat([sd[sa + "self.%s.bias" % n] for n in ("query", "key", "value")], 0)
                qkv_src = [(sa + "self.%s.weight" % n, sa + "self.%s.bias" % n, i * D_, (i + 1) * D_)
                           for i, n in enumerate(("query", "key", "value"))]
                kv_w.append(torch.cat([sd[ca + "self.key.weight"], sd[ca + "self.value.weight"]], 0))
                kv_b.append(torch.cat([sd[ca + "self.key.bias"], sd[ca + "self.value.bias"]], 0))
                kv_src += [(ca + "self.%s.weight" % n, ca + "self.%s.bias" % n, (2 * l + i) * D_, (2 * l + i + 1) * D_)
                           for i, n in enumerate(("key", "value"))]

                def ln(prefix):
                    k = prefix + "LayerNorm.weight"
                    return (self._live(sd[k]), self._live(sd[prefix + "LayerNorm.bias"])) if k in sd else None

                def one(prefix):
                    w = sd[prefix + ".weight"]
                    return self._lin(w, sd[prefix + ".bias"], [(prefix + ".weight", prefix + ".bias", 0, w.shape[0])])
                layers.append(dict(
                    qkv=self._lin(qkv_w, qkv_b, qkv_src), so=one(sa + "output.dense"), so_ln=ln(sa + "output."),
                    so_ln_key=sa + "output.LayerNorm",
                    cq=one(ca + "self.query"), co=one(ca + "output.dense"), co_ln=ln(ca + "output."),
                    co_ln_key=ca + "output.LayerNorm",
                    f1=one(lp + "intermediate.dense"), f2=one(lp + "output.dense"), f2_ln=ln(lp + "output."),
                    f2_ln_key=lp + "output.LayerNorm"))
            P["layers"] = layers
            P["kv_all"] = self._lin(torch.cat(kv_w, 0), torch.cat(kv_b, 0), kv_src)  # [L*2D, D]
            vb = "tgt_word_prj.bias" if "tgt_word_prj.bias" in sd else None
            P["vocab"] = self._lin(sd["tgt_word_prj.weight"], sd.get("tgt_word_prj.bias"),
                                   [("tgt_word_prj.weight", vb, 0, sd["tgt_word_prj.weight"].shape[0])])
        self.P = P
        self._refresh_table = self._build_refresh_table() if self.device is not None and self.device.type == "cuda" else None
        self.D = opt["dim_hidden"]
        self.H = opt["num_attention_heads"]
        self.nl = opt["num_hidden_layers_decoder"]
        self.act = L.ACT[opt["hidden_act"]]
        self.eps = float(opt["layer_norm_eps"])

    # ------------------------------------------------------------------------------------------
    # primitive wrappers
    # ------------------------------------------------------------------------------------------
    def _new(self, M, N, f32, bf, lo=False):
        """lo=True forces the bf16 lo copy also in plain bf16 mode (residual stream: hi + lo carries the
        fp32 value to ~2^-17 relative, so no fp32 copy of the stream has to be written or re-read)."""
        dev = self.device
        a = Act(M, N)
        if f32:
            a.f32 = torch.empty((M, N), dtype=torch.float32, device=dev)
        if bf and self.tc:
            a.hi = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
            if self.split or lo:
                a.lo = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
        return a

    def join_f32(self, a: Act) -> torch.Tensor:
        """fp32 view of an activation kept as a bf16 hi/lo pair (API boundary only)."""
        if a.f32 is None:
            a.f32 = torch.empty((a.M, a.N), dtype=torch.float32, device=self.device)
            L.call("navc_join_bf16", L.ptr(a.hi), L.ptr(a.lo), L.ptr(a.f32), a.M * a.N, L.stream())
        return a.f32

    def from_f32(self, x2d: torch.Tensor, need_bf=True) -> Act:
        x2d = x2d.contiguous().float()
        a = Act(x2d.shape[0], x2d.shape[1], f32=x2d)
        if self.tc and need_bf:
            a.hi = torch.empty(x2d.shape, dtype=torch.bfloat16, device=x2d.device)
            a.lo = torch.empty_like(a.hi) if self.split else None
            L.call("navc_split_bf16", L.ptr(x2d), L.ptr(a.hi), L.ptr(a.lo), x2d.numel(), L.stream())
        return a

    def linear(self, x: Act, lin: PackedLinear, act=0, residual=None,
               row_tokens: Optional[torch.Tensor] = None, f32=True, bf=True, tag=None, lo=False, m_dev=None, m_hint=0) -> Act:
        """nn.Linear + fused epilogue (include/navc.h navc_epilogue_t).  ``residual`` is an fp32 tensor
        or an Act; an Act without an fp32 copy is passed as its bf16 hi/lo pair (tcgen05 pair epilogue)."""
        M, N, K = x.M, lin.N, lin.K
        assert x.N == K, (x.N, K)
        use_tc = self.tc and (K % 64 == 0) and x.hi is not None
        out = self._new(M, N, f32 or not self.tc, bf, lo)
        res32 = res_hi = res_lo = None
        ld_res = 0
        if isinstance(residual, Act):
            ld_res = residual.N
            if residual.f32 is not None and (out.f32 is not None or not use_tc):
                res32 = residual.f32
            elif use_tc and out.f32 is None and residual.hi is not None:
                res_hi, res_lo = residual.hi, residual.lo
            else:
                res32 = self.join_f32(residual)
        elif residual is not None:
            res32, ld_res = residual, residual.shape[-1]
        ep = L.Epilogue(L.ptr(lin.b), L.ptr(res32), L.ptr(row_tokens), act, ld_res,
                        L.ptr(out.f32), L.ptr(out.hi), L.ptr(out.lo), N, 0, 1, 0, L.ptr(res_hi), L.ptr(res_lo),
                        m_dev.data_ptr() if m_dev is not None else None, int(m_hint), 0)
        e0 = self._t0(tag)
        if self.tf32 and x.f32 is not None and K % 32 == 0 and res_hi is None:
            L.call("navc_linear_tf32", L.ptr(x.f32), K, L.ptr(lin.w), K, M, N, K, ep, L.stream())
        elif use_tc:
            L.call("navc_linear_tc", self.tc_mode, L.ptr(x.hi), L.ptr(x.lo), K, L.ptr(lin.w_hi), L.ptr(lin.w_lo), K,
                   M, N, K, ep, L.stream())
        else:
            if x.f32 is None:
                raise L.NavcError("fp32 GEMM path needs an fp32 activation (K=%d not a multiple of 64?)" % K)
            L.call("navc_linear_f32", L.ptr(x.f32), K, L.ptr(lin.w), K, M, N, K, ep, L.stream())
        self._t1(tag, e0)
        return out

    def _t0(self, tag):
        if self.prof is None or tag is None:
            return None
        # inside a stream capture the events become event-record NODES of the graph (external=True): every replay
        # re-records them, so the per-class times come from the replayed graph itself, with no host gap between the
        # start event and the launch (eager timing charges each launch the host's latency to issue it)
        e0 = torch.cuda.Event(enable_timing=True, external=torch.cuda.is_current_stream_capturing())
        e0.record()
        return e0

    def _t1(self, tag, e0):
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True, external=torch.cuda.is_current_stream_capturing())
            e1.record()
            self.prof.setdefault(tag, []).append((e0, e1))

    def layernorm(self, x: Act, ln, row_tokens, f32=True, bf=True) -> Act:
        out = self._new(x.M, x.N, f32 or not self.tc, bf)
        L.call("navc_layernorm", L.ptr(x.f32), L.ptr(ln[0]), L.ptr(ln[1]), self.eps, L.ptr(row_tokens), x.M, x.N,
               L.ptr(out.f32), L.ptr(out.hi), L.ptr(out.lo), L.stream())
        return out

    def _proj_res(self, x: Act, lin, ln, residual: Act, row_tokens, pair=False, m_dev=None, tag=None, m_hint=0) -> Act:
        """dense -> (+residual) -> [LayerNorm] -> * non_pad_mask   (models/bert.py:192-200, 240-247, 271-299).
        pair: the residual stream lives as bf16 hi/lo pairs only (tensor-core modes without LayerNorm)."""
        if ln is None:
            return self.linear(x, lin, residual=residual, row_tokens=row_tokens, f32=not pair, bf=True, lo=pair, m_dev=m_dev, tag=tag,
                               m_hint=m_hint)
        y = self.linear(x, lin, residual=residual, row_tokens=None, f32=True, bf=False, tag=tag)
        return self.layernorm(y, ln, row_tokens)

    # ------------------------------------------------------------------------------------------
    # encoder  (models/seq2seq.py:35-63)
    # ------------------------------------------------------------------------------------------
    def encode(self, feats: List[torch.Tensor]):
        self.sync_weights()
        opt, P, D = self.opt, self.P, self.D
        assert len(feats) == len(P["streams"])
        if opt.get("fusion", "temporal_concat") not in ("temporal_concat", "none"):
            raise NotImplementedError("fusion=%r (broken in the reference, joint_representation.py:41)" % opt["fusion"])
        B = feats[0].shape[0]
        frames = [f.shape[1] for f in feats]
        E = sum(frames)
        dev = self.device
        enc = Act(B * E, D, f32=torch.empty((B, E, D), dtype=torch.float32, device=dev))
        if self.tc:
            enc.hi = torch.empty((B * E, D), dtype=torch.bfloat16, device=dev)
            enc.lo = torch.empty_like(enc.hi) if self.split else None
        enc_hidden = torch.empty((B, D), dtype=torch.float32, device=dev)
        no_norm = opt.get("fusion", "temporal_concat") == "none" or opt.get("no_encoder_bn", False)
        row0 = 0
        for i, (f, st) in enumerate(zip(feats, P["streams"])):
            F_ = f.shape[1]
            if len(set(frames)) != 1:
                raise NotImplementedError("modalities with different frame counts")
            x_in = self.from_f32(f.reshape(B * F_, f.shape[2]).to(dev))
            x = self.linear(x_in, st["l0"], f32=True, bf=True, tag="enc0")
            yg = self.linear(x, st["l12"], f32=True, bf=False, tag="enc12")
            norm = None if no_norm else P["norms"][i]
            if norm is not None and norm[0] == "ln":
                L.call("navc_highway_ln", L.ptr(x.f32), L.ptr(yg.f32), st["gate"], B, F_, D, E, i, len(feats), int(i > 0),
                       L.ptr(norm[1]), L.ptr(norm[2]), 1e-5, L.ptr(enc_hidden), L.ptr(enc.f32), L.ptr(enc.hi), L.ptr(enc.lo),
                       L.stream())
                row0 += F_
                continue
            rm = rv = bw = bb = None
            if norm is not None:
                _, rm, rv, bw, bb = norm
            L.call("navc_highway_bn", L.ptr(x.f32), L.ptr(yg.f32), st["gate"], B, F_, D, E, i, len(feats), int(i > 0),
                   L.ptr(rm), L.ptr(rv), L.ptr(bw), L.ptr(bb), 1e-5, L.ptr(enc_hidden), L.ptr(enc.f32),
                   L.ptr(enc.hi), L.ptr(enc.lo), L.stream())
            row0 += F_
        results = {}
        enc_mean = torch.empty((B, D), dtype=torch.float32, device=dev)
        if P["len_head"] is not None:
            w1, b1, w2, b2 = P["len_head"]
            max_len = w2.shape[0]
            pred = torch.empty((B, max_len), dtype=torch.float32, device=dev)
            L.call("navc_length_head", L.ptr(enc.f32), B, E, D, L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), max_len,
                   L.ptr(enc_mean), L.ptr(pred), L.stream())
            results["pred_length"] = pred
        else:
            L.call("navc_length_head", L.ptr(enc.f32), B, E, D, None, None, None, None, 0, L.ptr(enc_mean), None, L.stream())
        # the tensor handed to the caller is an ALIAS of enc.f32: Seq2Seq.encode hangs the cache below on it
        # (enc_output._navc_cache); with the same object that would be a reference cycle (tensor -> cache -> Act ->
        # tensor) and every batch's encoder memory + K|V projections (~440 MB at B = 128) would wait for the cyclic GC
        results["enc_output"] = enc.f32.detach()
        results["enc_hidden"] = enc_hidden
        results["_navc"] = dict(enc=enc, enc_mean=enc_mean, B=B, E=E, owner=id(self), pack_id=self.pack_id)
        return results

    # ------------------------------------------------------------------------------------------
    # decoder
    # ------------------------------------------------------------------------------------------
    def enc_inputs(self, enc_output: torch.Tensor, cache: Optional[dict] = None):
        """Encoder memory in the operand formats of this engine (fp32 + bf16 hi/lo copies) and the
        frame mean used by enhance_input=2; reuses what ``encode`` already produced."""
        self.sync_weights()
        if cache is not None and cache.get("owner") == id(self) and cache["enc"].f32.data_ptr() == enc_output.data_ptr() \
                and cache["enc"].f32.numel() == enc_output.numel():   # (a batch slice shares data_ptr with the full tensor)
            if cache.get("pack_id", self.pack_id) != self.pack_id:
                cache.pop("kv", None)       # K|V were projected with weights that have since changed
            cache["pack_id"] = self.pack_id
            return cache
        Bv, E, D = enc_output.shape
        enc = self.from_f32(enc_output.reshape(Bv * E, D))
        enc_mean = torch.empty((Bv, D), dtype=torch.float32, device=self.device)
        L.call("navc_length_head", L.ptr(enc.f32), Bv, E, D, None, None, None, None, 0, L.ptr(enc_mean), None, L.stream())
        return dict(enc=enc, enc_mean=enc_mean, B=Bv, E=E, owner=id(self), pack_id=self.pack_id)

    def memory(self, enc_output: torch.Tensor, cache: Optional[dict] = None):
        """Per-video decoder memory: cross-attention K|V of every layer (computed once)
        and the frame mean used by enhance_input=2."""
        cache = self.enc_inputs(enc_output, cache)
        if "kv" not in cache:
            tc_attn = self.tc_attention_ok(32, cache["E"])
            cache["kv"] = self.linear(cache["enc"], self.P["kv_all"], f32=not tc_attn, bf=tc_attn, tag="kv")  # [Bv*E, L*2D]
        return cache

    def tc_attention_ok(self, S, E):
        """tcgen05 attention cores need dk == 64, S <= 32 and E <= 128 (navc.h)."""
        return self.tc and self.opt["dim_hidden"] == self.opt["num_attention_heads"] * 64 and S <= 32 and E <= 128

    def _kv_f32(self, mem):
        if mem["kv"].f32 is None:  # fp32 cores requested (attention probabilities) after a tensor-core cache
            mem["kv"].f32 = self.linear(mem["enc"], self.P["kv_all"], f32=True, bf=False).f32
        return mem["kv"].f32

    def pack_rows(self, lens: torch.Tensor, S: int, hint: int = 0):
        """Packed-row bookkeeping for a fixed set of candidate lengths (include/navc.h "packed rows"):
        seq_off [N+1] (seq_off[N] = row count, device side), rowmap [N*S].  ``hint``: the host's estimate of the row
        count (navc_epilogue_t.m_hint: steers tile shapes only)."""
        N = lens.numel()
        seq_off = torch.empty((N + 1,), dtype=torch.int32, device=self.device)
        rowmap = torch.zeros((N * S,), dtype=torch.int32, device=self.device)
        L.call("navc_pack_rows", L.ptr(lens), N, S, L.ptr(seq_off), L.ptr(rowmap), L.stream())
        pk = dict(seq_off=seq_off, rowmap=rowmap, count=seq_off[N:], N=N, S=S, hint=int(hint))
        if os.environ.get("NAVC_ATTN2", "1") not in ("0", "no", "off"):
            # second-generation attention cores: 96-row windows of the packed row space (navc_pack_tiles)
            n_tiles = (N * S + L._lib.navc_attention_window() - 1) // L._lib.navc_attention_window()
            tile_seq = torch.empty((n_tiles + 1,), dtype=torch.int32, device=self.device)
            L.call("navc_pack_tiles", L.ptr(seq_off), N, L.ptr(tile_seq), n_tiles, L.stream())
            pk.update(tile_seq=tile_seq, n_tiles=n_tiles)
        return pk

    def can_pack(self, S, E):
        """Packed rows need the all-tensor-core layer (pair epilogues + tcgen05 attention cores)."""
        P = self.P
        return self.tc_attention_ok(S, E) and self.D % 64 == 0 and P["layers"][0]["f1"].N % 64 == 0 and \
            all(lw[k] is None for lw in P["layers"] for k in ("so_ln", "co_ln", "f2_ln"))

    def decoder_pass(self, tokens: torch.Tensor, mem: dict, group: int, category: Optional[torch.Tensor],
                     decoding_type: str, want_attn=False, want_f32=False, packed: Optional[dict] = None,
                     prune: Optional[dict] = None):
        """One BertDecoder forward (models/Decoder.py:96-178) -> hidden Act [N*S, D] (+ attention probs).
        ``prune`` (packed rows only; rows / count / seq_off of navc_compact_rows): the caller reads the hidden states of
        these rows only, so the last layer runs on them alone behind its self-attention core (which needs every row as
        key / value) and the result is the COMPACTED Act: row k = hidden state of packed row rows[k]."""
        P, D, H = self.P, self.D, self.H
        N, S = tokens.shape
        R = N * S
        E = mem["E"]
        assert N == mem["B"] * group, (N, mem["B"], group)
        tok_flat = tokens.reshape(R)
        emb = P["emb"]
        extra = None
        if decoding_type == "NARFormer":
            ei = self.opt.get("enhance_input", 2)
            if ei == 2:
                extra = mem["enc_mean"]
            elif ei != 0:
                raise NotImplementedError("enhance_input=1 fails in the reference itself (SURVEY 8c)")
        # residual stream as bf16 hi/lo pairs only (no fp32 copy written / re-read) when every GEMM of the
        # layer runs on the tensor cores and no per-sublayer LayerNorm needs the fp32 rows
        pair = self.tc and not self.tf32 and D % 64 == 0 and P["layers"][0]["f1"].N % 64 == 0 and \
            all(lw[k] is None for lw in P["layers"] for k in ("so_ln", "co_ln", "f2_ln"))
        x = self._new(R, D, not pair, True, lo=pair)
        m_dev, mh = None, 0
        if packed is not None:
            # only the sum(len) real positions are rows (R stays the launch maximum, the count is device side)
            assert (pair or self.tf32) and not want_attn and packed["N"] == N and packed["S"] == S
            m_dev, mh = packed["count"], packed.get("hint", 0)
            tok_flat = torch.empty((R,), dtype=torch.int64, device=self.device)  # token id of every packed row
            L.call("navc_embed_ln_packed", L.ptr(tokens), L.ptr(category), L.ptr(emb["word"]), L.ptr(emb["pos"]),
                   L.ptr(emb["cat"]), L.ptr(extra), group, L.ptr(emb["ln_w"]), L.ptr(emb["ln_b"]), self.eps, N, S, D,
                   L.ptr(packed["seq_off"]), L.ptr(packed["rowmap"]), L.ptr(tok_flat), L.ptr(x.f32), L.ptr(x.hi),
                   L.ptr(x.lo), L.stream())
        else:
            L.call("navc_embed_ln", L.ptr(tokens), L.ptr(category), L.ptr(emb["word"]), L.ptr(emb["pos"]), L.ptr(emb["cat"]),
                   L.ptr(extra), group, L.ptr(emb["ln_w"]), L.ptr(emb["ln_b"]), self.eps, N, S, D,
                   L.ptr(x.f32), L.ptr(x.hi), L.ptr(x.lo), L.stream())
        kv = mem["kv"]
        attns = []
        mask_kind = L.MASK_KIND[decoding_type]
        tc_attn = self.tc_attention_ok(S, E) and not want_attn and kv.hi is not None
        watch = int(self.opt.get("watch", 0))
        sfx = ""
        for l, lw in enumerate(P["layers"]):
            qkv = self.linear(x, lw["qkv"], f32=not tc_attn, bf=tc_attn, m_dev=m_dev, tag="qkv", m_hint=mh)
            ctx = self._new(R, D, not self.tc, True)
            p_self = p_cross = None
            e0 = self._t0("self")
            if packed is not None and "tile_seq" in packed:
                L.call("navc_self_attention_tc_tiles", self.tc_mode, L.ptr(qkv.hi), L.ptr(qkv.lo), 3 * D, L.ptr(tokens),
                       L.ptr(packed["seq_off"]), L.ptr(packed["tile_seq"]), packed["n_tiles"], R, N, S, D, H, mask_kind, watch,
                       L.ptr(ctx.hi), L.ptr(ctx.lo), L.stream())
            elif packed is not None:
                L.call("navc_self_attention_tc_packed", self.tc_mode, L.ptr(qkv.hi), L.ptr(qkv.lo), 3 * D, L.ptr(tokens),
                       L.ptr(packed["seq_off"]), N, S, D, H, mask_kind, watch, L.ptr(ctx.f32), L.ptr(ctx.hi), L.ptr(ctx.lo),
                       L.stream())
            elif tc_attn:
                L.call("navc_self_attention_tc", self.tc_mode, L.ptr(qkv.hi), L.ptr(qkv.lo), 3 * D, L.ptr(tokens), N, S,
                       D, H, mask_kind, watch, L.ptr(ctx.f32), L.ptr(ctx.hi), L.ptr(ctx.lo), L.stream())
            else:
                p_self = torch.empty((H, N, S, S), dtype=torch.float32, device=self.device) if want_attn else None
                L.call("navc_self_attention", L.ptr(qkv.f32), 3 * D, L.ptr(tokens), N, S, D, H, mask_kind,
                       watch, L.ptr(ctx.f32), L.ptr(ctx.hi), L.ptr(ctx.lo), L.ptr(p_self), L.stream())
            self._t1("self", e0)
            seq_off = packed["seq_off"] if packed is not None else None
            if prune is not None and l == len(P["layers"]) - 1:
                assert packed is not None and "tile_seq" in packed
                e0 = self._t0("gather")
                ctx_c, x_c = self._new(R, D, False, True, lo=ctx.lo is not None), self._new(R, D, False, True, lo=x.lo is not None)
                L.call("navc_gather_rows2", L.ptr(ctx.hi), L.ptr(ctx.lo), L.ptr(ctx_c.hi), L.ptr(ctx_c.lo), L.ptr(x.hi), L.ptr(x.lo),
                       L.ptr(x_c.hi), L.ptr(x_c.lo), D, L.ptr(prune["rows"]), L.ptr(prune["count"]), R, L.stream())
                self._t1("gather", e0)
                ctx, x, tok_flat = ctx_c, x_c, None          # (packed rows are never PAD: the row mask is a no-op)
                m_dev, mh, seq_off = prune["count"], prune.get("hint", 0), prune["seq_off"]
                sfx = "_p"                                   # profiling tags of the launches on the compacted rows
            a = self._proj_res(ctx, lw["so"], lw["so_ln"], x, tok_flat, pair, m_dev, tag="so" + sfx, m_hint=mh)
            q = self.linear(a, lw["cq"], f32=not tc_attn, bf=tc_attn, m_dev=m_dev, tag="cq" + sfx, m_hint=mh)
            ctx2 = self._new(R, D, not self.tc, True)
            e0 = self._t0("cross" + sfx)
            if packed is not None and "tile_seq" in packed:
                off = l * 2 * D
                L.call("navc_cross_attention_tc_tiles", self.tc_mode, L.ptr(q.hi), L.ptr(q.lo), D,
                       kv.hi[:, off:].data_ptr(), kv.lo[:, off:].data_ptr() if kv.lo is not None else None, kv.N,
                       L.ptr(seq_off), R, N, S, E, D, H, group, L.ptr(ctx2.hi), L.ptr(ctx2.lo), L.stream())
            elif packed is not None:
                off = l * 2 * D
                L.call("navc_cross_attention_tc_packed", self.tc_mode, L.ptr(q.hi), L.ptr(q.lo), D,
                       kv.hi[:, off:].data_ptr(), kv.lo[:, off:].data_ptr() if kv.lo is not None else None, kv.N,
                       L.ptr(packed["seq_off"]), N, S, E, D, H, group, L.ptr(ctx2.f32), L.ptr(ctx2.hi), L.ptr(ctx2.lo),
                       L.stream())
            elif tc_attn:
                off = l * 2 * D
                L.call("navc_cross_attention_tc", self.tc_mode, L.ptr(q.hi), L.ptr(q.lo), D,
                       kv.hi[:, off:].data_ptr(), kv.lo[:, off:].data_ptr() if kv.lo is not None else None, kv.N,
                       N, S, E, D, H, group, L.ptr(ctx2.f32), L.ptr(ctx2.hi), L.ptr(ctx2.lo), L.stream())
            else:
                p_cross = torch.empty((H, N, S, E), dtype=torch.float32, device=self.device) if want_attn else None
                kv_l = self._kv_f32(mem)[:, l * 2 * D:]
                L.call("navc_cross_attention", L.ptr(q.f32), D, kv_l.data_ptr(), kv.N, N, S, E, D, H, group,
                       L.ptr(ctx2.f32), L.ptr(ctx2.hi), L.ptr(ctx2.lo), L.ptr(p_cross), L.stream())
            self._t1("cross" + sfx, e0)
            c = self._proj_res(ctx2, lw["co"], lw["co_ln"], a, tok_flat, pair, m_dev, tag="co" + sfx, m_hint=mh)
            h = self.linear(c, lw["f1"], act=self.act, f32=not self.tc or self.tf32, bf=not self.tf32, tag="f1" + sfx, m_dev=m_dev, m_hint=mh)
            x = self._proj_res(h, lw["f2"], lw["f2_ln"], c, tok_flat, pair, m_dev, tag="f2" + sfx, m_hint=mh)
            if want_attn:
                attns.append((p_self, p_cross))
        if want_f32:
            self.join_f32(x)
        return x, attns

    def decoder_step(self, hist: torch.Tensor, anc: torch.Tensor, pos: int, caches, mem: dict, group: int,
                     category: Optional[torch.Tensor], decoding_type: str = "ARFormer") -> Act:
        """The causal BertDecoder for ONE new position per row (autoregressive beam search, decoding/ar_beam.py):
        position `pos` of every row's prefix, self-attention over the K/V cache (include/navc.h
        navc_self_attention_step) -> hidden Act [N, D].  Equals row `pos` of ``decoder_pass`` over the whole
        prefix (models/Decoder.py:96-178 with the causal mask): earlier positions never see later ones."""
        P, D, H = self.P, self.D, self.H
        N, T = hist.shape
        E = mem["E"]
        assert N == mem["B"] * group and decoding_type != "NARFormer"
        tokens = hist[:, pos].contiguous()  # this step's input token of every row
        emb = P["emb"]
        pair = self.tc and not self.tf32 and D % 64 == 0 and P["layers"][0]["f1"].N % 64 == 0 and \
            all(lw[k] is None for lw in P["layers"] for k in ("so_ln", "co_ln", "f2_ln"))
        x = self._new(N, D, not pair, True, lo=pair)
        # S = 1 with the position table offset to row `pos`
        L.call("navc_embed_ln", L.ptr(tokens), L.ptr(category), L.ptr(emb["word"]), emb["pos"][pos:].data_ptr(), L.ptr(emb["cat"]),
               None, group, L.ptr(emb["ln_w"]), L.ptr(emb["ln_b"]), self.eps, N, 1, D, L.ptr(x.f32), L.ptr(x.hi), L.ptr(x.lo),
               L.stream())
        kv = mem["kv"]
        # (one query row per beam: a 128-row tcgen05 tile holds only `group` rows, still 4x faster here than the
        # fp32 tile kernel -- 32 vs 128 us at 640 rows)
        tc_attn = self.tc_attention_ok(1, E) and kv.hi is not None
        watch = int(self.opt.get("watch", 0)) if decoding_type == "ARFormer" else 0
        for l, lw in enumerate(P["layers"]):
            qkv = self.linear(x, lw["qkv"], f32=True, bf=False)
            ctx = self._new(N, D, not self.tc, True)
            kc, vc = caches[l]
            L.call("navc_self_attention_step", L.ptr(qkv.f32), 3 * D, L.ptr(kc), L.ptr(vc), L.ptr(anc), L.ptr(hist), N, T, D, H,
                   pos, watch, L.ptr(ctx.f32), L.ptr(ctx.hi), L.ptr(ctx.lo), L.stream())
            a = self._proj_res(ctx, lw["so"], lw["so_ln"], x, tokens, pair)
            q = self.linear(a, lw["cq"], f32=not tc_attn, bf=tc_attn)
            ctx2 = self._new(N, D, not self.tc, True)
            if tc_attn:
                off = l * 2 * D
                L.call("navc_cross_attention_tc", self.tc_mode, L.ptr(q.hi), L.ptr(q.lo), D,
                       kv.hi[:, off:].data_ptr(), kv.lo[:, off:].data_ptr() if kv.lo is not None else None, kv.N,
                       N, 1, E, D, H, group, L.ptr(ctx2.f32), L.ptr(ctx2.hi), L.ptr(ctx2.lo), L.stream())
            else:
                kv_l = self._kv_f32(mem)[:, l * 2 * D:]
                L.call("navc_cross_attention", L.ptr(q.f32), D, kv_l.data_ptr(), kv.N, N, 1, E, D, H, group,
                       L.ptr(ctx2.f32), L.ptr(ctx2.hi), L.ptr(ctx2.lo), None, L.stream())
            c = self._proj_res(ctx2, lw["co"], lw["co_ln"], a, tokens, pair)
            h = self.linear(c, lw["f1"], act=self.act, f32=not self.tc or self.tf32, bf=not self.tf32)
            x = self._proj_res(h, lw["f2"], lw["f2_ln"], c, tokens, pair)
        return x

    # ------------------------------------------------------------------------------------------
    # vocabulary projection
    # ------------------------------------------------------------------------------------------
    def vocab_partials(self, hidden: Act, target: Optional[torch.Tensor] = None, m_dev: Optional[torch.Tensor] = None):
        """tgt_word_prj + softmax statistics without materialising logits (algorithms.py:7-15)."""
        lin = self.P["vocab"]
        R, V, K = hidden.M, lin.N, lin.K
        use_tc = self.tc and K % 64 == 0 and hidden.hi is not None
        tile = L._lib.navc_vocab_tile(1 if use_tc else 0)
        nt = (V + tile - 1) // tile
        dev = self.device
        pm = torch.empty((R, nt), dtype=torch.float32, device=dev)
        ps = torch.empty((R, nt), dtype=torch.float32, device=dev)
        pi = torch.empty((R, nt), dtype=torch.int32, device=dev)
        tl = torch.empty((R,), dtype=torch.float32, device=dev) if target is not None else None
        e0 = self._t0("vocab")
        if self.prof is not None and m_dev is not None:
            self.prof.setdefault("vocab_rows", []).append(m_dev.clone())  # device-side row count of this launch
        if m_dev is not None:
            assert use_tc and target is None
            L.call("navc_vocab_partials_tc_dyn", self.tc_mode, L.ptr(hidden.hi), L.ptr(hidden.lo), K, L.ptr(lin.w_hi),
                   L.ptr(lin.w_lo), K, L.ptr(lin.b), R, V, K, m_dev.data_ptr(), L.ptr(pm), L.ptr(ps), L.ptr(pi), L.stream())
        elif use_tc:
            L.call("navc_vocab_partials_tc", self.tc_mode, L.ptr(hidden.hi), L.ptr(hidden.lo), K, L.ptr(lin.w_hi),
                   L.ptr(lin.w_lo), K, L.ptr(lin.b), R, V, K, L.ptr(pm), L.ptr(ps), L.ptr(pi), L.ptr(target), L.ptr(tl),
                   L.stream())
        else:
            L.call("navc_vocab_partials_f32", L.ptr(hidden.f32), K, L.ptr(lin.w), K, L.ptr(lin.b), R, V, K,
                   L.ptr(pm), L.ptr(ps), L.ptr(pi), L.ptr(target), L.ptr(tl), L.stream())
        self._t1("vocab", e0)
        return pm, ps, pi, nt, tl

    def logits(self, hidden2d: torch.Tensor) -> torch.Tensor:
        """Materialised tgt_word_prj(hidden) for callers that use the attribute directly."""
        self.sync_weights()
        x = self.from_f32(hidden2d)
        return self.linear(x, self.P["vocab"], f32=True, bf=False).f32

    def log_softmax_(self, logits2d: torch.Tensor) -> torch.Tensor:
        M, V = logits2d.shape
        L.call("navc_log_softmax", L.ptr(logits2d), L.ptr(logits2d), M, V, V, L.stream())
        return logits2d
