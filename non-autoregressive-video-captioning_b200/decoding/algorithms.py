"""Device-resident iterative-refinement algorithms (mask-predict, easy-first, left-to-right, with
coarse-grained templates and teacher re-scoring).

Behavioural contract: reference decoding/algorithms.py (MaskPredict :231-273, Left2Right :282-344,
EasyFirst :354-418, Algorithm_Base :27-222).  Unlike the reference there is no per-row Python loop
and no logits tensor: each decoder pass ends in `navc_vocab_partials_*` and every refinement
iteration is ONE `navc_refine_step` launch (merge + pad rules + re-mask selection + next canvas).
Mask-predict needs no host synchronisation at all; easy-first / left-to-right read one counter per
pass for their data-dependent stopping rule (as the reference does).
"""
from __future__ import annotations

import os

import torch

from .. import _lib as L
from ..config import Constants


class Refiner:
    def __init__(self, opt, model, teacher_model, mem, teacher_mem, category, beam, S, dict_mapping=None, rows_hint=0):
        self.opt, self.model, self.teacher = opt, model, teacher_model
        self.eng = model.engine
        self.mem, self.tmem = mem, teacher_mem
        self.category = category
        self.lens = beam.reshape(-1).contiguous()          # int32 [N]
        self.N, self.S = self.lens.numel(), S
        self.group = self.N // mem["B"]
        dev = self.eng.device
        N, S = self.N, S
        self.tokens = torch.empty((N, S), dtype=torch.int64, device=dev)
        self.canvas = torch.empty((N, S), dtype=torch.int64, device=dev)
        self.probs = torch.empty((N, S), dtype=torch.float32, device=dev)
        self.upd = torch.zeros((N, S), dtype=torch.uint8, device=dev)
        self.lprobs = torch.empty((N, S), dtype=torch.float32, device=dev)
        self.visual = torch.zeros((N, S), dtype=torch.uint8, device=dev)
        self.masked0 = torch.zeros((N, S), dtype=torch.uint8, device=dev)
        self.counters = torch.zeros((256, 2), dtype=torch.int32, device=dev)
        # packed rows (only the sum(len) real positions are decoder rows) when the whole layer runs on the
        # tensor cores; opt['navc_packed'] / $NAVC_PACKED = 0 keeps the reference's padded [N, S] layout
        want = opt.get("navc_packed", os.environ.get("NAVC_PACKED", "1"))
        self.packed = None
        if str(want).lower() not in ("0", "false", "no", "off") and self.eng.can_pack(S, mem["E"]):
            self.packed = self.eng.pack_rows(self.lens, S, rows_hint)
        # second-level packing of the vocabulary projection: logits only at the positions the previous step
        # re-masked (double-buffered row lists / slot maps; counts are device scalars)
        self.vsel = None       # (rows, count, slot) written by the last selecting step
        self.vbuf = None
        if self.packed is not None:
            R = N * S
            self.vbuf = [(torch.empty((R,), dtype=torch.int32, device=dev), torch.zeros((1,), dtype=torch.int32, device=dev),
                          torch.zeros((R + 1,), dtype=torch.int32, device=dev), torch.zeros((N + 1,), dtype=torch.int32, device=dev))
                         for _ in range(2)]
        # the LAST decoder layer of a pass that only merges re-masked positions runs on those rows alone (everything behind
        # its self-attention core; include/navc.h navc_compact_rows); opt['navc_prune'] / $NAVC_PRUNE = 0: all rows + a gather
        self.prune = str(opt.get("navc_prune", os.environ.get("NAVC_PRUNE", "1"))).lower() not in ("0", "false", "no", "off")
        self.rows_hint = int(rows_hint)
        self.n_steps = 0
        self.passes = 0
        self.pending = None  # (partials, merge kind, is_ct) of the pass not yet merged
        self.teacher_probs = None
        self.use_ct = bool(opt.get("use_ct", False))
        self.masking_decision = bool(opt.get("masking_decision", False)) and teacher_model is not None
        self.final_teacher = teacher_model is not None and not opt.get("no_candidate_decision", False)
        self.map = dict_mapping if isinstance(dict_mapping, torch.Tensor) else None  # id remap table (na_generate.py)

    # -- building blocks ---------------------------------------------------------------------
    def init_canvas(self, fill):
        L.call("navc_init_canvas", L.ptr(self.lens), self.N, self.S, int(fill), L.ptr(self.canvas),
               L.ptr(self.tokens), L.ptr(self.probs), L.stream())

    def run_pass(self, merge, is_ct=False):
        """decoder + vocabulary statistics on the current canvas (algorithms.py:143-167)."""
        sub_rows = merge == L.MERGE_MASKED and self.vsel is not None
        prune = None
        if sub_rows and self.prune:
            rows, count, slot, seq_off_c = self.vsel
            prune = dict(rows=rows, count=count, seq_off=seq_off_c, hint=self.rows_hint // 2)
        hid, _ = self.eng.decoder_pass(self.canvas, self.mem, self.group, self.category, "NARFormer", packed=self.packed, prune=prune)
        m_dev = self.packed["count"] if self.packed is not None else None
        slot = None
        if sub_rows:
            # only the re-masked positions are merged: project just their rows onto the vocabulary
            rows, count, slot, _ = self.vsel
            if prune is None:
                from ..engine import Act
                sub = Act(hid.M, hid.N, hi=torch.empty_like(hid.hi), lo=None if hid.lo is None else torch.empty_like(hid.lo))
                L.call("navc_gather_rows", L.ptr(hid.hi), L.ptr(hid.lo), hid.N, L.ptr(rows), L.ptr(count), hid.M,
                       L.ptr(sub.hi), L.ptr(sub.lo), L.stream())
                hid = sub
            m_dev = count
        self.vsel = None
        self.pending = (self.eng.vocab_partials(hid, m_dev=m_dev), merge, is_ct, slot)
        self.passes += 1

    def step(self, select, ratio=0.0, given=None, q=1, win=(0, 0), use_teacher=False, emit_flags=False):
        """One fused launch: merge the pending pass, pick the next positions to re-mask, write the
        next canvas.  Returns the index of the counter row written by this launch."""
        st = L.Step()
        if self.pending is not None:
            (pm, ps, pi, nt, _), merge, is_ct, part_slot = self.pending
            st.part_slot = L.ptr(part_slot)
            st.part_max, st.part_sum, st.part_idx, st.n_tiles = L.ptr(pm), L.ptr(ps), L.ptr(pi), nt
            st.merge, st.is_ct = merge, int(is_ct)
        else:
            st.merge, st.is_ct = L.MERGE_NONE, 0
        st.select, st.q, st.ratio = select, int(q), float(ratio)
        st.win_lo, st.win_hi = int(win[0]), int(win[1])
        st.lens = L.ptr(self.lens)
        st.teacher = L.ptr(self.teacher_probs) if use_teacher and self.teacher_probs is not None else None
        st.given = L.ptr(given)
        st.tokens, st.probs, st.upd_mask = L.ptr(self.tokens), L.ptr(self.probs), L.ptr(self.upd)
        st.canvas, st.lprobs = L.ptr(self.canvas), L.ptr(self.lprobs)
        slot = self.n_steps % self.counters.shape[0]
        if self.n_steps and slot == 0:
            self.counters.zero_()
        st.counters = self.counters[slot].data_ptr()
        st.visual = L.ptr(self.visual) if emit_flags else None
        st.masked0 = L.ptr(self.masked0) if emit_flags else None
        st.seq_off = L.ptr(self.packed["seq_off"]) if self.packed is not None else None
        if self.packed is not None and select in (L.SELECT_WORST, L.SELECT_MASKTOK, L.SELECT_GIVEN):
            rows, count, slot_map, seq_off_c = self.vbuf[self.n_steps % 2]
            st.sel_rows, st.sel_count, st.sel_slot = None, L.ptr(count), L.ptr(slot_map)   # flag mode
            self.vsel = (rows, count, slot_map, seq_off_c)
        L.call("navc_refine_step", st, self.N, self.S, L.stream())
        if self.vsel is not None and st.sel_slot:
            rows, count, slot_map, seq_off_c = self.vsel
            L.call("navc_compact_rows", L.ptr(slot_map), L.ptr(self.packed["seq_off"]), self.N, self.N * self.S, L.ptr(rows),
                   L.ptr(count), L.ptr(seq_off_c), L.stream())
        self.pending = None
        self.n_steps += 1
        return slot

    def score_with_teacher(self):
        """algorithms.py:175-204: causal teacher pass over [BOS]+tokens[:-1]; prob of each token."""
        N, S = self.N, self.S
        dev = self.eng.device
        shifted = torch.empty((N, S), dtype=torch.int64, device=dev)
        mapped = torch.empty((N, S), dtype=torch.int64, device=dev)
        L.call("navc_teacher_inputs", L.ptr(self.tokens), L.ptr(self.map), N, S, L.ptr(shifted), L.ptr(mapped), L.stream())
        teng = self.teacher.engine
        hid, _ = teng.decoder_pass(shifted, self.tmem, self.group, self.category, self.teacher.opt["decoding_type"])
        pm, ps, _, nt, tl = teng.vocab_partials(hid, target=mapped.view(-1))
        if self.teacher_probs is None:
            self.teacher_probs = torch.empty((N, S), dtype=torch.float32, device=dev)
        L.call("navc_teacher_probs", L.ptr(pm), L.ptr(ps), nt, L.ptr(tl), L.ptr(self.lens), N, S,
               L.ptr(self.teacher_probs), L.stream())

    def finish(self):
        """Merge the last pass, optional final teacher re-scoring, lprobs = log(prob * teacher)."""
        if self.final_teacher:
            if self.pending is not None:
                self.step(L.SELECT_KEEP)
            self.score_with_teacher()
            self.step(L.SELECT_NONE, use_teacher=True)
        else:
            self.step(L.SELECT_NONE)
        return self.tokens, self.lprobs

    def count(self, slot, which):
        return int(self.counters[slot, which].item())  # host sync (data-dependent loop control)

    # -- algorithms ---------------------------------------------------------------------------
    def mask_predict(self):
        T = int(self.opt.get("iterations", 5)) + (1 if self.use_ct else 0)
        self.init_canvas(Constants.VIS if self.use_ct else Constants.MASK)
        self.run_pass(L.MERGE_ALL, is_ct=self.use_ct)
        for t in range(1, T):
            if self.masking_decision:
                self.step(L.SELECT_KEEP)
                self.score_with_teacher()
            if self.use_ct and t == 1:
                self.step(L.SELECT_MASKTOK)
            else:
                self.step(L.SELECT_WORST, ratio=1.0 - (t / T), use_teacher=self.masking_decision)
            self.run_pass(L.MERGE_MASKED)
        return self.finish()

    def _start(self):
        """Common start of easy-first / left-to-right (algorithms.py:287-294, 359-366)."""
        if self.use_ct:
            self.init_canvas(Constants.VIS)
            self.run_pass(L.MERGE_ALL, is_ct=True)
        else:
            self.init_canvas(Constants.MASK)
        return self.step(L.SELECT_KEEP, emit_flags=True)

    def _refine_tail(self):
        Tq = int(self.opt.get("q_iterations", 1))
        visual = self.visual.clone() if self.use_ct else None
        for i in range(Tq):
            if i == 0 and self.use_ct:
                self.step(L.SELECT_GIVEN, given=visual)
            else:
                self.step(L.SELECT_WORST, ratio=0.4 * (1.0 - (i / Tq)))
            self.run_pass(L.MERGE_MASKED)
        return self.finish()

    def easy_first(self):
        q = int(self.opt.get("q", 1))
        slot = self._start()
        prev = 0
        while True:
            remain = self.count(slot, 0)
            if remain == 0 or remain == prev:
                break
            prev = remain
            self.run_pass(L.MERGE_EF)
            slot = self.step(L.SELECT_KEEP, q=q)
        return self._refine_tail()

    def left_to_right(self):
        q = int(self.opt.get("q", 1))
        self._start()
        masked0 = self.masked0.clone()
        for cur in range(0, self.S, q):
            slot = self.step(L.SELECT_WINDOW, given=masked0, win=(cur, cur + q))
            if self.count(slot, 1) == 0:
                break
            self.run_pass(L.MERGE_MASKED)
        return self._refine_tail()


ALGORITHMS = {"mp": Refiner.mask_predict, "ef": Refiner.easy_first, "l2r": Refiner.left_to_right}
