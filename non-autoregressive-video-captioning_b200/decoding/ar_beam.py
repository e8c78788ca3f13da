"""Autoregressive beam search (contract: reference models/Translator.py:94-161 + models/Beam.py).

``beam_search(model, opt, encoder_outputs, category) -> (hyps, scores)`` with
``hyps[b] = [[token ids without BOS] x n_best]`` and ``scores[b] = [length-normalised log-prob] x n_best``.

SURVEY.md section 8(f) row 3.  ``beam_search`` (default) keeps everything on the device: every step runs the
causal decoder for ONE new position per beam row over a self-attention K/V cache (``Engine.decoder_step``,
O(S) per step instead of re-running the whole prefix), the top-k over ``beam x vocab`` is one batched
torch.topk, and ``navc_beam_advance`` does Beam.advance (back-pointers as a re-gathered ancestry table,
finished list, stopping rules: Beam.py:68-117) for all videos in one launch; the host reads a done-counter
every 4 steps and the finished lists once at the end (length penalty + n-best sort: Beam.py:119-150).
Encoder memory is shared by the beams of a video through ``group = beam_size`` instead of being repeated;
finished videos stay in the batch (their rows are ignored) instead of being compacted away.

``beam_search_host`` (``opt['navc_ar_host_beam']``) is the first implementation -- whole-prefix decoder pass per
step, per-video Python bookkeeping as in the reference -- kept as a cross-check.
"""
from __future__ import annotations

import torch

from ..config import Constants


class _Beam:
    """Host-side state of one video's beam (reference models/Beam.py)."""

    def __init__(self, size, max_len, n_sents):
        self.size, self.max_len = size, max_len
        self.want = max(size, n_sents)
        self.done = False
        self.scores = [0.0] * size
        self.prev_ks, self.next_ys = [], [[Constants.BOS] + [Constants.PAD] * (size - 1)]
        self.finished = []

    def prefix(self, k):
        hyp = []
        for j in range(len(self.prev_ks) - 1, -1, -1):
            hyp.append(self.next_ys[j + 1][k])
            k = self.prev_ks[j][k]
        return [Constants.BOS] + hyp[::-1]

    def current(self):
        """[size, t] decoder input rows (Beam.get_current_state; beams are kept in score order)."""
        if len(self.next_ys) == 1:
            return [[t] for t in self.next_ys[0]]
        return [self.prefix(k) for k in range(self.size)]

    def advance(self, best_scores, best_ids, n_words):
        """Beam.advance after the device-side top-k (Beam.py:68-117)."""
        self.scores = list(best_scores)
        prev_k = [i // n_words for i in best_ids]
        self.prev_ks.append(prev_k)
        self.next_ys.append([i - k * n_words for i, k in zip(best_ids, prev_k)])
        for i, tok in enumerate(self.next_ys[-1]):
            if tok == Constants.EOS:
                self.finished.append([self.scores[i], len(self.next_ys) - 1, i])
                if len(self.finished) >= self.want:
                    self.done = True
                    return True
        if len(self.next_ys) == self.max_len:
            self.done = True
            if not self.finished:
                for i in range(self.size):
                    self.finished.append([self.scores[i], len(self.next_ys) - 1, i])
        return self.done

    def best(self, alpha, n_best):
        items = sorted(([sc / (t ** alpha), t, k] for sc, t, k in self.finished), key=lambda a: -a[0])[:n_best]
        hyps = []
        for _, t, k in items:
            hyp = []
            for j in range(t - 1, -1, -1):
                hyp.append(self.next_ys[j + 1][k])
                k = self.prev_ks[j][k]
            hyps.append(hyp[::-1])
        return hyps, [it[0] for it in items]


class _BeamState:
    """Device-side state of one beam search of shape (B videos, K beams, T = max_len): token history and ancestry
    (ping-pong), K/V caches, beam scores, finished lists.  Allocated once per shape and reset per call, so that the
    per-step CUDA graphs can be replayed on it."""

    def __init__(self, eng, B, K, T, n_best, E, has_cat):
        dev, D = eng.device, eng.D
        N = B * K
        self.B, self.K, self.T, self.N = B, K, T, N
        self.want = max(K, n_best)
        self.cap = self.want
        self.hist = [torch.zeros((N, T), dtype=torch.int64, device=dev) for _ in range(2)]
        self.anc = [torch.zeros((N, T), dtype=torch.int32, device=dev) for _ in range(2)]
        self.caches = [(torch.empty((T, N, D), dtype=torch.float32, device=dev),
                        torch.empty((T, N, D), dtype=torch.float32, device=dev)) for _ in eng.P["layers"]]
        self.scores = torch.zeros((B, K), dtype=torch.float32, device=dev)
        self.done = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.n_done = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.fin_count = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.fin_score = torch.zeros((B, self.cap), dtype=torch.float32, device=dev)
        self.fin_len = torch.zeros((B, self.cap), dtype=torch.int32, device=dev)
        self.fin_tok = torch.zeros((B, self.cap, T), dtype=torch.int64, device=dev)
        # graph-owned copies of the per-call inputs
        self.enc = torch.empty((B, E, D), dtype=torch.float32, device=dev)
        self.cat = torch.zeros((B, 1), dtype=torch.int64, device=dev) if has_cat else None
        self.mem = None
        self.pack_id = eng.pack_id
        self.graphs = {}      # "memory" / step t -> torch.cuda.CUDAGraph
        self.launches = {}    # our kernels inside each graph
        self.pool = None
        self.calls = 0

    def reset(self):
        for h in self.hist:
            h.zero_()
        self.hist[0][:, 0] = Constants.BOS   # only beam 0 is read at the first step (Beam.py:75-76)
        for t in (self.scores, self.done, self.n_done, self.fin_count):
            t.zero_()


def _beam_step(eng, st: _BeamState, t, V, decoding_type):
    """Step t (1-based): decode position t-1 of every beam row, top-K over beam x vocab, Beam.advance."""
    from .. import _lib as L
    B, K, T = st.B, st.K, st.T
    cur = (t - 1) % 2
    pos = t - 1
    hist, anc = st.hist[cur], st.anc[cur]
    x = eng.decoder_step(hist, anc, pos, st.caches, st.mem, K, st.cat, decoding_type)
    logits = eng.logits(eng.join_f32(x))                                        # tgt_word_prj, Translator.py:113
    best_scores = torch.empty((B, K), dtype=torch.float32, device=logits.device)
    best_ids = torch.empty((B, K), dtype=torch.int64, device=logits.device)
    if K <= 8:
        # log_softmax + beam score + EOS fill + top-K over beam x vocab in one launch (Translator.py:114, Beam.py:68-83)
        L.call("navc_beam_topk", L.ptr(logits), logits.stride(0), B, K, V, L.ptr(st.scores), L.ptr(hist), T, pos, int(t == 1),
               L.ptr(best_scores), L.ptr(best_ids), L.stream())
    else:
        logp = eng.log_softmax_(logits).view(B, K, V)
        if t == 1:
            lk = logp[:, 0, :]                                                  # only beam 0 holds <BOS> (Beam.py:75-76)
        else:
            lk = logp + st.scores.unsqueeze(-1)
            lk = lk.masked_fill(hist[:, pos].view(B, K, 1).eq(Constants.EOS), -1e20)   # Beam.py:71-74
            lk = lk.view(B, K * V)
        best_scores, best_ids = lk.topk(K, dim=1, largest=True, sorted=True)
    L.call("navc_beam_advance", L.ptr(best_scores), L.ptr(best_ids), B, K, V, t, T, st.want, T, L.ptr(hist),
           L.ptr(st.hist[1 - cur]), L.ptr(anc), L.ptr(st.anc[1 - cur]), L.ptr(st.scores), L.ptr(st.done), L.ptr(st.fin_count),
           L.ptr(st.fin_score), L.ptr(st.fin_len), L.ptr(st.fin_tok), st.cap, L.ptr(st.n_done), L.stream())


def _run_graphed(st: _BeamState, key, fn, capture):
    """Replay the graph of `fn` (capturing it first when allowed), else run it eagerly."""
    from .. import _lib as L
    g = st.graphs.get(key)
    if g is None and capture:
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        n0 = L.launches
        with torch.cuda.graph(g, pool=st.pool):
            fn()
        st.launches[key] = L.launches - n0
        if st.pool is None:
            st.pool = g.pool()
        st.graphs[key] = g
    if g is None:
        fn()
    else:
        g.replay()
        L.launches += st.launches[key]


def beam_search(model, opt, encoder_outputs, category):
    if opt.get("navc_ar_host_beam", False):
        return beam_search_host(model, opt, encoder_outputs, category)
    from .na_generate import graphs_enabled
    eng = model.engine
    eng.sync_weights()
    K = int(opt["beam_size"])
    max_len = int(opt["max_len"])
    n_best = int(opt.get("topk", 1))
    enc_output = encoder_outputs["enc_output"]
    if isinstance(enc_output, list):
        enc_output = enc_output[0]
    B, E = enc_output.shape[0], enc_output.shape[1]
    V = eng.P["vocab"].N
    decoding_type = opt.get("decoding_type", "ARFormer")
    cat = category.contiguous() if category is not None else None

    # state (and the CUDA graphs recorded on it) is kept per shape; dropped with the packed weights it points into
    key = ("ar_beam", B, K, max_len, n_best, E, decoding_type, cat is None, int(opt.get("watch", 0)))
    st = eng.graphs.get(key)
    if st is None or st.pack_id != eng.pack_id:
        for k in [k for k in eng.graphs if isinstance(k, tuple) and k and k[0] == "ar_beam"][1:]:
            del eng.graphs[k]   # each state owns its K/V caches: keep at most two shapes
        st = eng.graphs[key] = _BeamState(eng, B, K, max_len, n_best, E, cat is not None)
    st.calls += 1
    # first call per shape runs eagerly (warms lazily initialised kernels), the second records one graph per step
    capture = graphs_enabled(opt) and st.calls >= 2 and not torch.cuda.is_current_stream_capturing()
    st.reset()
    st.enc.copy_(enc_output.view(B, E, -1))
    if cat is not None:
        st.cat.copy_(cat.view(B, 1))

    def project_memory():  # cross-attention K|V of all layers, once per video
        st.mem = eng.memory(st.enc, None)

    _run_graphed(st, "memory", project_memory, capture)
    for t in range(1, max_len):
        _run_graphed(st, t, lambda: _beam_step(eng, st, t, V, decoding_type), capture)
        if t % 4 == 0 and t + 1 < max_len and int(st.n_done.item()) == B:   # every video finished early
            break
    alpha = opt.get("beam_alpha", 1.0)
    fc, fs, fl, ft = st.fin_count.tolist(), st.fin_score.tolist(), st.fin_len.tolist(), st.fin_tok.cpu()
    hyps, outs = [], []
    for b in range(B):  # Beam.sort_finished / get_hypothesis (Beam.py:119-150): length-normalised, best first, stable
        items = sorted(([fs[b][e] / (fl[b][e] ** alpha), e] for e in range(fc[b])), key=lambda a: -a[0])[:n_best]
        hyps.append([ft[b, e, :fl[b][e]].tolist() for _, e in items])
        outs.append([sc for sc, _ in items])
    beam_search.last_stats = {"steps": t, "graph": bool(st.graphs), "launches_per_step": st.launches.get(1)}
    return hyps, outs


beam_search.last_stats = {}


def beam_search_host(model, opt, encoder_outputs, category):
    eng = model.engine
    eng.sync_weights()
    n_bm = int(opt["beam_size"])
    max_len = int(opt["max_len"])
    enc_output = encoder_outputs["enc_output"]
    if isinstance(enc_output, list):
        enc_output = enc_output[0]
    B = enc_output.shape[0]
    dev = enc_output.device
    mem = eng.memory(enc_output.contiguous().float(), encoder_outputs.get("_navc"))
    cat = category.contiguous() if category is not None else None
    V = eng.P["vocab"].N
    beams = [_Beam(n_bm, max_len, int(opt.get("topk", 1))) for _ in range(B)]
    scores = torch.zeros((B, n_bm), dtype=torch.float32, device=dev)
    decoding_type = opt.get("decoding_type", "ARFormer")
    for t in range(1, max_len):
        rows = [r for b in beams for r in (b.current() if not b.done else [[Constants.BOS] + [Constants.PAD] * (t - 1)] * n_bm)]
        tokens = torch.tensor(rows, dtype=torch.int64).to(dev)              # [B*n_bm, t]
        hid, _ = eng.decoder_pass(tokens, mem, n_bm, cat, decoding_type, want_f32=True)
        last = hid.f32.view(B * n_bm, t, -1)[:, -1, :].contiguous()          # hidden state of the newest position
        logp = eng.log_softmax_(eng.logits(last)).view(B, n_bm, V)            # F.log_softmax(tgt_word_prj(.)), Translator.py:113-114
        if t == 1:
            lk = logp[:, 0, :]                                                 # only beam 0 holds <BOS> (Beam.py:75-76)
        else:
            lk = logp + scores.unsqueeze(-1)
            lk = lk.masked_fill(tokens[:, -1].view(B, n_bm, 1).eq(Constants.EOS), -1e20)   # Beam.py:71-74
            lk = lk.view(B, n_bm * V)
        best_scores, best_ids = lk.topk(n_bm, dim=1, largest=True, sorted=True)
        scores = best_scores
        bs, bi = best_scores.cpu().tolist(), best_ids.cpu().tolist()
        active = False
        for b, beam in enumerate(beams):
            if not beam.done:
                beam.advance(bs[b], bi[b], V)
                active = active or not beam.done
        if not active:
            break
    alpha = opt.get("beam_alpha", 1.0)
    out = [b.best(alpha, int(opt.get("topk", 1))) for b in beams]
    return [h for h, _ in out], [s for _, s in out]
