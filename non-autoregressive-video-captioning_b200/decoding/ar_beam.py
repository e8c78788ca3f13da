"""Autoregressive beam search (contract: reference models/Translator.py:94-161 + models/Beam.py).

``beam_search(model, opt, encoder_outputs, category) -> (hyps, scores)`` with
``hyps[b] = [[token ids without BOS] x n_best]`` and ``scores[b] = [length-normalised log-prob] x n_best``.

SURVEY.md section 8(f) row 3: the AR path is not on the NA hot path and is not separately optimised.
Like the reference it re-runs the causal decoder over the whole prefix at every step (no K/V cache);
the decoder pass, the vocabulary projection and the log-softmax are navc kernels (encoder memory is
shared by the beams of a video through ``group = beam_size`` instead of being repeated), the top-k
over ``beam x vocab`` is one batched torch.topk, and the per-video beam bookkeeping (back-pointers,
finished list, length penalty: Beam.py:68-150) stays host-side Python as in the reference.
Finished videos are kept in the batch (their rows are ignored) instead of being compacted away.
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


def beam_search(model, opt, encoder_outputs, category):
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
