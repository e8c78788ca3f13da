"""Helpers callers import from misc.utils (contract: reference misc/utils.py:21-30, 205-229)."""
from ..config import Constants


def to_sentence(hyp, vocab, break_words=(Constants.EOS, Constants.PAD), skip_words=()):
    words = []
    for wid in hyp:
        if wid in skip_words:
            continue
        if wid in break_words:
            break
        words.append(vocab[wid])
    return " ".join(words)


def enlarge(info, beam_size):
    """[B, ...] -> [B*beam_size, ...], candidate-major (row = b*beam_size + j)."""
    return info.unsqueeze(1).expand(info.shape[0], beam_size, *info.shape[1:]).reshape(
        info.shape[0] * beam_size, *info.shape[1:])


def auto_enlarge(info, beam_size):
    if isinstance(info, list):
        return [auto_enlarge(x, beam_size) for x in info]
    if isinstance(info, tuple):
        return tuple(auto_enlarge(x, beam_size) for x in info)
    return enlarge(info, beam_size)
