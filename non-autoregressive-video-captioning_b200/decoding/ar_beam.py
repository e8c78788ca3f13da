"""Autoregressive beam search (reference models/Translator.py:94-161 + models/Beam.py).
SURVEY.md section 8(f) item 3 ("next" row, not on the NACF hot path): not built yet."""


def beam_search(model, opt, encoder_outputs, category):
    raise NotImplementedError("AR beam search (ARB inference) is a section-8(f) 'next' row and is not built yet; "
                              "the AR decoder itself (teacher re-scoring, ARFormer forward) is supported")
