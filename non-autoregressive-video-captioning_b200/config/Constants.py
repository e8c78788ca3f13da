"""Token ids and result-key mapping shared with callers (contract: reference config/Constants.py:1-18)."""
PAD, UNK, BOS, EOS, MASK, VIS = 0, 1, 2, 3, 4, 5

PAD_WORD, UNK_WORD, BOS_WORD = "<pad>", "<unk>", "<bos>"
EOS_WORD, MASK_WORD, VIS_WORD = "<eos>", "<mask>", "<vis>"

# criterion name -> (key of the model output, key of the target in the batch dict)
mapping = {
    "lang": ("tgt_word_logprobs", "tgt_word_labels"),
    "length": ("pred_length", "tgt_length"),
}
