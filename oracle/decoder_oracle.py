"""Greedy CTC decode oracle — TEST INFRASTRUCTURE ONLY.  Restates decoder.py:3-30 in numpy:
argmax over the vocabulary (first maximum wins, as numpy/torch CPU argmax do), keep a token iff
it is not blank and differs from the previous frame's token."""
import numpy as np


def ctc_greedy_decode(log_probs, input_lengths, blank=0):
    preds = np.argmax(np.asarray(log_probs, dtype=np.float64), axis=-1)
    out = []
    for b in range(preds.shape[0]):
        seq, prev = [], None
        for tok in preds[b, :int(input_lengths[b])]:
            tok = int(tok)
            if tok != blank and tok != prev:
                seq.append(tok)
            prev = tok
        out.append(seq)
    return out
