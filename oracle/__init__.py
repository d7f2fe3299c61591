"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference algorithms on the LucyRNN + CTC (+RNN-T) hot
path.  Nothing under ``statecatcher_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may.  The product path never routes through here.

Parity status
-------------
* LucyRNN (lucy_oracle.py): PINNED — checked against golden vectors produced by
  importing the unmodified reference ``/root/reference/lucyrnn.py`` in the
  authoring container (tests/golden/make_golden.py, committed with its output).
* CTC (ctc_oracle.py): the reference calls ``torch.nn.CTCLoss`` (train.py:142,
  model.py:70-71); torch is unpinned in requirements.txt:5, installed 2.11.0.
  PINNED against ``torch.nn.functional.ctc_loss`` 2.11.0 CPU golden vectors.
* RNN-T (rnnt_oracle.py): PARITY UNPINNED by the reference (warp_rnnt is absent,
  not in requirements.txt, and its call site model.py:97-105 matches no
  published API).  Restates the Graves-2012 transducer DP and is cross-checked
  against ``torchaudio.functional.rnnt_loss`` and against the known-answer vector
  warp-transducer / warp-rnnt publish in their own tests (``rnnt_oracle.WARP_KAT_*``).
* Glue around the encoder: ``lucy_oracle.detach_states`` + the CTC composition, the RNN-T
  predictor/joiner (joiner_oracle.py, with a hand backward) and the greedy decoder
  (decoder_oracle.py) are PINNED on tests/golden/glue_cases.npz, produced by running the
  reference's own ``model.compute_loss`` / ``ASRModel`` / ``RNNTPredictorJoiner`` /
  ``RNNTCompactPredictorJoiner`` / ``decoder.ctc_greedy_decoder`` (tests/golden/make_glue_golden.py;
  ``xlstm`` stubbed, the Triton encoder replaced by the reference's native ``LucyRNN``).  The
  transducer loss VALUE under the joiner stays the torchaudio substitute pin.
* Frontend (frontend_oracle.py): PINNED on golden vectors from the reference's own torchaudio
  objects (tests/golden/make_frontend_golden.py).
"""
