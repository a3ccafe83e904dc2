"""CPU oracle for the ML-VAE hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker or as the
reported CPU baseline -- never as the thing that is shipped or measured as the
B200 path.  The product package (``ml_vae_b200``) never imports this package
and raises when its CUDA library is missing.

Parity pinning status (see DESIGN.md "Oracle"):

* VAE latent block (FC stacks, reparameterise, KL, biLSTM decoder, recon loss,
  length-masked reduction, weighted loss sum): PINNED.  ``vae_ref.py`` is a
  functional restatement that ``oracle/gen_golden.py`` checks, in this
  container, against the reference's own modules imported unmodified from
  ``/root/reference/src`` (with the tiny ``sb_shim`` standing in for the one
  SpeechBrain symbol they import).  The golden vectors it writes under
  ``tests/golden/`` are outputs of the reference code itself.
* Joint boundary / mispronunciation decoder (``utils/decode_utils.py:374-565``): PINNED.
  ``decode_ref.py`` restates the dynamic programme; ``oracle/gen_golden_decode.py`` runs the
  reference FUNCTION itself (imported unmodified, numpy + joblib are installed here) on seeded
  batches and asserts that the restatement reproduces its three integer outputs bit for bit
  (``tests/golden/md_decode_cases.npz``).
* Acoustic front-end (SpeechBrain ``Fbank``): PARITY UNPINNED.  The arithmetic
  lives in the third-party ``speechbrain`` package (requirements.txt:1,
  unpinned, 0.5.x by API usage) which is neither vendored in the reference
  tree nor installed here, and the reference ships no feature fixtures.
  ``fbank_ref.py`` restates its published algorithm on top of ``torch.stft``;
  it is cross-checked against an independent float64 numpy DFT
  (``fbank_np.py``) and, as a third-party anchor, against ``torchaudio``'s
  Spectrogram / AmplitudeToDB(top_db=80) / compute_deltas where the definitions
  coincide (tests/test_oracle_golden.py::test_front_end_pieces_agree_with_torchaudio;
  the mel matrix differs from torchaudio's by design), but not against SpeechBrain
  itself -- so it stays "unpinned".
"""
