"""CPU oracle for the DCASE2019-task4 hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement (numpy float64 / plain torch
fp32) of the reference's algorithm for the path

    waveform -> log-mel -> CRNN fwd/bwd -> mean-teacher step.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or the timed CPU baseline.  The product package
(``dcase2019_task4_b200``) never imports it and fails loudly when its CUDA
library is missing.

Pinning status
--------------
The reference ships NO tests, golden vectors or fixtures, and
``librosa`` is not installed, so:

* ``oracle.crnn`` / ``oracle.train_step`` are pinned against the reference's
  own ``baseline/models/CRNN.py`` imported unmodified from ``/root/reference``
  (``tests/test_oracle_vs_reference.py``, and the committed fixtures made by
  ``tests/golden/make_golden.py``); ``oracle.train_step.train_batch`` additionally
  against the reference's own ``main.train`` / ``main_simple_CRNN.train`` run
  unmodified on the CPU (``tests/scripts/ref_train_vs_oracle.py``).
* ``oracle.mel.scaler_means`` / ``scaler_std`` are pinned against the reference's own ``baseline/utils/Scaler.py``
  (live and through ``tests/golden/scaler_reference.npz``).
* the rest of ``oracle.mel`` restates librosa's published algorithm (un-vendored, unpinned
  third-party dependency, environment.yml:17) and is cross-checked against
  ``torch.stft`` and ``torchaudio.functional.melscale_fbanks``:
  **parity unpinned** against librosa itself.
"""
