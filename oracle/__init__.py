"""CPU oracle for the Transformer4SED hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker / timed CPU baseline.  Nothing under
``transformer4sed_b200/`` imports it; the product path fails loudly if the CUDA library is missing.

Parity status: PINNED.  The reference is pure Python/PyTorch and importable in the build container
with the shim in ``oracle/ref_shim``; ``oracle/make_golden.py`` runs the *unmodified* reference
(`/root/reference`) and commits its outputs under ``tests/golden/``.  ``tests/test_oracle_*.py`` check
every function here against those vectors.  The reference itself ships no tests / golden vectors
(SURVEY §4), and the arithmetic below that lives in third-party packages is restated from their
published algorithms: timm==0.4.5 (`requirements.txt:16`: Block/Attention/Mlp), torchaudio
(`compliance.kaldi.get_mel_banks`, Kaldi mel-bank construction), torch (`stft`, `LayerNorm`,
`MultiheadAttention`, `interpolate`).
"""
