"""CPU oracle for the detectInBlur motion-blur hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs import it, and
only as the checker / the timed CPU baseline.  The product path (``detectinblur_b200``) never falls
back to this code; it raises if the CUDA library is missing.

Pinning status: the reference ships NO tests, golden vectors or known-answer fixtures for this path
(SURVEY.md section 4 / 8c), so the restatement is pinned against outputs of the reference itself,
executed in the authoring container by ``tools/make_golden.py`` (reference imported from
``/root/reference``), and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` replays
every fixture through this restatement (bit-exact for taps / fp32 / fp16 pixels / PSF rasters).
"""
