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

Modules and what pins them:
  blur_oracle.py     manual_blur / blur_image_list / normalize            blur_cases.npz, normalize_case.npz (bit-exact)
  psf_oracle.py      trajectory, PSF raster, centring, stored format      psf_cases.npz, transform_cases.npz (bit-exact)
                     philox4x32 + trajectory_philox: the counter-based walk the GPU trajectory generator evaluates --
                     not a reference function; Philox is pinned on the Random123 known answers, the walk on the
                     invariants and statistics of the reference walk (tests/test_gpu_transforms.py)
  fourier_oracle.py  BlurImageHandler (the CPU Fourier blur), as the FFT    fourier_cases.npz (<= 2e-6; 7 cases incl. odd sizes,
                     and restated as the tap sum the CUDA path evaluates    grey, upscaled, oldDeltaPad)
  resize_oracle.py   GeneralizedRCNNTransform.forward: normalize +          resize_cases.npz (<= 1e-6: torch contracts the
                     bilinear resize + zero-padded batch                    interpolation's multiply-adds)
"""
