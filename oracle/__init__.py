"""TEST INFRASTRUCTURE ONLY — CPU/fp32 restatement of the reference hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the timed CPU baseline.
The product package (``video_depth_normal_v2_b200``) never imports it and fails
loudly when its CUDA library is missing.

Parity pinning: the reference ships no golden vectors or tests, so the
oracle is pinned against outputs of the *live reference* run in the build container
(``tests/golden/gen_golden.py`` -> ``tests/golden/*.npz``); see ``tests/test_oracle_golden.py``.
"""
