"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the cost-volume + disparity-regression path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as
the checker or the timed CPU baseline.  The product package
(``realtime_stereo_matcher_b200``) never imports this package and has no CPU fallback.

Parity pinning: the reference (babiking/realtime_stereo_matcher) has no tests, golden
vectors or known-answer files for this path, so the oracle is
pinned against *outputs of the reference itself*, executed in the build container by
``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every oracle function against those fixtures.
"""
from .numpy_oracle import *  # noqa: F401,F403
