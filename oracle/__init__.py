"""CPU oracle for the CHORE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker (or as
the timed CPU baseline) -- never as part of the shipped product path
(``chore_b200/``), which must fail loudly when its CUDA library is missing.
"""
