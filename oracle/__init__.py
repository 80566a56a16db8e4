"""CPU oracle for the FRLW-EvD event-representation path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and there only as the checker
(or as the timed CPU baseline), never as a fallback for the CUDA path.

Parity status: PINNED.  The reference ships no golden vectors, so
the restatement is pinned against outputs of the reference itself, executed in
the build container by ``oracle/ref_harness.py`` + ``oracle/make_golden.py``;
the resulting vectors are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py`` on every run (CPU, no reference needed).
"""
