"""CPU oracle for the ClairS-TO per-candidate hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``clairs_to_b200/`` may import this
package; the only allowed callers are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Every function restates a piece of the reference (HKU-BAL/ClairS-TO v0.4.4) and
cites the ``file:line`` it follows.  Parity pinning: the reference ships no
tests or golden vectors, so the oracle is pinned against
outputs of the reference's own Python code run in the build container
(``tests/golden/make_golden.py`` imports ``/root/reference`` and writes the
fixtures under ``tests/golden/``), and, when ``/root/reference`` is importable,
directly against the reference functions (``tests/test_oracle_vs_reference.py``).
"""
