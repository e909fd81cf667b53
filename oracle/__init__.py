"""CPU oracle for the max-cut / QUBO environment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rlsolver_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs
of ``bench.py`` do, and there only as the checker / the timed CPU baseline.

Parity status: pinned.  Every function here is checked bit-for-bit (integers)
or to 1e-6 (floats) against fixtures under ``tests/golden/`` that were produced
by importing the reference itself (``tools/make_goldens.py``, CPU, fixed seeds,
RNG draws recorded) -- see ``tests/test_oracle_golden.py`` -- plus the one
known answer the reference documents (greedy on BA_100_ID0 -> 273).
"""
