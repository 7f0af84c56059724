"""CPU oracle for the MSAU hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement (numpy / torch-fp32 / plain C) of
what datvo06/MSAU computes on the hot path.  It exists so that ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` can check (and time) the reference algorithm on a box where
``/root/reference`` does not exist.  Nothing in ``msau_b200/`` (the product) may
import it: the product path is CUDA-only and fails loudly without its extension.

Parity status: PINNED.  The reference ships no tests / golden vectors (SURVEY.md
section 4), so the oracle is pinned by running the unmodified reference itself in the
build container (``tests/golden/make_golden.py`` imports ``/root/reference``) and
committing its outputs as fixtures under ``tests/golden/``; ``tests/test_oracle_*.py``
replays them against this restatement (bit-exact for grids / label maps, exact-equal
float32 for the model because both sides call the same ATen CPU operators).
"""
