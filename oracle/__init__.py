"""oracle/ -- TEST INFRASTRUCTURE, not product code.

A CPU restatement (plain torch ATen ops, the arithmetic the reference itself runs on:
`requirements.txt:19` pins torch) of the reference's CTR forward hot path, used ONLY by
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs as the *checker* and the *reported CPU baseline*.  Nothing under `torecsys_b200/`
imports it; the product path has no CPU fallback and raises when the CUDA library is absent.

Parity pinning: the reference's own tests for this path are shape-only, so
there are no upstream golden vectors.  The restatement is pinned instead against outputs of the
reference itself, imported in the build container through `oracle/ref_shim.py` and frozen as
fixtures under `tests/golden/` by `oracle/make_golden.py` (committed with the fixtures).
"""
