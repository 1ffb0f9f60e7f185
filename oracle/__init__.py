"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the similarity-weighted NT-Xent hot path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this package, and only as the checker.  The product
(`simhand_b200/`) never imports it: there is no CPU fallback.

Parity pin: the restatements here are checked against the reference's own
functions executed from `/root/reference/src/models/utils.py` in the build
container (`oracle/ref_loader.py`, `tests/test_oracle.py`) and against the
golden vectors those functions produced (`tests/golden/`, made by
`oracle/gen_golden.py`).  The reference ships no tests or golden vectors of its
own.
"""
