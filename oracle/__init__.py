"""CPU oracle for the DV-Matcher hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU and in plain torch/numpy fp32 (fp64 where an
arbiter is needed), the algorithms of the reference's dense correspondence +
deformation path.  Every function cites the reference file:line it follows.

Who may import it: `tests/`, `__graft_entry__.smoke()`, and the `cpu_baseline`
/ `--impl reference` legs of `bench.py` -- as the checker or the timed CPU
baseline, never as the thing shipped.  Nothing under `dv_matcher_b200/`
imports it; the product path fails loudly when the CUDA library is missing.

Pinning status
--------------
The reference ships no tests, golden vectors or known-answer files
(SURVEY.md section 4 / 8c).  The oracle is therefore pinned against *outputs of
the reference itself*: `tests/golden/make_golden.py` imports the reference's
own `models/loss.py`, `lib/deformation_graph_point.py`, `models/model.py`
unmodified (through `oracle/refimport.py`'s stub-import hook, in the build
container where /root/reference exists), runs them on seeded inputs and real
SCAPE geometry, and commits the results as `tests/golden/*.npz`.
`tests/test_oracle_golden.py` checks every oracle function against them.

One exception: Chamfer (`oracle.geometry.chamfer_3d`).  Its arithmetic lives in
ThibaultGROUEIX/ChamferDistancePytorch (`chamfer3D/chamfer3D.cu`), which is NOT
vendored under /root/reference (git-ignored at `.gitignore:6`, no version
pinned in `requirements.txt`).  We restate its published algorithm (squared
direct-difference distance, ascending scan with strict `<`, int32 arg-min) and
anchor on the reference's call sites (`models/loss.py:1216-1226`, `867-882`).
For that one function: **parity unpinned** against upstream code.
"""
