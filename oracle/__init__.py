"""ORACLE — test infrastructure only.

CPU restatement (torch fp32 / numpy fp64 / plain C) of the afford-motion diffusion hot path,
used ONLY as the checker by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`.  Nothing in `afford-motion_b200/` imports this package;
the product path fails loudly when its CUDA library is missing instead of falling back here.

Pinning: every function here is checked against outputs of the reference's own Python modules
(imported from /root/reference in the build container by `tests/golden/make_golden.py`, fixtures
committed under `tests/golden/`).  Two boundaries are PARITY UNPINNED because their source is not
under /root/reference: `pointops_cuda` (FPS / kNN tie order, see pointops_ref.c) and CLIP
(replaced by a synthetic text-feature provider behind the `encode_text_clip` hook).
"""
