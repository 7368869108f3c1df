#!/bin/bash
# compute-sanitizer pass over the kernels written in the second half of round 2 (pipelined attention incl. the query-row window and the
# TMA-store epilogue, last-layer row compaction, quarter-width GEMM tails, row-block flags + overlapped LayerNorm).  Bounded by `timeout`.
set -u
OUT=gpurun_out/sanitize_b
mkdir -p $OUT
SMOKE="python -c 'import __graft_entry__ as g; g.smoke()'"
run() {  # tool, tag, command
  timeout 300 compute-sanitizer --tool "$1" --print-limit 10 --error-exitcode 7 bash -c "$3" > "$OUT/$2.log" 2>&1
  echo "$2: exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/$2.log" | tail -1)" | tee -a "$OUT/summary.txt"
}
: > "$OUT/summary.txt"
run memcheck  memcheck_smoke  "$SMOKE"
run synccheck synccheck_smoke "$SMOKE"
run memcheck  memcheck_attn   "python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'mha_tc and (130 or 212 or 40 or 300)'"
run racecheck racecheck_attn  "python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'test_mha_tc_vs_fp64 and (130 or 212)'"
run memcheck  memcheck_flags  "python -m pytest tests/test_gpu_gemm_ln.py -x -q -m gpu -k 'overlapped and 978'"
cat "$OUT/summary.txt"
grep -h "Race reported\|Invalid\|at .*kernel" $OUT/racecheck_attn.log | sort | uniq -c | sort -rn | head -12
