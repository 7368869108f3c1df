#!/bin/bash
# compute-sanitizer over the hot path (run on the GPU box: gpurun -- 'bash tools/sanitize.sh').  memcheck + racecheck + synccheck of
# one small CMDM chain + CDM forward (smoke) and of the collapsed-CDM / fast-mode GPU tests; summaries -> gpurun_out/sanitize/,
# the committed copy lives in profiles/r2_sanitizer_summary.txt.  Every tool run is bounded by `timeout`.
set -u
OUT=gpurun_out/sanitize
mkdir -p $OUT
SMOKE="python -c 'import __graft_entry__ as g; g.smoke()'"
run() {  # tool, tag, command
  timeout 420 compute-sanitizer --tool "$1" --print-limit 20 --error-exitcode 7 bash -c "$3" > "$OUT/$2.log" 2>&1
  echo "$2: exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/$2.log" | tail -1)" | tee -a "$OUT/summary.txt"
}
: > "$OUT/summary.txt"
run memcheck  memcheck_smoke  "$SMOKE"
run racecheck racecheck_smoke "$SMOKE"
run synccheck synccheck_smoke "$SMOKE"
run memcheck  memcheck_cdm    "python -m pytest tests/test_gpu_cdm_collapsed.py -x -q -m gpu -k 'collapsed_forward and (1000 or 77)'"
run racecheck racecheck_cdm   "python -m pytest tests/test_gpu_cdm_collapsed.py -x -q -m gpu -k 'collapsed_forward and 1000'"
run memcheck  memcheck_fast   "python -m pytest tests/test_gpu_fast_mode.py -x -q -m gpu -k 'linear_tc_fast and 392'"
cat "$OUT/summary.txt"
