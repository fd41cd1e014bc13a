#!/bin/bash
# compute-sanitizer over the asynchronous-machinery kernels at small shapes (run on the GPU box: gpurun -- bash scripts/sanitize.sh).
# memcheck: out-of-bounds / misaligned accesses;  racecheck: shared-memory hazards between the warp roles;  synccheck: barrier misuse.
# Summaries land in gpurun_out/sanitize_<tool>.log; copy the tails into profiles/ when they are to be judged.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
rc_all=0
for tool in ${T4S_SANITIZE_TOOLS:-memcheck racecheck synccheck}; do
  timeout ${T4S_SANITIZE_TIMEOUT:-900} compute-sanitizer --tool "$tool" --error-exitcode 7 --print-limit 20 python scripts/sanitize_cases.py \
      > "gpurun_out/sanitize_$tool.log" 2>&1
  rc=$?
  echo "compute-sanitizer $tool rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -n 1)"
  grep -E "cases ok" "gpurun_out/sanitize_$tool.log" | tr '\n' ' '; echo
  [ $rc -ne 0 ] && rc_all=1 && grep -E "=========" "gpurun_out/sanitize_$tool.log" | head -n 30
done
exit $rc_all
