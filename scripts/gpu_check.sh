#!/bin/bash
# Run on the GPU box via gpurun: each test file in its own process with a timeout (a trapped kernel kills only its file),
# then optional micro-benchmarks.  Everything lands in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
FILES=${T4S_TEST_FILES:-$(ls tests/test_*gpu*.py)}
rc_all=0
for f in $FILES; do
  name=$(basename "$f" .py)
  timeout ${T4S_TEST_TIMEOUT:-300} python -m pytest "$f" -m gpu -q ${T4S_PYTEST_ARGS:-} > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "$name rc=$rc $(tail -n 1 gpurun_out/$name.log)"
  [ $rc -ne 0 ] && rc_all=1 && tail -n 40 "gpurun_out/$name.log"
done
if [ -n "${T4S_EXTRA:-}" ]; then
  timeout ${T4S_EXTRA_TIMEOUT:-600} bash -c "$T4S_EXTRA" > gpurun_out/extra.log 2>&1
  echo "extra rc=$?"; tail -n 60 gpurun_out/extra.log
fi
exit $rc_all
