#!/bin/bash
# One GPU-box visit: GPU tests, bench (both arms), ncu launch list of a bench step, ncu --set full of the hot kernels.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
TAG=${T4S_TAG:-r1}
bash scripts/gpu_check.sh
rc=$?
if [ "${T4S_SKIP_BENCH:-0}" != "1" ]; then
  timeout 600 python bench.py ${T4S_BENCH_ARGS:-} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -n 5 gpurun_out/bench_${TAG}.err
fi
if [ "${T4S_REF:-0}" = "1" ]; then
  timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1
  tail -c 1500 gpurun_out/bench_ref_${TAG}.json
fi
if [ "${T4S_NCU_LIST:-0}" = "1" ]; then
  T4S_BREAKDOWN=/dev/null timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_${TAG}.log 2>&1
  echo "ncu list rc=$?"
fi
for k in ${T4S_NCU_FULL:-}; do
  # k = name:regex:skip:count
  IFS=: read name regex skip count <<< "$k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -f -o gpurun_out/${TAG}_$name \
      python scripts/prof_kernels.py $name > gpurun_out/ncu_full_${TAG}_$name.log 2>&1
  echo "ncu full $name rc=$?"
done
exit $rc
