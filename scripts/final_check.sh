#!/bin/bash
# Round-end validation on the GPU box: every GPU test file, smoke(), the default bench line, the per-workload lines and the ncu launch list.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
bash scripts/gpu_check.sh
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 4000 gpurun_out/bench_default.json
for w in matsed_finetune2 pmam dasm; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-extra-legs > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"; tail -c 1500 gpurun_out/bench_$w.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2_launches.csv
