#!/bin/bash
# N=8 lines of the other BASELINE configurations (run with `gpurun --gpus 8 -- bash scripts/variants_n8.sh`): profiles/variants_r2.json
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
port=29520
for spec in "matsed_finetune2 0" "pmam 0" "dasm 0" "matsed 32"; do
  set -- $spec
  port=$((port + 1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 \
      --workload $1 --batch $2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_n8_$1_b$2.json 2> gpurun_out/bench_n8_$1_b$2.err
  echo "$1 batch=$2 rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_n8_$1_b$2.json')); print(round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms', d['config'].get('params_in_sync'))"
done
