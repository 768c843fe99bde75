#!/bin/bash
# bench lines of the BASELINE configurations that fit one GPU (1, 2, 3 and the per-GPU shards of 4 and 5) + the reference arm of each
TAG=${1:-r02}
CFGS=${2:-"1 3 4 5"}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for c in $CFGS; do
  timeout 900 python bench.py --config $c --steps ${STEPS:-5} --warmup 3 > gpurun_out/${TAG}_bench_cfg${c}.json 2> gpurun_out/${TAG}_bench_cfg${c}.err
  echo "config $c exit $?"; cut -c1-330 gpurun_out/${TAG}_bench_cfg${c}.json; tail -3 gpurun_out/${TAG}_bench_cfg${c}.err
done
