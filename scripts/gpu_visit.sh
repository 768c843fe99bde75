#!/bin/bash
# One GPU visit: the parity suite (optionally a -k filter), then one bench line.   gpurun -- 'bash scripts/gpu_visit.sh TAG ["pytest args"] [bench args]'
TAG=${1:-r02}
PYARGS=${2:-}
BARGS=${3:---steps 10 --warmup 3}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short -x $PYARGS > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-300
if [ "$BARGS" != "none" ]; then
  timeout 600 python bench.py $BARGS > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  cut -c1-600 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
fi
