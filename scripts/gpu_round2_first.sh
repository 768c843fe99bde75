#!/bin/bash
# First GPU visit of the next round (about 6 GPU-minutes):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round2_first.sh'
# 1. the design probes (scripts/gpu_probe_round2.py: tf32 raw-operand handling, TMA tiles as SAME padding, the TMA-fed
#    candidate kernel against production) - under their own timeout, they have never run on hardware;
# 2. the parity suite (the Python-side changes made after the last validated tree: checkpoint order of the real encoder,
#    set-up draws and checkpoint cadence in train(), read-only memmap uploads);
# 3. one bench line.
TAG=${1:-r02_first}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python scripts/gpu_probe_round2.py > gpurun_out/${TAG}_probes.log 2>&1
echo "probes exit $?"; tail -25 gpurun_out/${TAG}_probes.log | cut -c1-260
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-400 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
