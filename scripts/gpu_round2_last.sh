#!/bin/bash
# Round-2 LAST measurement visit (shorter CPU arm, no HBM-kernel sweep): parity suite, smoke, the five bench configurations (+ the CPU arm of config 2), warm step
# timeline, ncu launch list of the bench command, ncu --set full of the tensor-core conv kernel.
#   gpurun --timeout 3000 -- 'bash scripts/gpu_round2_last.sh TAG'
TAG=${1:-r02_last}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/${TAG}_conv_breakdown.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
for c in 1 3 4 5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cfg${c}.json 2> gpurun_out/${TAG}_bench_cfg${c}.err
  echo "config $c exit $?"; cut -c1-220 gpurun_out/${TAG}_bench_cfg${c}.json
done
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
timeout 400 python scripts/gpu_step_timeline.py --steps 2 --out gpurun_out/${TAG}_step_timeline.txt 2>&1 | tail -1; head -1 gpurun_out/${TAG}_step_timeline.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_tc -c 8 -f -o gpurun_out/${TAG}_tc_full python scripts/ncu_conv_one.py > gpurun_out/${TAG}_ncu_full.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_full.log
ncu -i gpurun_out/${TAG}_tc_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_tc_full_raw.csv 2>/dev/null
python scripts/summarize_ncu_full.py gpurun_out/${TAG}_tc_full_raw.csv > gpurun_out/${TAG}_ncu_full_tc_summary.txt 2>&1; cut -c1-400 gpurun_out/${TAG}_ncu_full_tc_summary.txt | head -3
bash scripts/gpu_launch_list.sh ${TAG} 2>&1 | tail -4
