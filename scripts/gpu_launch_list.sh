#!/bin/bash
# ncu launch list of one bench step (every launch visible: CN_GRAPHS=0) + summary.   bash scripts/gpu_launch_list.sh TAG
TAG=${1:-r02}
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CN_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --roofline-pass inline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-200
python scripts/summarize_launches.py gpurun_out/${TAG}_launches.csv 2 gpurun_out/${TAG}_tc_dram_traffic.json > gpurun_out/${TAG}_launches_summary.txt
head -45 gpurun_out/${TAG}_launches_summary.txt | cut -c1-150; tail -3 gpurun_out/${TAG}_launches_summary.txt
