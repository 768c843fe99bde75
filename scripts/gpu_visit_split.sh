mkdir -p gpurun_out
timeout 250 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 200 --tb=short -k "conv_tcgen05 or small_batch_split" 2>&1 | tail -4 | cut -c1-300
P='import json,sys; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][0]); print(round(d["value"],1), round(d["ms_per_step"],4), d.get("e2e",{}).get("value"), d.get("e2e",{}).get("latency_ms_median_of_5"))'
for v in 1 0; do echo "config 1 CN_ACT_SPLIT=$v: $(CN_ACT_SPLIT=$v timeout 120 python bench.py --config 1 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P")"; done
for v in 1 0; do echo "config 2 CN_ACT_SPLIT=$v: $(CN_ACT_SPLIT=$v timeout 150 python bench.py --steps 10 --warmup 4 --no-e2e --no-cpu-baseline 2>/dev/null | cut -c1-100)"; done
timeout 100 python scripts/gpu_generate_timeline.py 1 2>/dev/null > gpurun_out/r02_m8_generate_timeline_b1.txt; head -4 gpurun_out/r02_m8_generate_timeline_b1.txt | cut -c1-150
