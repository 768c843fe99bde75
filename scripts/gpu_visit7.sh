cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for l in 2 4 8; do
  echo "== LPR $l"
  CN_LIB=$GRAFT_REPO_ROOT/confignet_b200/lib/libcn_lpr$l.so timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -k "conv" 2>&1 | tail -3
  CN_LIB=$GRAFT_REPO_ROOT/confignet_b200/lib/libcn_lpr$l.so CN_DBG=32 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/r02_breakdown_lpr$l.txt 2>/dev/null | cut -c1-120
done
echo "== LPR 1"
CN_DBG=32 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/r02_breakdown_lpr1.txt 2>/dev/null | cut -c1-120
