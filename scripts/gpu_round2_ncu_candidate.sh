#!/bin/bash
# ncu --set full of the candidate kernel next to the production kernel on VGG block3 (B=16, 64x64, 256 -> 256):
#   gpurun --timeout 600 -- 'bash scripts/gpu_round2_ncu_candidate.sh'
# then here:  ncu -i gpurun_out/r02_candidate_full.ncu-rep --page raw --csv > /tmp/c.csv; python scripts/summarize_ncu_full.py /tmp/c.csv
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CN_PROBE_NCU=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv_tma_fast|igemm_tc_pixel" -c 6 -f \
    -o gpurun_out/r02_candidate_full python scripts/gpu_probe_round2.py > gpurun_out/r02_candidate_ncu.log 2>&1
tail -5 gpurun_out/r02_candidate_ncu.log | cut -c1-240
