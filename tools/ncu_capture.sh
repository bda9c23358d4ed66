#!/bin/bash
# ncu evidence for the round (run on the GPU box, one GPU):  bash tools/ncu_capture.sh
#  1. launch list of ONE sequential pass over the bench workload, plain stream launches (tools/one_pass.py; per-launch
#     device times are cold-cache and serialised: compare SHARES with the bench line, not absolutes)
#  2. `--set full` captures of the dominant GEMM class (m=14,n=13,k=10: the join profiles/gemm_traffic.json names),
#     a ~1 ms GEMM of the sliced plans (m=11,n=11,k=12), a store-bound join (k=4: the row-streamed persistent kernel; k=2: the
#     whole-tile one) and config 3's dominant join (m=11,n=10,k=10: the stream-K kernel).
# Outputs land in gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into the JSON summaries under profiles/.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 500 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file gpurun_out/launches.csv \
    python tools/one_pass.py 1 > gpurun_out/launches_pass.log 2>&1
echo "launch list rc=$?"
timeout 200 $NCU --set full --import-source on -k regex:k_gemm_dmma -s 2 -c 1 -o gpurun_out/gemm_14_13_10 -f python tools/one_join.py 14 13 10 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
timeout 200 $NCU --set full --import-source on -k regex:k_gemm_dmma -s 2 -c 1 -o gpurun_out/gemm_11_11_12 -f python tools/one_join.py 11 11 12 >> gpurun_out/ncu_gemm.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:k_gemm_dmma_p1 -s 2 -c 1 -o gpurun_out/gemm_14_14_4 -f python tools/one_join.py 14 14 4 >> gpurun_out/ncu_gemm.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:k_gemm_dmma_p -s 2 -c 1 -o gpurun_out/gemm_14_14_2 -f python tools/one_join.py 14 14 2 >> gpurun_out/ncu_gemm.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:k_gemm_dmma_sk -s 2 -c 1 -o gpurun_out/gemm_11_10_10 -f python tools/one_join.py 11 10 10 >> gpurun_out/ncu_gemm.log 2>&1
echo "captures done"; ls -la gpurun_out/*.ncu-rep
