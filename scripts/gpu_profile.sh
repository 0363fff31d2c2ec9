# GPU box: ncu evidence for the DEFAULT bench command (run: gpurun -- bash scripts/gpu_profile.sh)
#  1. launch list (every search-pipeline launch with its device time)
#  2. one --set full capture of each scoring kernel (K-score pairs at batch 4096, K-scan at batch 1)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
KERNELS='regex:score_tc|scan_bf16|select_|seed_|rerank_|finalize_|exact_|query_prep|merge_|normalize_|pack_rows'
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 1200 --csv --log-file gpurun_out/launches.csv \
    $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
# main (seeded) pass of the pair kernel = every second score_tc2 launch; skip 3 -> a steady-state main pass
ncu --set full --clock-control none --import-source on -k regex:score_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_score_tc2 \
    $CMD > gpurun_out/ncu_tc2.log 2>&1
echo "score_tc2 capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:scan_bf16_kernel -s 5 -c 1 -o gpurun_out/prof_scan \
    $CMD > gpurun_out/ncu_scan.log 2>&1
echo "scan capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:pack_rows_kernel -s 170 -c 1 -o gpurun_out/prof_pack \
    $CMD > gpurun_out/ncu_pack.log 2>&1
echo "pack capture exit $?"
ls -la gpurun_out | head -30
