# GPU box: ncu evidence for the DEFAULT bench command (run: gpurun -- bash scripts/gpu_profile.sh)
#  1. launch list (every search-pipeline launch with its device time)
#  2. one --set full capture of each scoring kernel: K-score CTA pairs at batch 4096 (main pass) and the
#     single-CTA K-score at batch 1 (main pass), plus K-pack
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
KERNELS='regex:score_tc|scan_bf16|select_|seed_|collect_|rerank_|finalize_|exact_|query_prep|merge_|normalize_|pack_rows'
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 1500 --csv --log-file gpurun_out/launches.csv \
    $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
# seeded searches launch (sampling pass, main pass): odd launch indices are main passes
ncu --set full --clock-control none --import-source on -k regex:score_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_score_tc2 \
    $CMD > gpurun_out/ncu_tc2.log 2>&1
echo "score_tc2 capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 11 -c 1 -o gpurun_out/prof_score_tc_b1 \
    $CMD > gpurun_out/ncu_tc_b1.log 2>&1
echo "score_tc (batch 1) capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:pack_rows_kernel -s 156 -c 1 -o gpurun_out/prof_pack \
    $CMD > gpurun_out/ncu_pack.log 2>&1
echo "pack capture exit $?"
ls -la gpurun_out/*.ncu-rep
