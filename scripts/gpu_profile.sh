# GPU box: ncu launch list of the default bench command + one full capture of each scoring kernel.
mkdir -p gpurun_out
SMALL="--rows 4000000 --batch 2048 --steps 2 --warmup 1 --no-cpu-baseline"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score_tc|scan_bf16|select_|rerank_|finalize_|exact_|query_prep|merge_|normalize_" -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
python bench.py $SMALL > gpurun_out/plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_tc2_kernel -s 2 -c 1 -o gpurun_out/prof_score_tc2 \
    python bench.py $SMALL > gpurun_out/ncu_tc.log 2>&1
echo "score_tc capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:scan_bf16_kernel -s 2 -c 1 -o gpurun_out/prof_scan \
    python bench.py $SMALL > gpurun_out/ncu_scan.log 2>&1
echo "scan capture exit $?"
ls -la gpurun_out
