mkdir -p gpurun_out
ARGS="--rows 10000000 --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_tc2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_score_tc2_full \
    python bench.py $ARGS > gpurun_out/ncu_tc2.log 2>&1
echo "capture exit $?"; tail -2 gpurun_out/ncu_tc2.log | cut -c1-300
