# GPU box: refresh of the launch list and the large-batch tail capture after the one-wave sampling pass and the
# 3-CTAs-per-SM tail instantiation (run: gpurun --timeout 1500 -- bash scripts/gpu_profile_r2b.sh)
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --hnsw-rows 0 --sweep none"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail_kernel -s 5 -c 1 -o gpurun_out/prof_tail_dense_b4096 $CMD > gpurun_out/ncu_tail_dense.log 2>&1
echo "capture tail dense exit $?"
ls -la gpurun_out/*.ncu-rep
