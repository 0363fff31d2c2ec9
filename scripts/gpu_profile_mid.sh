# GPU box: ncu evidence for the 129..512-query regime on a 10 M-row shard
#   gpurun --timeout 1500 -- bash scripts/gpu_profile_mid.sh
mkdir -p gpurun_out
rm -f gpurun_out/prof_mid_*.ncu-rep
for spec in "256 score_tc2_kernel" "384 score_tc_kernel" "512 score_tc2_kernel"; do
  set -- $spec
  ARGS="--rows 10000000 --batch $1 --steps 2 --warmup 1 --no-cpu-baseline"
  python bench.py $ARGS > gpurun_out/plain_mid_$1.log 2>&1 || { echo "plain bench $1 failed"; tail -3 gpurun_out/plain_mid_$1.log; continue; }
  tail -1 gpurun_out/plain_mid_$1.log | cut -c1-300
  # seeded searches launch (sampling pass, main pass): odd launch indices are main passes
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/prof_mid_$1 \
      python bench.py $ARGS > gpurun_out/ncu_mid_$1.log 2>&1
  echo "capture batch $1 exit $?"
done
ls -la gpurun_out/*.ncu-rep
