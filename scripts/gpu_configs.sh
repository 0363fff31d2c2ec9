# GPU box: the other BASELINE configs on one GPU (run as: gpurun -- bash scripts/gpu_configs.sh)
mkdir -p gpurun_out
for c in ${CONFIGS:-2 4 4s 5}; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_cfg$c.log 2>&1
  echo "config $c exit $?"; tail -1 gpurun_out/bench_cfg$c.log | cut -c1-2500
done
