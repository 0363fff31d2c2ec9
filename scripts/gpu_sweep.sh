# GPU box: batch sweep on one GPU (run as: gpurun -- bash scripts/gpu_sweep.sh)
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sweep ${SWEEP:-1,2,4,8,16,32,64,128,129,192,256,384,512,768,1024,2048} > gpurun_out/bench_sweep.log 2>&1; echo "bench sweep exit $?"; tail -1 gpurun_out/bench_sweep.log | cut -c1-6000
