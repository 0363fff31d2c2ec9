mkdir -p gpurun_out
set -x
timeout 600 python bench.py --steps 5 --warmup 3 --batch 4 --path 1 --no-cpu-baseline > gpurun_out/bench_scan.log 2>&1; echo "bench scan exit $?"; tail -2 gpurun_out/bench_scan.log
timeout 900 python bench.py --steps 5 --warmup 3 --sweep 1,2,4,8,16,32,64,128,256,512,1024,2048 > gpurun_out/bench_tc.log 2>&1; echo "bench tc exit $?"; tail -2 gpurun_out/bench_tc.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?"; tail -1 gpurun_out/bench_ref.log
