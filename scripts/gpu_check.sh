# GPU box: parity tests + smoke + default bench (run as `gpurun -- bash scripts/gpu_check.sh`)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "gpu tests exit $?"; tail -15 gpurun_out/t_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?"; tail -1 gpurun_out/bench_ref.log
fi
