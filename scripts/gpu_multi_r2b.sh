# single-process group checkpoint (run as: gpurun --gpus N -- bash scripts/gpu_multi_r2b.sh N)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python bench.py --gpus $N --single-process --steps 10 --warmup 3 > gpurun_out/r2_bench_sp_n$N.log 2>&1; echo "single-process bench n=$N exit $?"; tail -1 gpurun_out/r2_bench_sp_n$N.log | cut -c1-1500
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q > gpurun_out/r2_pytest_group_n$N.log 2>&1; echo "round2 tests n=$N exit $?"; tail -5 gpurun_out/r2_pytest_group_n$N.log | cut -c1-300
