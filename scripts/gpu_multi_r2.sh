# multi-GPU checkpoint (run as: gpurun --gpus N -- bash scripts/gpu_multi_r2.sh N [full])
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1; echo "torchrun bench n=$N exit $?"; tail -1 gpurun_out/r2_bench_n$N.log | cut -c1-600
timeout 600 python bench.py --gpus $N --single-process --steps 10 --warmup 3 > gpurun_out/r2_bench_sp_n$N.log 2>&1; echo "single-process bench n=$N exit $?"; tail -1 gpurun_out/r2_bench_sp_n$N.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  scripts/check_sharded.py > gpurun_out/r2_check_sharded_n$N.log 2>&1; echo "check_sharded n=$N exit $?"; tail -2 gpurun_out/r2_check_sharded_n$N.log | cut -c1-400
if [ "${2:-}" = "full" ]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_n$N.log 2>&1; echo "all gpu tests n=$N exit $?"; tail -3 gpurun_out/r2_pytest_n$N.log | cut -c1-200
else
  timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "group or device_all" > gpurun_out/r2_pytest_group_n$N.log 2>&1; echo "group tests n=$N exit $?"; tail -3 gpurun_out/r2_pytest_group_n$N.log | cut -c1-200
fi
