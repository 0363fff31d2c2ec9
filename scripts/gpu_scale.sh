# multi-GPU bench (run as: gpurun --gpus N -- bash scripts/gpu_scale.sh N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_n$N.log 2>&1; echo "bench n=$N exit $?"; tail -3 gpurun_out/bench_n$N.log | cut -c1-3000
