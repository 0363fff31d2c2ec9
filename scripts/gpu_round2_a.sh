# GPU box: round-2 checkpoint A — gpu tests, default bench (both arms), sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/r2a_bench.log | cut -c1-600
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sweep 1,8,32,64,128,160,192,256,384,512,1024 > gpurun_out/r2a_sweep.log 2>&1; echo "sweep rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.log 2>&1; echo "ref rc $?"; tail -1 gpurun_out/r2a_ref.log | cut -c1-400
