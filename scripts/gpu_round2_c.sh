mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/r2c_pytest.log | cut -c1-200
python scripts/exp_pack.py > gpurun_out/r2c_pack.log 2>&1; cat gpurun_out/r2c_pack.log | cut -c1-200
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sweep 1,8,32,128 > gpurun_out/r2c_bench.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/r2c_bench.log | cut -c1-300
