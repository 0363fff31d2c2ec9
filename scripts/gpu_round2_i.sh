mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc $?"; tail -8 gpurun_out/r2i_pytest.log | cut -c1-220
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sweep 1,8,32,64,128,160,192,256,384,512,1024 > gpurun_out/r2i_bench.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/r2i_bench.log | cut -c1-300
python scripts/exp_power.py > gpurun_out/r2i_power.log 2>&1; cut -c1-300 gpurun_out/r2i_power.log
timeout 900 python scripts/exp_tn.py > gpurun_out/r2i_tn.log 2>&1; echo "tn rc $?"
