# round-2 checkpoint: full GPU tests, default bench, sweep, launch list of the default bench command
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2l_pytest.log | cut -c1-220
python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/r2l_bench.log | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --hnsw-rows 0 --sweep 1,8,32,64,128,130,160,192,224,256,320,384,512,768,1024,2048 > gpurun_out/r2l_sweep.log 2>&1; echo "sweep rc $?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --hnsw-rows 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2l_launches.csv $CMD > gpurun_out/r2l_ncu_launch.log 2>&1
echo "launch list rc $?"
