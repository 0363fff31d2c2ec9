# round-2 final single-GPU checkpoint: GPU tests, default bench, reference arm, smoke, sweep, side configs
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2z_pytest.log | cut -c1-220
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/r2z_bench.log | cut -c1-300
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2z_ref.log 2>&1; echo "ref rc $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc $?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --hnsw-rows 0 --sweep 1,8,32,64,128,130,160,192,224,256,320,384,512,768,1024,2048 > gpurun_out/r2z_sweep.log 2>&1; echo "sweep rc $?"
for c in 1 2 4s; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_cfg$c.log 2>&1
  echo "config $c exit $?"; tail -1 gpurun_out/r2z_cfg$c.log | cut -c1-400
done
