mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc $?"; tail -6 gpurun_out/r2j_pytest.log | cut -c1-220
python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/r2j_bench.log | cut -c1-300
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2j_ref.log 2>&1; echo "ref rc $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2j_smoke.log | cut -c1-300
