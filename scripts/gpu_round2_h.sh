mkdir -p gpurun_out
python scripts/exp_power.py > gpurun_out/r2h_power.log 2>&1; cut -c1-300 gpurun_out/r2h_power.log
python scripts/exp_pack.py > gpurun_out/r2h_pack.log 2>&1; B2K_PACK_NO_BULK=1 python scripts/exp_pack.py >> gpurun_out/r2h_pack.log 2>&1; cut -c1-230 gpurun_out/r2h_pack.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pack" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2h_pytest.log | cut -c1-200
