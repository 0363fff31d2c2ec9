mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2b_pytest.log
python scripts/exp_pack.py > gpurun_out/r2b_pack.log 2>&1; B2K_PACK_TWO_PASS=1 python scripts/exp_pack.py >> gpurun_out/r2b_pack.log 2>&1; cat gpurun_out/r2b_pack.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pack_rows_reg_kernel -s 20 -c 1 -o gpurun_out/prof_pack_reg python scripts/exp_pack.py > gpurun_out/ncu_pack_reg.log 2>&1; echo "ncu rc $?"
