# GPU box: round-2 ncu evidence (run: gpurun --timeout 2400 -- bash scripts/gpu_profile_r2.sh)
#  launch list of the DEFAULT bench command, then one --set full capture each of: K-score pairs (batch 4096, main
#  pass), K-score single CTA (batch 1), fused tail (batch 4096 and batch 1), K-pack, transposed K-score (160
#  queries), exchange push / merge (two ranks in one process)
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --hnsw-rows 0 --sweep none"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
cap() {  # name kernel-regex skip command...
  local name=$1 k=$2 s=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "capture $name exit $?"
}
# seeded searches launch (sampling pass, main pass): odd launch indices are main passes
cap score_tc2 score_tc2_kernel 3 $CMD
cap score_tc_b1 score_tc_kernel 11 $CMD
cap tail_b4096 tail_kernel 5 $CMD
cap tail_b1 tail_kernel 35 $CMD
cap pack pack_rows_kernel 10 $CMD
MID="python bench.py --rows 10000000 --batch 160 --steps 2 --warmup 1 --no-cpu-baseline --hnsw-rows 0"
$MID > gpurun_out/plain_tn_160.log 2>&1 && cap score_tn_160 score_tn_kernel 2 $MID
XT="python -m pytest tests/test_gpu_parity.py -q -m gpu -k peer_exchange"
$XT > gpurun_out/plain_xchg.log 2>&1 && { cap xchg_push xchg_push_kernel 1 $XT; cap xchg_merge xchg_merge_kernel 1 $XT; }
ls -la gpurun_out/*.ncu-rep
