mkdir -p gpurun_out
python scripts/exp_tail.py > gpurun_out/tail_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:score_tc|select_|seed_|rerank_|finalize_|exact_|query_prep|collect_' -c 400 --csv --log-file gpurun_out/tail_launches.csv python scripts/exp_tail.py > gpurun_out/tail_ncu.log 2>&1
echo "exit $?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/tail_launches.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[h]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); gi=H.index('Grid Size')
for r in rows[h+1:]:
    if len(r)>vi:
        try: print(r[ki][:28], r[gi], float(r[vi].replace(',',''))/1e3)
        except ValueError: pass
PY
