cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ingest_launches.csv python scripts/bench_ingest.py 4 > gpurun_out/ib.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ingest_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:16]: print(r[4][:50], r[-1])
PY
