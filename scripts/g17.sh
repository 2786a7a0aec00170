cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_phase -c 1 -o gpurun_out/resample_phase python scripts/bench_ingest.py 2 > gpurun_out/ib2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
