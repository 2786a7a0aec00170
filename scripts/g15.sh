cd /root/repo
timeout 900 python -m pytest tests/test_ingest.py tests/test_evaluate.py tests/test_gpu_parity.py -m gpu -x -q -s 2>&1 | grep -v Warning | tail -12
timeout 300 python scripts/bench_ingest.py 32 > gpurun_out/ingest_bench.json 2> gpurun_out/ingest_bench.err; echo "rc=$?"; tail -3 gpurun_out/ingest_bench.err; cat gpurun_out/ingest_bench.json
