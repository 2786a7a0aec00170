cd /root/repo
timeout 900 python -m pytest tests/test_device_metrics.py -m gpu -x -q -s 2>&1 | grep -v Warn | tail -30
timeout 600 python scripts/bench_metrics.py 100000 > gpurun_out/metrics_bench.json 2> gpurun_out/metrics_bench.err; echo "rc=$?"; tail -3 gpurun_out/metrics_bench.err; cat gpurun_out/metrics_bench.json
