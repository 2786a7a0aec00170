cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_evaluate_files.py 192 > gpurun_out/evaluate_files.json 2> gpurun_out/evaluate_files.err; echo "rc=$?"; tail -3 gpurun_out/evaluate_files.err; cat gpurun_out/evaluate_files.json
