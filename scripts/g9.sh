cd /root/repo
timeout 1200 python -m pytest tests/test_ptq.py -m gpu -x -q 2>&1 | tail -25
