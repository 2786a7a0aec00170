cd /root/repo
BN_DEBUG=1 python profiles/run_wave.py 64 1 2>&1 | tail -20
bash scripts/g3.sh
