cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_ -c 20 -o gpurun_out/full_wave2368 -f python profiles/run_wave.py 2368 1 > gpurun_out/rw_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/full_wave2368.ncu-rep
