cd /root/repo
BN_DS_CTAS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ds -s 1 -c 1 -o gpurun_out/ds01 python profiles/run_wave.py 2368 1 > gpurun_out/rw2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ds01.ncu-rep
