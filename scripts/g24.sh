cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_stft_mag|k_stem_sat|k_head_tc' -c 3 -o gpurun_out/front python profiles/run_wave.py 2368 1 > gpurun_out/rw3.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/front.ncu-rep
