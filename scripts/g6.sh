set -x
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 30 -c 30 -o /tmp/r1_full python profiles/run_wave.py 2048 2 > gpurun_out/r1_ncu_full.log 2>&1; tail -2 gpurun_out/r1_ncu_full.log
ncu -i /tmp/r1_full.ncu-rep --page raw --csv > gpurun_out/r1_full_raw.csv
ncu -i /tmp/r1_full.ncu-rep --page details --csv > gpurun_out/r1_full_details.csv
for k in k_stft_mag k_head k_stem k_tail; do
  ncu -i /tmp/r1_full.ncu-rep --page source --print-source sass --csv --kernel-name regex:$k > gpurun_out/r1_src_$k.csv
done
ncu -i /tmp/r1_full.ncu-rep --page source --print-source sass --csv --kernel-name regex:k_ds --launch-skip 1 --launch-count 1 > gpurun_out/r1_src_k_ds_01.csv
ncu -i /tmp/r1_full.ncu-rep --page source --print-source sass --csv --kernel-name regex:k_ds --launch-skip 0 --launch-count 1 > gpurun_out/r1_src_k_ds_00.csv
ncu -i /tmp/r1_full.ncu-rep --page source --print-source sass --csv --kernel-name regex:k_ds --launch-skip 6 --launch-count 1 > gpurun_out/r1_src_k_ds_06.csv
gzip -f gpurun_out/r1_src_*.csv
ls -la gpurun_out /tmp/r1_full.ncu-rep
