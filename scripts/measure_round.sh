#!/bin/bash
# How the artefacts under profiles/rN/ are produced on a B200 box (run from the repo root; about 7 minutes).
#   usage: scripts/measure_round.sh <tag> [out dir]      e.g. scripts/measure_round.sh 7620eb1 gpurun_out
# Afterwards: python scripts/profile_tables.py <out>/<tag>_launches.csv <out>/<tag>_full_raw.csv 21710 <tag> profiles/rN
TAG=${1:?tag}; OUT=${2:-gpurun_out}; mkdir -p "$OUT"
KERNELS='regex:^k_|k_stage|k_ds|k_stft|k_head|k_stem'
timeout 1500 python -m pytest tests -m gpu -q -s > "$OUT/${TAG}_gpu_tests.log" 2>&1; echo "suite rc=$?" >> "$OUT/${TAG}_gpu_tests.log"
timeout 900 python bench.py --steps 10 --warmup 3 > "$OUT/${TAG}_bench.json" 2> "$OUT/${TAG}_bench.err"
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > "$OUT/${TAG}_bench_ref.json" 2> "$OUT/${TAG}_bench_ref.err"
# A/B lines: CUDA-core stem, no stage kernel, quantising frontend, warp-specialised DS blocks, stem inside the first block
for f in 11 131 171 203 395; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-files --fusion $f > "$OUT/${TAG}_bench_f$f.json" 2> "$OUT/${TAG}_bench_f$f.err"
done
BN_DS_NC=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-files > "$OUT/${TAG}_bench_nc0.json" 2> "$OUT/${TAG}_bench_nc0.err"
timeout 600 python bench_configs.py --out "$OUT/${TAG}_configs34.json" > "$OUT/${TAG}_configs34.log" 2>&1
BN_GENERIC_TC=0 timeout 600 python bench_configs.py --out "$OUT/${TAG}_configs34_one_kernel_per_op.json" > "$OUT/${TAG}_configs34_one_kernel_per_op.log" 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" --csv --log-file "$OUT/${TAG}_launches.csv" python profiles/run_wave.py 21710 3 > "$OUT/${TAG}_launches.log" 2>&1
timeout 900 ncu --set full --clock-control none -k "$KERNELS" -o "$OUT/${TAG}_full" python profiles/run_wave.py 21710 1 > "$OUT/${TAG}_full.log" 2>&1
ncu -i "$OUT/${TAG}_full.ncu-rep" --page raw --csv > "$OUT/${TAG}_full_raw.csv" 2>/dev/null && rm -f "$OUT/${TAG}_full.ncu-rep"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/${TAG}_smoke.log" 2>&1
grep -h "passed\|failed\|rc=" "$OUT/${TAG}_gpu_tests.log" | tail -3; cut -c1-200 "$OUT/${TAG}_bench.json"; tail -1 "$OUT/${TAG}_smoke.log"
