#!/bin/bash
set -x
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -3 gpurun_out/r02_bench_1gpu.err
python bench.py --steps 5 --warmup 3 --realtime 0 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --realtime 0 --no-cpu-baseline > /dev/null 2>&1
bash scripts/capture_traffic.sh 2>&1 | tail -2
python scripts/bigfft_probe.py 8192 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:'k_rfft_forward|k_irfft_ola' -s 24 -c 2 -f -o /tmp/tailfft_wide python scripts/bigfft_probe.py 8192 > /dev/null 2>&1
ncu -i /tmp/tailfft_wide.ncu-rep --page raw --csv > gpurun_out/r02_tailfft_wide_raw.csv
python scripts/extra_bench.py all 320 2>&1 | tail -2 | tee gpurun_out/r02_extra.jsonl
python scripts/offline_bench.py 2>&1 | tee gpurun_out/r02_offline.jsonl | tail -3
du -sh gpurun_out
