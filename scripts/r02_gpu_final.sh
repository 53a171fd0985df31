#!/bin/bash
# end-of-round capture: suite, smoke, bench line, launch list, hash-stamped DRAM traffic, tail-transform ncu summary, extras
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
( time python bench.py ) > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -3 gpurun_out/r02_bench_1gpu.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -3 gpurun_out/r02_bench_ref.err
python bench.py --steps 5 --warmup 3 --realtime 0 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --realtime 0 --no-cpu-baseline > /dev/null 2>&1
bash scripts/capture_traffic.sh 2>&1 | tail -2
python scripts/bigfft_probe.py 8192 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:'k_rfft_forward|k_irfft_ola' -s 24 -c 2 -f -o /tmp/tailfft_wide python scripts/bigfft_probe.py 8192 > /dev/null 2>&1
ncu -i /tmp/tailfft_wide.ncu-rep --page raw --csv > gpurun_out/r02_tailfft_wide_raw.csv
python scripts/extra_bench.py all 320 2>&1 | tail -2 | tee gpurun_out/r02_extra.jsonl
python scripts/configs_bench.py > gpurun_out/r02_configs.jsonl 2>&1; cat gpurun_out/r02_configs.jsonl | cut -c1-200
du -sh gpurun_out
