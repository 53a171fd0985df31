#!/bin/bash
# end-of-round capture, second half of round 2 (after k_mac_rt; the last capture of the round ran its first six steps): suite, smoke, bench line, reference arm, launch list,
# hash-stamped DRAM traffic, ncu --set full of k_mac_rt at 16 and 32 streams, matrix sweep by stream count.
# Every step under its own timeout; outputs under gpurun_out/ (copy what is judged into profiles/).
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_gpu_suite.txt 2>&1; tail -4 gpurun_out/r02_gpu_suite.txt
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r02_smoke.txt
( time timeout 400 python bench.py ) > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -4 gpurun_out/r02_bench_1gpu.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -4 gpurun_out/r02_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --realtime 0 --no-cpu-baseline > /dev/null 2>&1
timeout 400 bash scripts/capture_traffic.sh 2>&1 | tail -2
STEPS=40 timeout 300 python scripts/r02_rt_sweep.py 1: 2: 4: 8: 16: 24: 32: 64:mimo_tc=1 128:mimo_tc=1 > gpurun_out/r02_mimo_by_streams.jsonl 2> gpurun_out/r02_mimo_by_streams.err; cat gpurun_out/r02_mimo_by_streams.jsonl
for NS in 16 32; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mac_rt -s 6 -c 1 -f -o /tmp/rt$NS python scripts/r02_rt_sweep.py "$NS:" > /dev/null 2>&1
  ncu -i /tmp/rt$NS.ncu-rep --page raw --csv > /tmp/rt${NS}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py < /tmp/rt${NS}_raw.csv > gpurun_out/r02_rt${NS}_summary.txt 2>&1
  python scripts/ncu_stalls.py < /tmp/rt${NS}_raw.csv >> gpurun_out/r02_rt${NS}_summary.txt 2>&1
done
du -sh gpurun_out
