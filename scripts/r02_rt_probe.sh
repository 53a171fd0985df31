#!/bin/bash
# k_mac_rt vs K4 by stream count (16 x 16 matrix, 10 s IRs, block 512) and one ncu --set full capture of k_mac_rt at 16 streams
mkdir -p gpurun_out
: > gpurun_out/r02_mimo_by_streams.jsonl
for spec in "2 0" "8 0" "32 0" "48 0" "64 0" "32 1" "64 1"; do
  set -- $spec
  timeout 150 python scripts/mimo_bench.py --streams $1 --tc $2 --steps 40 2>/dev/null | tail -1 >> gpurun_out/r02_mimo_by_streams.jsonl
done
cut -c1-420 gpurun_out/r02_mimo_by_streams.jsonl
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_mac_rt -s 8 -c 1 -f -o /tmp/rt16 python scripts/mimo_bench.py --streams 16 --tc 0 --steps 10 --warmup 5 > /dev/null 2>&1
ncu -i /tmp/rt16.ncu-rep --page raw --csv > gpurun_out/r02_rt16_raw.csv 2>/dev/null
python scripts/ncu_summary.py < gpurun_out/r02_rt16_raw.csv
