#!/bin/bash
# Re-captures the DRAM traffic of the dominant kernel of the default bench (ncu --set full) and stamps
# profiles/fused_traffic.json with the hash of the kernel sources it was taken from; bench.py only quotes
# `roofline.traffic` while that hash matches the sources being timed.  Run under gpurun as the last act of a round:
#   gpurun -- 'bash scripts/capture_traffic.sh'      (then copy gpurun_out/fused_traffic.json to profiles/)
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --realtime 0 --no-cpu-baseline"
$CMD > gpurun_out/traffic_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_block_fused -s 100 -c 2 -f -o gpurun_out/r02_fused $CMD > gpurun_out/traffic_ncu.log 2>&1
ncu -i gpurun_out/r02_fused.ncu-rep --page raw --csv > gpurun_out/r02_fused_raw.csv
python - <<'PY'
import csv, json, sys
sys.path.insert(0, '.')
import bench
rows = list(csv.reader(open('gpurun_out/r02_fused_raw.csv')))
hdr, units = rows[0], rows[1]
def col(r, name):
    v, u = float(r[hdr.index(name)]), units[hdr.index(name)]
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
tot = [col(r, 'dram__bytes_read.sum') + col(r, 'dram__bytes_write.sum') for r in rows[2:]]
out = {"kernel": rows[2][hdr.index('Kernel Name')], "channels": 4096, "dram_bytes_per_launch": sum(tot) / len(tot),
       "kernel_source_sha256_16": bench.kernel_source_hash(),
       "duration_us": [float(r[hdr.index('gpu__time_duration.sum')]) for r in rows[2:]],
       "dram_throughput_pct": [float(r[hdr.index('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')]) for r in rows[2:]],
       "source": "scripts/capture_traffic.sh: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of %d launches" % len(tot)}
json.dump(out, open('gpurun_out/fused_traffic.json', 'w'), indent=1)
print(json.dumps(out))
PY
rm -f gpurun_out/r02_fused.ncu-rep gpurun_out/r02_fused_raw.csv  # gpurun brings back at most 64 MiB
