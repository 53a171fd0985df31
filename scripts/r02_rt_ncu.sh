#!/bin/bash
# ncu --set full of k_mac_rt (source page with stall samples included): bash scripts/r02_rt_ncu.sh <streams> [tunes]
NS=${1:-16}; TUNE=${2:-}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mac_rt -s 6 -c 1 -f -o /tmp/rt python scripts/r02_rt_sweep.py "$NS:$TUNE" > /dev/null 2>&1
ncu -i /tmp/rt.ncu-rep --page raw --csv > gpurun_out/r02_rt${NS}_raw.csv 2>/dev/null
ncu -i /tmp/rt.ncu-rep --page source --csv > gpurun_out/r02_rt${NS}_source.csv 2>/dev/null
python scripts/ncu_summary.py < gpurun_out/r02_rt${NS}_raw.csv
ls -la gpurun_out/r02_rt${NS}_source.csv
