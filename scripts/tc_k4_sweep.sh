#!/bin/bash
# K4 harness over ring positions, odd/even segment counts, shards and input groups.  Build the harness first (no GPU needed):
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Ifft_convolution_b200/csrc -Iinclude \
#        -o build/tc_k4_test scripts/tc_k4_test.cu
# perf mode: TC_PERF=1 build/tc_k4_test 512 16 938 <streams> 7 2
fail=0
for args in "4 2 32 5 7" "4 2 38 5 7" "3 2 37 5 0" "3 2 37 5 1" "3 2 37 5 36" "3 2 37 5 20" "2 3 5 2 3" "2 1 1 1 0" \
            "2 3 40 7 13 2" "2 3 41 7 13 3 11 29" "2 3 41 7 30 1 11 29" "2 3 41 7 31 1 12 29" "2 2 41 7 40 1 0 13" "2 2 64 128 9 2"; do
    out=$(timeout 60 build/tc_k4_test $args 2>&1 | tail -1)
    echo "$args -> $out"
    case "$out" in *"max err / rms = "[0-9].[0-9]*e-0[6-9]*) ;; *) fail=1 ;; esac
done
exit $fail
