#!/bin/bash
# per-call latency at the C ABI (tests/abi/latency.c) for configs[0..2]: defaults, and the crossfade gains on the critical path
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_crossfade_speculation.py tests/test_gpu_random_sequences.py tests/test_gpu_realtime.py tests/test_gpu_parity.py tests/test_gpu_reference_tests.py -q -x -m gpu 2>&1 | tail -4
/usr/bin/gcc -O2 -Iinclude tests/abi/latency.c -o /tmp/latency -Lfft_convolution_b200 -lfftconv_b200 -Wl,-rpath,$PWD/fft_convolution_b200 -lm || exit 1
: > gpurun_out/r02_abi_latency2.jsonl
for c in 0 1 2; do
  for mode in "pinned" "" ; do
    for tune in "" ; do
      echo "# config $c $mode $tune" >> gpurun_out/r02_abi_latency2.jsonl
      timeout 120 /tmp/latency $c 3000 $mode $tune >> gpurun_out/r02_abi_latency2.jsonl 2>&1
    done
  done
done
for mode in "pinned" "" ; do
  echo "# config 2 $mode xf_speculate=0" >> gpurun_out/r02_abi_latency2.jsonl
  timeout 120 /tmp/latency 2 3000 $mode xf_speculate=0 >> gpurun_out/r02_abi_latency2.jsonl 2>&1
done
cut -c1-170 gpurun_out/r02_abi_latency2.jsonl
