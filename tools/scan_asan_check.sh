#!/bin/bash
# The lane code of the scan solvers (csrc/ibs_scan_core.cuh, csrc/ibs_scan2_core.cuh) under AddressSanitizer + UBSan through the CPU harness:
# every row write and record read (including the prefetches) of the per-lane arithmetic the CUDA kernel runs.
# (compute-sanitizer is not available on the GPU pool.)   usage: tools/scan_asan_check.sh
set -e
cd "$(dirname "$0")/.."
out=${SCAN_ASAN_DIR:-/tmp/sch_asan}
mkdir -p "$out"
g++ -O1 -g -std=c++17 -shared -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-unknown-pragmas \
    -o "$out/scan_core_host.so" tools/scan2_core_host.cpp        # (includes scan_core_host.cpp: both kernels' lane code)
SCAN_ASAN_SO="$out/scan_core_host.so" LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
    python tools/scan_asan_run.py
