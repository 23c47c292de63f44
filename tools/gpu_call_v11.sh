#!/bin/bash
# Round-2 call V11 (1 GPU): compute-sanitizer memcheck of the final tree (the zeroing kernel + chained setup pass came after V6).
set -u
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/v11_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 2 gpurun_out/v11_memcheck.log
