#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.txt
tail -4 gpurun_out/r02_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.txt
tail -4 gpurun_out/r02_sanitizer_racecheck.txt
