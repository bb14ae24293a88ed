#!/bin/bash
# each case in its own process
cd "$(dirname "$0")"
run() { echo -n "nx=$1 ny=$2 nz=$3 box=$4x$5 at ($6,$7,$8) via=$9: "; timeout 30 ./tma_probe "$@" 2>&1 | tail -1; }
run 512 512 8 84 48 132 77 3 0
run 512 512 8 84 48 131 76 3 0
run 512 512 8 84 48 130 76 3 0
run 512 512 8 84 48 128 75 3 0
run 84 48 8 84 48 0 0 3 0
run 84 48 8 84 48 -8 -8 3 0
run 84 48 8 84 48 20 30 7 0
