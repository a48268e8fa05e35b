#!/bin/bash
# usage: tools/ncu_summary.sh <kernel-regex> <outfile-stem> <python script...>
# full-set capture of one kernel on the GPU box; only text summaries are kept (the .ncu-rep is too large to ship)
set -e
K="$1"; OUT="$2"; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$K" -s 6 -c 1 -o /tmp/ncu_$OUT "$@" > gpurun_out/${OUT}_ncu.log 2>&1 || true
ncu -i /tmp/ncu_$OUT.ncu-rep --page details > gpurun_out/${OUT}_details.txt 2>&1 || true
ncu -i /tmp/ncu_$OUT.ncu-rep --page raw --csv > gpurun_out/${OUT}_raw.csv 2>&1 || true
ncu -i /tmp/ncu_$OUT.ncu-rep --page source --csv > gpurun_out/${OUT}_source.csv 2>&1 || true
ls -la /tmp/ncu_$OUT.ncu-rep gpurun_out/${OUT}_* || true
