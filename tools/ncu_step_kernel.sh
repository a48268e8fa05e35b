#!/bin/bash
# usage: tools/ncu_step_kernel.sh <kernel-regex> <stem> [skip]   -- full-set capture of ONE launch inside a step
K="$1"; OUT="$2"; SKIP="${3:-0}"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$K" -s $SKIP -c 1 \
    -o /tmp/ncu_$OUT python tools/profile_step.py ncu > gpurun_out/${OUT}_ncu.log 2>&1 || true
ncu -i /tmp/ncu_$OUT.ncu-rep --page details > gpurun_out/${OUT}_details.txt 2>&1 || true
ncu -i /tmp/ncu_$OUT.ncu-rep --page source --csv > gpurun_out/${OUT}_source.csv 2>&1 || true
