#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 14 gpurun_out/$name.log | cut -c1-400; }
export PYTHONPATH=$PWD
run overhead python scripts/cpu_overhead.py
