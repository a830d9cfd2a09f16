#!/bin/sh
# Quick GPU call: the full -m gpu suite, one bench line, the launch list of one step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
    name=$1; t=$2; shift 2
    echo "=== $name" >> gpurun_out/call.log
    timeout "$t" "$@" > "gpurun_out/c_$name.log" 2>&1
    echo "rc=$? $(tail -1 "gpurun_out/c_$name.log" | cut -c1-600)" >> gpurun_out/call.log
}
: > gpurun_out/call.log
run suite 900 python -u -m pytest -q -m gpu -s --timeout 300 --timeout-method=thread tests
run bench 300 python bench.py --steps 30 --warmup 6 --no-cpu-baseline
run ncu_list 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
cut -c1-300 gpurun_out/call.log
