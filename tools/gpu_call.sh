#!/bin/sh
# Generic GPU call: the full -m gpu suite, the two sweeps, bench.py, and (optionally) ncu captures.  Short per-step
# summaries go to gpurun_out/call.log; full logs next to it.
#   gpurun --timeout 1500 -- 'sh tools/gpu_call.sh [ncu-kernel-regex]'
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
run mel_sweep 200 python tools/mel_sweep.py
run gru_sweep 120 python tools/gru_sweep.py --iters 50
run bench 300 python bench.py --steps 30 --warmup 6
if [ -n "$1" ]; then
    run ncu_full 400 ncu --set full --clock-control none --import-source on -k "regex:$1" -c 3 -o gpurun_out/ncu_full -f python tools/profile_step.py
fi
run ncu_list 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
cat gpurun_out/call.log | cut -c1-700
