#!/bin/sh
# End-of-round evidence call: -m gpu suite, the two sweeps, bench.py (with the CPU baseline), an `ncu --set full` capture of
# every kernel of one step (-> profiles/r2_ncu_summary_final.csv, profiles/ncu_traffic.json), the launch list, the timeline.
#   gpurun --timeout 1500 -- 'sh tools/final_call.sh'
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
run bench 400 python bench.py --steps 50 --warmup 6
run bench_ref 400 python bench.py --impl reference --steps 2 --warmup 1
run ncu_full 600 ncu --set full --clock-control none -c 70 -o /tmp/ncu_step -f python tools/profile_step.py --steps 1
# the report of a whole step exceeds what gpurun copies back: reduce it here
python tools/ncu_summary.py /tmp/ncu_step.ncu-rep > gpurun_out/ncu_summary_step.csv 2> gpurun_out/c_ncu_summary.log
run ncu_list 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
run timeline 120 python tools/timeline.py --pipelined
cut -c1-300 gpurun_out/call.log
