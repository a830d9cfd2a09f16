#!/bin/sh
# STFT-only GPU call: log-mel parity tests, the mel sweep, one bench line and an ncu --set full capture of the kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
    name=$1; t=$2; shift 2
    echo "=== $name" >> gpurun_out/call.log
    timeout "$t" "$@" > "gpurun_out/c_$name.log" 2>&1
    echo "rc=$? $(tail -1 "gpurun_out/c_$name.log" | cut -c1-600)" >> gpurun_out/call.log
}
: > gpurun_out/call.log
run logmel_tests 300 python -u -m pytest -q -m gpu -s --timeout 200 --timeout-method=thread tests/test_gpu_logmel.py tests/test_gpu_scaler.py tests/test_gpu_api.py
run mel_sweep 200 python tools/mel_sweep.py
run bench 300 python bench.py --steps 30 --warmup 6 --no-cpu-baseline
run ncu_full 400 ncu --set full --clock-control none --import-source on -k "regex:stft_mel" -c 1 -o gpurun_out/ncu_stft -f python tools/profile_step.py --steps 1
cut -c1-400 gpurun_out/call.log
