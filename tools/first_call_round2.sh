#!/bin/sh
# ROUND2_PLAN.md section 1 as one command (single GPU; about 3-4 minutes of box time).  Every step has its own timeout and
# log under gpurun_out/; a hang in an experimental piece cannot eat the call.
#   gpurun --timeout 420 -- 'sh tools/first_call_round2.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, timeout seconds, command...
    name=$1; t=$2; shift 2
    echo "=== $name" | tee -a gpurun_out/first_call.log
    timeout "$t" "$@" > "gpurun_out/fc_$name.log" 2>&1
    echo "rc=$? $(tail -1 "gpurun_out/fc_$name.log")" | tee -a gpurun_out/first_call.log
}
: > gpurun_out/first_call.log
run default_suite 90 python -u -m pytest -q -m gpu --timeout 60 --timeout-method=thread tests
run exp_bigru 60 env DCASE_EXPERIMENTAL=1 python -u -m pytest -q -m gpu --timeout 40 --timeout-method=thread tests/test_gpu_bigru.py
run exp_p2p_world1 40 env DCASE_EXPERIMENTAL=1 python -u -m pytest -q -m gpu --timeout 30 --timeout-method=thread tests/test_gpu_crnn.py -k p2p
run exp_pipelined 60 env DCASE_EXPERIMENTAL=1 python -u -m pytest -q -m gpu --timeout 40 --timeout-method=thread tests/test_gpu_api.py -k "pipelined or reference_train_fixture"
run gru_sweep 60 env DCASE_EXPERIMENTAL=1 python tools/gru_sweep.py --iters 50
run scaler_bench 40 python tools/scaler_bench.py
run main_synthetic 150 python examples/main_synthetic.py --clips 96 --epochs 2
run bench_default 240 python bench.py --steps 30 --warmup 6
run bench_pipelined 240 env DCASE_PIPELINE=1 python bench.py --steps 30 --warmup 6 --no-cpu-baseline
cat gpurun_out/first_call.log
