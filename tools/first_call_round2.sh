#!/bin/sh
# Round-2 first GPU call (single GPU): the default suite incl. the new full-size training parity tests, then every piece
# that was written after round 1's GPU budget was spent (gated by DCASE_EXPERIMENTAL=1), each with its own timeout and log
# under gpurun_out/ so that a hang in one piece cannot eat the call.
#   gpurun --timeout 1200 -- 'sh tools/first_call_round2.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, timeout seconds, command...
    name=$1; t=$2; shift 2
    echo "=== $name" | tee -a gpurun_out/first_call.log
    timeout "$t" "$@" > "gpurun_out/fc_$name.log" 2>&1
    echo "rc=$? $(tail -1 "gpurun_out/fc_$name.log")" | tee -a gpurun_out/first_call.log
}
: > gpurun_out/first_call.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/fc_smi.log 2>&1
run default_suite 420 python -u -m pytest -q -m gpu -s --timeout 300 --timeout-method=thread tests --deselect tests/test_gpu_fullsize.py
run fullsize 420 python -u -m pytest -q -m gpu -s --timeout 300 --timeout-method=thread tests/test_gpu_fullsize.py
run exp_bigru 60 env DCASE_EXPERIMENTAL=1 python -u -m pytest -q -m gpu --timeout 40 --timeout-method=thread tests/test_gpu_bigru.py
run exp_p2p_world1 60 env DCASE_EXPERIMENTAL=1 python -u -m pytest -q -m gpu --timeout 30 --timeout-method=thread tests/test_gpu_crnn.py -k p2p
run exp_pipelined 90 env DCASE_EXPERIMENTAL=1 python -u -m pytest -q -m gpu --timeout 60 --timeout-method=thread tests/test_gpu_api.py -k "pipelined"
run gru_sweep 60 env DCASE_EXPERIMENTAL=1 python tools/gru_sweep.py --iters 50
run main_synthetic 150 python examples/main_synthetic.py --clips 96 --epochs 2
run bench_default 300 python bench.py --steps 30 --warmup 6
run bench_pipelined 240 env DCASE_PIPELINE=1 python bench.py --steps 30 --warmup 6 --no-cpu-baseline
cat gpurun_out/first_call.log
