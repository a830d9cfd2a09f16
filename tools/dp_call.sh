#!/bin/sh
# Multi-GPU call (gpurun --gpus N): the 2-rank parity tests, then bench.py at the requested world sizes.
#   gpurun --gpus 2 -- 'sh tools/dp_call.sh 2'        gpurun --gpus 8 -- 'sh tools/dp_call.sh 4 8'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
    name=$1; t=$2; shift 2
    echo "=== $name" >> gpurun_out/dp_call.log
    timeout "$t" "$@" > "gpurun_out/dp_$name.log" 2>&1
    echo "rc=$? $(tail -1 "gpurun_out/dp_$name.log" | cut -c1-900)" >> gpurun_out/dp_call.log
}
: > gpurun_out/dp_call.log
run tests 600 python -u -m pytest -q -m gpu -s --timeout 280 --timeout-method=thread tests/test_gpu_dp.py tests/test_gpu_syncbn.py tests/test_gpu_crnn.py tests/test_gpu_api.py
for n in "$@"; do
    run "bench_n$n" 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$n" --steps 30 --warmup 6
done
if [ -n "$SYNCBN_ALSO" ]; then
    for n in "$@"; do
        run "bench_syncbn_n$n" 300 env DCASE_SYNC_BN=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus "$n" --steps 30 --warmup 6
    done
fi
if [ -n "$NCCL_ALSO" ]; then
    for n in "$@"; do
        run "bench_nccl_n$n" 300 env DCASE_DP_NCCL=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus "$n" --steps 30 --warmup 6
    done
fi
cut -c1-400 gpurun_out/dp_call.log
