#!/bin/sh
# Two-GPU call (gpurun --gpus 2): the 2-rank parity test (NCCL eager / NCCL in the CUDA graph / fused p2p exchange), then
# bench.py at N = 2 in each mode.  Every step under its own timeout: a missed flag in the p2p protocol is a hang.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
    name=$1; t=$2; shift 2
    echo "=== $name" | tee -a gpurun_out/dp_call.log
    timeout "$t" "$@" > "gpurun_out/dp_$name.log" 2>&1
    echo "rc=$? $(tail -1 "gpurun_out/dp_$name.log" | cut -c1-1500)" | tee -a gpurun_out/dp_call.log
}
: > gpurun_out/dp_call.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run test_eager 200 python -u -m pytest -q -m gpu -s --timeout 180 --timeout-method=thread tests/test_gpu_dp.py -k "bit_identical"
run test_graph 200 python -u -m pytest -q -m gpu -s --timeout 180 --timeout-method=thread tests/test_gpu_dp.py -k "cuda_graph"
run test_p2p 200 python -u -m pytest -q -m gpu -s --timeout 180 --timeout-method=thread tests/test_gpu_dp.py -k "p2p_exchange_and"
run test_p2p_graph 200 python -u -m pytest -q -m gpu -s --timeout 180 --timeout-method=thread tests/test_gpu_dp.py -k "p2p_exchange_inside"
run bench_eager 200 $TR bench.py --gpus 2 --steps 30 --warmup 6
run bench_graph 200 env DCASE_DP_GRAPH=1 $TR bench.py --gpus 2 --steps 30 --warmup 6
run bench_p2p 200 env DCASE_DP_P2P=1 $TR bench.py --gpus 2 --steps 30 --warmup 6
run bench_p2p_graph 200 env DCASE_DP_P2P=1 DCASE_DP_GRAPH=1 $TR bench.py --gpus 2 --steps 30 --warmup 6
cat gpurun_out/dp_call.log
