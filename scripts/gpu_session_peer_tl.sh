#!/bin/bash
# Timelines of the peer exchange under a few diagnostic knobs: gpurun --gpus 2 -- 'bash scripts/gpu_session_peer_tl.sh r02t'
tag=${1:-tl}; mkdir -p gpurun_out
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) scripts/timeline_peer.py --peer "$@"; }
run > gpurun_out/${tag}_default.txt 2> gpurun_out/${tag}_default.err
SGR_PEER_SERIAL=1 run > gpurun_out/${tag}_serial.txt 2> gpurun_out/${tag}_serial.err
SGR_PEER_SERIAL=1 SGR_PEER_LOCAL_FACTORS=1 run > gpurun_out/${tag}_serial_local.txt 2> gpurun_out/${tag}_serial_local.err
SGR_PEER_REDUCE_BLOCKS=148 run > gpurun_out/${tag}_rb148.txt 2> gpurun_out/${tag}_rb148.err
SGR_PEER_LOCAL_FACTORS=1 run > gpurun_out/${tag}_local.txt 2> gpurun_out/${tag}_local.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29777 scripts/timeline_peer.py > gpurun_out/${tag}_nccl.txt 2> gpurun_out/${tag}_nccl.err
grep -H step_ms gpurun_out/${tag}_*.txt
