#!/bin/bash
# 8 GPUs, one box, short: the peer-memory exchange at 8 ranks (its own exchange check runs inside bench.py), then NCCL.
#   gpurun --gpus 8 --timeout 300 -- 'bash scripts/gpu_session_n8.sh r02n8'
tag=${1:-n8}; mkdir -p gpurun_out
run() { n=$1; shift; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
B="--steps 10 --warmup 3 --no-cpu-baseline"
run 8 bench.py --gpus 8 $B --exchange peer > gpurun_out/${tag}_bench_n8_peer.json 2> gpurun_out/${tag}_bench_n8_peer.err
tail -c 1500 gpurun_out/${tag}_bench_n8_peer.json | head -c 400; echo
run 8 bench.py --gpus 8 $B --exchange nccl > gpurun_out/${tag}_bench_n8_nccl.json 2> gpurun_out/${tag}_bench_n8_nccl.err
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n=%d  %.1f views/s  %.3f ms  e2e %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]), d["parallelism"]["exchange_check"],
              {k: round(v["ms"] * v["launches_per_step"], 3) for k, v in d.get("stages", {}).items() if k in ("preprocess_backward", "view_finalize", "peer_reduce", "peer_sync")})
    except Exception as e:
        print(f, "unreadable:", e); print(open(f.replace(".json", ".err")).read()[-800:])
PY
