#!/bin/bash
# Peer-memory exchange session on ONE box with N >= 2 GPUs:
#   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_session_peer.sh r02p 2'
# single-rank peer tests, the multi-rank exchange check (peer + NCCL), then bench lines: N=1, N ranks with the peer
# exchange (default), N ranks over NCCL, other chunk layouts, and a kernel timeline.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-peer}; N=${2:-2}
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/${tag}_gpus.txt
timeout 300 python -m pytest tests/test_gpu_parallel.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -4 gpurun_out/${tag}_tests.log
run $N scripts/check_view_parallel.py > gpurun_out/${tag}_check_n${N}.json 2> gpurun_out/${tag}_check_n${N}.err
tail -1 gpurun_out/${tag}_check_n${N}.json; tail -5 gpurun_out/${tag}_check_n${N}.err
B="--steps 20 --warmup 5 --no-cpu-baseline"
timeout 300 python bench.py $B > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
run $N bench.py --gpus $N $B --exchange peer > gpurun_out/${tag}_bench_n${N}_peer.json 2> gpurun_out/${tag}_bench_n${N}_peer.err
run $N bench.py --gpus $N $B --exchange nccl > gpurun_out/${tag}_bench_n${N}_nccl.json 2> gpurun_out/${tag}_bench_n${N}_nccl.err
grep -q '"value"' gpurun_out/${tag}_bench_n${N}_peer.json || { echo "peer bench failed; stopping"; tail -5 gpurun_out/${tag}_bench_n${N}_peer.err; head -c 300 gpurun_out/${tag}_bench_n${N}_peer.json; exit 1; }
if [ "${3:-}" != "quick" ]; then
run $N bench.py --gpus $N $B --exchange peer --chunks 3 > gpurun_out/${tag}_bench_n${N}_peer_chunks3.json 2> gpurun_out/${tag}_bench_n${N}_peer_chunks3.err
run $N bench.py --gpus $N $B --exchange peer --chunks 5 > gpurun_out/${tag}_bench_n${N}_peer_chunks5.json 2> gpurun_out/${tag}_bench_n${N}_peer_chunks5.err
run $N bench.py --gpus $N $B --exchange peer --no-taper > gpurun_out/${tag}_bench_n${N}_peer_notaper.json 2> gpurun_out/${tag}_bench_n${N}_peer_notaper.err
fi
run $N scripts/timeline_peer.py --peer > gpurun_out/${tag}_timeline_n${N}.txt 2> gpurun_out/${tag}_timeline_n${N}.err
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if "value" in d:
            print(f, "n=%d  %.1f views/s  %.3f ms  e2e %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]),
                  (d.get("parallelism", {}).get("exchange_check") or {}).get("peer_memory"),
                  {k: round(v["ms"] * v["launches_per_step"], 3) for k, v in d.get("stages", {}).items() if k in ("preprocess_backward", "view_finalize", "peer_reduce", "peer_sync", "blend_backward")})
        else:
            print(f, str(d)[:400])
    except Exception as e:
        print(f, "unreadable:", e)
        try: print(open(f.replace(".json", ".err")).read()[-1500:])
        except Exception: pass
PY
