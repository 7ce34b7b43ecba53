mkdir -p gpurun_out; tag=r02x
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 300 python -m pytest tests/test_gpu_parallel.py -m gpu -q -x 2>&1 | tail -3
B="--steps 20 --warmup 5 --no-cpu-baseline"
run 2 scripts/check_view_parallel.py > gpurun_out/${tag}_check_n2.json 2> gpurun_out/${tag}_check_n2.err; tail -1 gpurun_out/${tag}_check_n2.json | cut -c1-300
run 2 bench.py --gpus 2 $B --exchange nccl > gpurun_out/${tag}_bench_n2_nccl.json 2> gpurun_out/${tag}_bench_n2_nccl.err
run 2 bench.py --gpus 2 $B --exchange nccl --no-taper > gpurun_out/${tag}_bench_n2_nccl_notaper.json 2> gpurun_out/${tag}_bench_n2_nccl_notaper.err
run 2 bench.py --gpus 2 $B --exchange nccl --chunks 5 > gpurun_out/${tag}_bench_n2_nccl_chunks5.json 2> gpurun_out/${tag}_bench_n2_nccl_chunks5.err
run 2 bench.py --gpus 2 $B --exchange peer > gpurun_out/${tag}_bench_n2_peer.json 2> gpurun_out/${tag}_bench_n2_peer.err
run 2 bench.py --gpus 2 $B --exchange peer --chunks 3 > gpurun_out/${tag}_bench_n2_peer_chunks3.json 2> gpurun_out/${tag}_bench_n2_peer_chunks3.err
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n=%d  %.1f views/s  %.3f ms  e2e %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]),
              {k: round(v["ms"] * v["launches_per_step"], 3) for k, v in d.get("stages", {}).items() if k in ("preprocess_backward", "view_finalize", "peer_reduce", "peer_sync")})
    except Exception as e:
        print(f, "unreadable:", e); print(open(f.replace(".json", ".err")).read()[-800:])
PY
