tag=r02m8; mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/${tag}_gpus.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
for n in 2 4 8; do run $n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err; done
run 8 bench.py --gpus 8 --steps 20 --warmup 5 --chunks 2 --no-cpu-baseline > gpurun_out/${tag}_bench_n8_chunks2.json 2> gpurun_out/${tag}_bench_n8_chunks2.err
run 8 scripts/check_view_parallel.py > gpurun_out/${tag}_check_n8.json 2> gpurun_out/${tag}_check_n8.err
python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_c5_n1.json 2> gpurun_out/${tag}_c5_n1.err
run 8 bench.py --workload c5 --gpus 8 --steps 5 --warmup 3 > gpurun_out/${tag}_c5_n8.json 2> gpurun_out/${tag}_c5_n8.err
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if "value" in d:
            print(f, "n=%d  %.1f views/s  %.3f ms  e2e %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("parallelism", {}).get("exchange_check"))
        else:
            print(f, str(d)[:300])
    except Exception as e:
        print(f, "unreadable:", e)
        try: print(open(f.replace(".json", ".err")).read()[-1500:])
        except Exception: pass
PY
