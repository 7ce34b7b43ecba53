#!/bin/bash
# Multi-GPU session on ONE box:  gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_session_multi.sh r02m N'
# exchange check (both SH modes, several chunk counts), the driver's scaling line at 2..N ranks, and (N = 8) BASELINE
# config 5 as a strong-scaling batch.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-multi}; N=${2:-2}
mkdir -p gpurun_out
run() {  # ranks, script + args -> stdout
  n=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"
}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/${tag}_gpus.txt
run 2 scripts/check_view_parallel.py > gpurun_out/${tag}_check_n2.json 2> gpurun_out/${tag}_check_n2.err
tail -1 gpurun_out/${tag}_check_n2.json; tail -3 gpurun_out/${tag}_check_n2.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
for n in 2 4 8; do
  [ $n -le $N ] || continue
  run $n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
done
run $N bench.py --gpus $N --steps 20 --warmup 5 --chunks 2 --no-cpu-baseline > gpurun_out/${tag}_bench_n${N}_chunks2.json 2> gpurun_out/${tag}_bench_n${N}_chunks2.err
run $N bench.py --gpus $N --steps 20 --warmup 5 --chunks 8 --no-cpu-baseline > gpurun_out/${tag}_bench_n${N}_chunks8.json 2> gpurun_out/${tag}_bench_n${N}_chunks8.err
if [ $N -ge 8 ]; then
  run 8 scripts/check_view_parallel.py > gpurun_out/${tag}_check_n8.json 2> gpurun_out/${tag}_check_n8.err
  run 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-sh-factors > gpurun_out/${tag}_bench_n8_plain.json 2> gpurun_out/${tag}_bench_n8_plain.err
  python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_c5_n1.json 2> gpurun_out/${tag}_c5_n1.err
  for n in 2 4 8; do
    run $n bench.py --workload c5 --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_c5_n${n}.json 2> gpurun_out/${tag}_c5_n${n}.err
  done
fi
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if "value" in d:
            print(f, "n=%d  %.1f views/s  %.3f ms  e2e %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]),
                  d.get("parallelism", {}).get("exchange_check"))
        else:
            print(f, d)
    except Exception as e:
        print(f, "unreadable:", e)
        try: print(open(f.replace(".json", ".err")).read()[-1200:])
        except Exception: pass
PY
