#!/bin/bash
# Exchange-path session:  gpurun [--gpus 2] --timeout 1200 -- 'bash scripts/gpu_session_trace.sh r02t [N]'
# Forced-exchange tests, bench with the exchange path's kernels on one rank (no NCCL), and with N=2: the exchange check,
# bench lines with / without the side-stream finalize, torch.profiler timelines (scripts/trace_step.py).
tag=${1:-trace}; N=${2:-1}
mkdir -p gpurun_out
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) "$@"; }
line() { python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n=%d  %.1f views/s  %.3f ms  e2e %.1f (%.3f ms)" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
        print("   ", "  ".join("%s=%.3f" % (k, v["ms"]) for k, v in d.get("stages", {}).items()))
    except Exception as e:
        print(f, "unreadable:", e)
        try: print(open(f.replace(".json", ".err")).read()[-1500:])
        except Exception: pass
PY
}
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -q 2>&1 | tail -5 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline"
timeout 300 python bench.py $B > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 300 python bench.py $B --force-exchange > gpurun_out/${tag}_bench_n1_forced.json 2> gpurun_out/${tag}_bench_n1_forced.err
timeout 300 python bench.py $B --force-exchange --side-stream > gpurun_out/${tag}_bench_n1_forced_side.json 2> gpurun_out/${tag}_bench_n1_forced_side.err
timeout 300 python bench.py $B --force-exchange --side-stream --chunks 8 > gpurun_out/${tag}_bench_n1_forced_side_c8.json 2> gpurun_out/${tag}_bench_n1_forced_side_c8.err
line gpurun_out/${tag}_bench_n1.json gpurun_out/${tag}_bench_n1_forced.json gpurun_out/${tag}_bench_n1_forced_side.json gpurun_out/${tag}_bench_n1_forced_side_c8.json
if [ "$N" -ge 2 ]; then
  run 2 scripts/check_view_parallel.py > gpurun_out/${tag}_check_n2.json 2> gpurun_out/${tag}_check_n2.err; tail -1 gpurun_out/${tag}_check_n2.json; tail -2 gpurun_out/${tag}_check_n2.err
  run 2 bench.py --gpus 2 $B > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
  run 2 bench.py --gpus 2 $B --side-stream > gpurun_out/${tag}_bench_n2_side.json 2> gpurun_out/${tag}_bench_n2_side.err
  run 2 bench.py --gpus 2 $B --side-stream --chunks 8 > gpurun_out/${tag}_bench_n2_side_c8.json 2> gpurun_out/${tag}_bench_n2_side_c8.err
  run 2 bench.py --gpus 2 $B --side-stream --chunks 2 > gpurun_out/${tag}_bench_n2_side_c2.json 2> gpurun_out/${tag}_bench_n2_side_c2.err
  line gpurun_out/${tag}_bench_n2.json gpurun_out/${tag}_bench_n2_side.json gpurun_out/${tag}_bench_n2_side_c8.json gpurun_out/${tag}_bench_n2_side_c2.json
  run 2 scripts/trace_step.py --chunks 4 --out gpurun_out/${tag}_trace_c4 > gpurun_out/${tag}_c4_n2.txt 2> gpurun_out/${tag}_c4_n2.err
  run 2 scripts/trace_step.py --chunks 4 --side-stream --out gpurun_out/${tag}_trace_c4_side > gpurun_out/${tag}_c4_side_n2.txt 2> gpurun_out/${tag}_c4_side_n2.err
  tail -2 gpurun_out/${tag}_c4_n2.err
else
  timeout 300 python scripts/trace_step.py --force-exchange --side-stream --out gpurun_out/${tag}_trace_forced_side > gpurun_out/${tag}_forced_side_n1.txt 2> gpurun_out/${tag}_forced_side_n1.err
fi
gzip -f gpurun_out/${tag}_trace*.json 2>/dev/null
ls gpurun_out | grep ${tag} | head -40
