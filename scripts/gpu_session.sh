#!/bin/bash
# One GPU-box session: tests, blend-loop statistics, both bench arms, ncu launch list + full capture of one step.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh r02a [quick]'
# Everything lands in gpurun_out/<tag>_*; nothing here is a bench value except <tag>_bench*.json.
tag=${1:-sess}; mode=${2:-full}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/${tag}_tests.log
tail -15 gpurun_out/${tag}_tests.log
if [ -f sugar_b200/lib/variants/lib_stats2.so ]; then
  SGR_LIB_PATH=$PWD/sugar_b200/lib/variants/lib_stats2.so python scripts/blend_stats.py > gpurun_out/${tag}_stats.json 2> gpurun_out/${tag}_stats.err
fi
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("views/s %.1f  ms/step %.3f  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    print("  ".join("%s=%.3f" % (k, v["ms"]) for k, v in d["stages"].items()))
except Exception as e:
    print("bench failed:", e); print(open("gpurun_out/${tag}_bench.err").read()[-2000:])
PY
[ "$mode" = "quick" ] && exit 0
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
ls sugar_b200/lib/variants/*.so 2>/dev/null | grep -v stats > /dev/null && bash scripts/ab_bench.sh > gpurun_out/${tag}_ab.log 2>&1
for wl in coarse_sdf_step refine_step; do
  python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${tag}_${wl}.json 2> gpurun_out/${tag}_${wl}.err
  python bench.py --workload $wl --impl reference --steps 5 --warmup 3 > gpurun_out/${tag}_${wl}_ref.json 2> gpurun_out/${tag}_${wl}_ref.err
  python - <<PY
import json
for f in ("gpurun_out/${tag}_${wl}.json", "gpurun_out/${tag}_${wl}_ref.json"):
    try:
        d = json.load(open(f)); print("${wl}", d.get("impl", "ours"), "%.2f steps/s  %.2f ms" % (d["value"], d["ms_per_step"]), d.get("workload_stats"))
    except Exception as e:
        print("${wl} failed:", f, e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
done
# ncu: launch list of three steps (kernel shares), then one full-set capture of the third step's kernels
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'blend_|preprocess|scatter|tile_|finalize|meshbind' --csv --log-file gpurun_out/${tag}_launches.csv \
    python scripts/profile_step.py 3000000 1920 1080 3 > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'blend_|preprocess|scatter|tile_|finalize|meshbind' --launch-skip 20 --launch-count 10 -f \
    -o gpurun_out/${tag}_full python scripts/profile_step.py 3000000 1920 1080 3 > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out | tail -20
