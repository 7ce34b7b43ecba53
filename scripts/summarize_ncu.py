"""Turn an `ncu --set full` report into the committed summaries under profiles/:
   python scripts/summarize_ncu.py gpurun_out/prof_X.ncu-rep r01
 -> profiles/r01_ncu_summary.csv  (one row per captured launch, key metrics)
 -> profiles/ncu_kernels.json     (per bench.py stage: dram bytes read+written and warp-instructions per
                                   launch, summed over the launches of one step; bench.py reads it)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
STAGE = {"preprocess_kernel": "preprocess", "tile_scan_kernel": "tile_scan", "scatter_kernel": "scatter",
         "tile_sort_merge_kernel": "tile_sort_smem", "tile_sort_bucket_kernel": "tile_sort_smem",
         "tile_sort_kernel": "tile_sort_global", "blend_forward_kernel": "blend_forward",
         "blend_backward_kernel": "blend_backward", "blend_backward_chunked_kernel": "blend_backward",
         "preprocess_backward_kernel": "preprocess_backward"}
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.csv")
    traffic, insts = {}, {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        cols = [c for c in KEEP if c in ix]
        w.writerow(["kernel", "block"] + [f"{c} [{units[ix[c]]}]" for c in cols])
        for r in rows[2:]:
            name = r[ix["Kernel Name"]]
            short = name.split("(")[0].replace("void ", "").replace("sgr::", "")
            w.writerow([short, r[ix["Block Size"]]] + [r[ix[c]] for c in cols])
            key = next((v for k, v in STAGE.items() if short.startswith(k)), None)
            if key:
                b = sum(float(r[ix[c]]) * MULT.get(units[ix[c]], 1) for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                traffic[key] = traffic.get(key, 0) + int(b)
                insts[key] = insts.get(key, 0) + int(float(r[ix["smsp__inst_executed.sum"]]))
    facts = {k: {"dram_bytes": traffic[k], "warp_instructions": insts[k], "source": f"profiles/{tag}_ncu_summary.csv"}
             for k in traffic}
    with open(os.path.join(ROOT, "profiles", "ncu_kernels.json"), "w") as f:
        json.dump(facts, f, indent=1)
    print("wrote", out, facts)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
