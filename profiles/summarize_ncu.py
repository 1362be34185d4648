"""Markdown summary of an `ncu --set full` report: per captured launch, the metrics DESIGN.md / bench.py quote.
usage: python profiles/summarize_ncu.py gpurun_out/X.ncu-rep|X_raw.csv > profiles/rN_name.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"), ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/shared throughput %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (warp inst/clk/SM)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA inst % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU inst % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU inst % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU inst % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall mio throttle"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math pipe throttle"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not selected"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):      # a raw page exported on the GPU box (`ncu -i X.ncu-rep --page raw --csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{rep.split('/')[-1]}` (`ncu --set full --clock-control none`; cold-cache, serialised launches)\n")
    seen = {}
    for r in rows[2:]:
        name = r[col['Kernel Name']]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1 and "--all" not in sys.argv:
            continue   # one launch per kernel name (the others repeat it)
        print(f"## {r[col['Kernel Name']][:110]}\n")
        print("| metric | value |\n|---|---|")
        for k, name in KEYS:
            if k in col and r[col[k]] != "":
                print(f"| {name} (`{k}`) | {r[col[k]]} {units[col[k]]} |")
        print()
    if "--traffic" in sys.argv:      # profiles/r2_traffic.json: dram bytes per launch and kernel, read by bench.py's roofline.traffic
        import json
        import re
        out_path = sys.argv[sys.argv.index("--traffic") + 1]
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tr, order = {}, {}
        for r in rows[2:]:
            name = r[col['Kernel Name']]
            short = re.sub(r"\(.*$", "", name).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("dspfe::", "").replace("void ", "").strip()
            short = re.sub(r"^mfcc_delta_kernel<.*", "mfcc_delta_kernel", short)
            tot = sum(float(r[col[k]]) * unit[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            order[short] = order.get(short, 0) + 1
            key = short if order[short] == 1 else f"{short}#{order[short]}"      # a second launch of the same kernel in the step (the autocorrelation chain)
            tr[key] = tot
        tr["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full of profiles/run_frontend_once.py (the bench step), {rep.split('/')[-1]}"
        with open(out_path, "w") as f:
            json.dump(tr, f, indent=1)


if __name__ == "__main__":
    main()
