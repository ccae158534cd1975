"""Summarise an .ncu-rep here (no GPU needed): key raw metrics + source lines ranked by stall samples.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--top 40] [--json out.json]"""
import csv, io, json, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio"]

def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout

def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr = raw[0]
    out = {}
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        d = {}
        for h, u, v in zip(hdr, raw[1], row):
            hh = h.split("TriageCompute.")[-1]
            if hh in KEYS or any(hh.startswith("smsp__average_warps_issue_stalled") and hh.endswith("per_issue_active.ratio") for _ in [0]) :
                d[hh] = (v, u)
        out[name[:60]] = d
        print("==", name[:100])
        for k, (v, u) in d.items():
            try:
                if float(v.replace(",", "")) == 0: continue
            except Exception: pass
            print("   %-90s %s %s" % (k, v, u))
    if "--nosrc" not in sys.argv:
        txt = run(["-i", rep, "--page", "source", "--csv"])
        lines = txt.splitlines()
        # first line names the kernel; the table follows
        k = 0
        while k < len(lines) and not lines[k].startswith('"Address"'): k += 1
        src = list(csv.reader(io.StringIO("\n".join(lines[k:]))))
        h = src[0]
        ci, cs, cx = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        rows = []
        for n, r in enumerate(src[1:]):
            if len(r) <= cs or r[0].startswith("Kernel") or r[0] == "Address": continue
            try: smp = float(r[cs])
            except Exception: continue
            st = sorted(((float(r[i] or 0), x) for i, x in stall_cols), reverse=True)[:2]
            rows.append((n, smp, r[ci].strip(), r[cx], " ".join("%s=%d" % (x[6:], v) for v, x in st if v > 0)))
        tot = sum(x[1] for x in rows) or 1
        if "--listing" in sys.argv:
            with open(sys.argv[sys.argv.index("--listing") + 1], "w") as f:
                for n, smp, ins, ex, st in rows:
                    f.write("%5d %6.2f%% %8s  %-70s %s\n" % (n, 100 * smp / tot, ex, ins[:70], st))
        print("total samples", tot)
        for n, smp, ins, ex, st in sorted(rows, key=lambda x: -x[1])[:top]:
            print("%5d %6.2f%% %8s  %-70s %s" % (n, 100 * smp / tot, ex, ins[:70], st))
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)

main()
