"""Turn ncu reports / launch lists under gpurun_out/ into the committed text summaries under profiles/."""
import csv, collections, io, re, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]

def summarize(rep, title):
    hdr, units, rows = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"## {title}", f"source: `{rep}` (ncu --set full --clock-control none --import-source on)", ""]
    for r in rows:
        lines.append(f"kernel: `{r[ix['Kernel Name']]}`")
        for k in KEYS:
            if k in ix:
                lines.append(f"  {k:75s} {r[ix[k]]} {units[ix[k]]}")
        lines.append("")
    return "\n".join(lines)

def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    tot = collections.Counter(); op = collections.Counter(); ninst = 0; nsamp = 0
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        try:
            n = float(r[ix["Instructions Executed"]]); s = float(r[ix["# Samples"]])
        except ValueError:
            continue
        ninst += n; nsamp += s
        for h, i in ix.items():
            if h.startswith("stall_") and "Not Issued" not in h:
                try: tot[h] += float(r[i])
                except ValueError: pass
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        if m: op[m.group(2).split(".")[0]] += n
    lines = [f"warp-instructions executed: {ninst:.4g}; stall samples: {int(nsamp)}",
             "stall reasons: " + ", ".join(f"{k[6:]} {100 * v / nsamp:.1f}%" for k, v in tot.most_common(8)),
             "opcode mix: " + ", ".join(f"{k} {100 * v / ninst:.1f}%" for k, v in op.most_common(16)), ""]
    return "\n".join(lines)

def launches(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try: v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError: continue
        u = r[ix["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        nm = r[ix["Kernel Name"]].split("(")[0]
        agg[nm][0] += 1; agg[nm][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"## {title}", f"source: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)", "",
             "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
    return "\n".join(lines) + "\n"

if __name__ == "__main__":
    what, path, title = sys.argv[1], sys.argv[2], sys.argv[3]
    if what == "launches": print(launches(path, title))
    else:
        print(summarize(path, title)); print(stalls(path))
