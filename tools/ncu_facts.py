#!/usr/bin/env python3
"""Extract the per-launch facts bench.py quotes (DRAM traffic, duration, pipe utilisation) from an `ncu --set full` report and the
JSON line the profiled bench run printed, and write them to profiles/r02_ncu_facts.json (+ a markdown summary).

    python tools/ncu_facts.py gpurun_out/prof_r2c.ncu-rep gpurun_out/r2_prof_c.json [--md profiles/r02_ncu_summary.md] [--key k1_wave]

Kernels: tc_nn_kernel (K6), tim_graph_kernel (K8), kcore_warp_kernel / clique_cta_kernel (K9).  For every kernel the LARGEST launch of
the capture is reported (the captures hold a 64-pair wave plus single-pair passes)."""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
           "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m.get(unit, 1)


def to_ms(v, unit):
    m = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}
    return float(v) * m.get(unit, 1)


def main():
    rep, bench_json = sys.argv[1], sys.argv[2]
    md = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    best = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("qb::", "")
        key = name.split("<")[0]
        rec = {"launch": name}
        for m in METRICS:
            if m in ix and r[ix[m]] != "":
                v, u = r[ix[m]].replace(",", ""), units[ix[m]]
                try:
                    if "bytes" in m:
                        rec[m] = to_bytes(v, u)
                    elif m.startswith("gpu__time"):
                        rec[m] = to_ms(v, u)
                    else:
                        rec[m] = float(v)
                except ValueError:
                    pass
        if key not in best or rec.get("gpu__time_duration.sum", 0) > best[key].get("gpu__time_duration.sum", 0):
            best[key] = rec
    line = json.loads([ln for ln in Path(bench_json).read_text().splitlines() if ln.startswith("{")][-1])   # ncu prints its own lines around it
    facts = {"source": {"report": Path(rep).name, "bench_line": Path(bench_json).name, "bench_config": line["config"]["workload"]}}
    for key, rec in best.items():
        f = {"launch": rec["launch"], "duration_ms": rec.get("gpu__time_duration.sum"),
             "dram_bytes_per_launch": rec.get("dram__bytes_read.sum", 0) + rec.get("dram__bytes_write.sum", 0),
             "dram_read_bytes": rec.get("dram__bytes_read.sum"), "dram_write_bytes": rec.get("dram__bytes_write.sum"),
             "issue_active_pct": rec.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "warps_active_pct": rec.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
             "tensor_pipe_active_pct": rec.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
             "fma_pipe_active_pct": rec.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
             "alu_pipe_active_pct": rec.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
             "warp_instructions": rec.get("sm__inst_executed.sum"), "registers_per_thread": rec.get("launch__registers_per_thread"),
             "grid": rec.get("launch__grid_size"), "block": rec.get("launch__block_size")}
        if key == "tc_nn_kernel":
            f["algorithmic_flops_per_launch"] = line["roofline"]["flops_per_launch"]
            f["algorithmic_bytes_per_launch"] = line["roofline"].get("algorithmic_bytes_per_launch")
        facts[key] = f
    out = ROOT / "profiles" / "r02_ncu_facts.json"
    if "--key" in sys.argv:   # merge this capture under its own key, leave the rest of the file alone
        allf = json.loads(out.read_text()) if out.exists() else {}
        allf[sys.argv[sys.argv.index("--key") + 1]] = facts
        out.write_text(json.dumps(allf, indent=1))
    else:
        out.write_text(json.dumps(facts, indent=1))
    print(json.dumps(facts, indent=1))
    if md:
        with open(md, "w") as fh:
            fh.write(f"# Round 2 - ncu --set full summary\n\nReport `{Path(rep).name}` (one {line['config']['workload']}); values per launch, the largest launch of every kernel.\n\n")
            fh.write("| kernel | ms | DRAM read MB | DRAM write MB | issue active % | warps active % | tensor pipe % | fma pipe % | alu pipe % | regs |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
            for key, f in facts.items():
                if key == "source":
                    continue
                g = lambda k, s=1.0: "-" if f.get(k) is None else f"{f[k] * s:.2f}"
                fh.write(f"| `{f['launch'][:48]}` | {g('duration_ms')} | {g('dram_read_bytes', 1e-6)} | {g('dram_write_bytes', 1e-6)} | {g('issue_active_pct')} | "
                         f"{g('warps_active_pct')} | {g('tensor_pipe_active_pct')} | {g('fma_pipe_active_pct')} | {g('alu_pipe_active_pct')} | {g('registers_per_thread')} |\n")


if __name__ == "__main__":
    main()
