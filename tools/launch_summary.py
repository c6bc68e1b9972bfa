#!/usr/bin/env python3
"""profiles/r02_launches_summary.md from the ncu launch lists (profiles/r02_launches_street.csv / _dense.csv).

    python tools/launch_summary.py
"""
import collections
import csv
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CMD = ("QB200_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv python bench.py [--scene dense] --pairs 64 "
       "--slots 64 --steps 1 --warmup 0 --no-cpu-baseline --graph-L 0 --sync-steps")


def table(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki]].append(float(r[vi].replace(",", "")) / 1000.0)
        except ValueError:
            pass
    total = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | max us per launch | share of the capture |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k[:60]}` | {len(v)} | {max(v):.1f} | {100 * sum(v) / total:.1f}% |")
    wave = sum(max(v) for v in agg.values() if len(v) <= 16) + sum(max(v) * (len(v) / 8.0) for v in agg.values() if len(v) > 16)
    return "\n".join(out), wave


def main():
    md = ["# Round 2 - ncu launch lists", "", f"Command: `{CMD}`",
          "(one device-resident 64-pair wave, one host-buffer wave, single-pair passes; serialised and cold-cache: compare SHARES; "
          "`max us` = the 64-pair launch).", ""]
    for name, f, what in (("street", "r02_launches_street.csv", "config/params.yaml defaults, mean L = 306"),
                          ("dense preset", "r02_launches_dense.csv", "voxel 0.22, tuple test off, mean L = 2970")):
        p = ROOT / "profiles" / f
        if not p.exists():
            continue
        t, wave = table(p)
        md += [f"## {name} ({what})", "", f"Sum of the 64-pair launches of one wave (serialised, cold): {wave / 1000:.2f} ms.", "", t, ""]
    (ROOT / "profiles" / "r02_launches_summary.md").write_text("\n".join(md))
    print("\n".join(md[:12]))


if __name__ == "__main__":
    main()
