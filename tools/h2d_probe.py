"""Pinned host -> device copy rate of this box (the ceiling of bench.py's e2e line): one JSON line.

  python tools/h2d_probe.py [--mb 928] [--chunks 1,4,16]
"""
import argparse, json, torch

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=928)
    ap.add_argument("--chunks", default="1,5,512")
    a = ap.parse_args()
    n = a.mb * (1 << 20)
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    out = {}
    for c in [int(x) for x in a.chunks.split(",")]:
        step = (n + c - 1) // c
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for o in range(0, n, step):
                dev[o:o + step].copy_(host[o:o + step], non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[f"h2d_GBps_{c}_copies"] = round(best, 2)
    # device -> host for completeness
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); host.copy_(dev, non_blocking=True); e1.record(); torch.cuda.synchronize()
    out["d2h_GBps"] = round(n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
    out["mb"] = a.mb
    print(json.dumps(out))

if __name__ == "__main__":
    main()
