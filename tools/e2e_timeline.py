"""Stage boundaries of every wave of one host-buffer batch (QB200_TIMELINE=1 prints them to stderr).

  QB200_TIMELINE=1 python tools/e2e_timeline.py [--pairs 256] [--slots 64] [--device]
"""
import argparse, os, sys
os.environ.setdefault("QB200_TIMELINE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from quatro_b200 import capi, synth
import bench

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--slots", type=int, default=64)
    a = ap.parse_args()
    p = bench.scene_params("street")
    prs = [synth.outdoor_pair(1000 + i)[:2] for i in range(a.pairs)]
    total = sum(len(s) + len(t) for s, t in prs)
    host = torch.empty((total, 4), dtype=torch.float32).pin_memory()
    hv = host.numpy()
    arr = (capi.Pair * a.pairs)()
    o = 0
    for i, (s, t) in enumerate(prs):
        hv[o:o + len(s)] = s; arr[i].src, arr[i].n_src = host.data_ptr() + o * 16, len(s); o += len(s)
        hv[o:o + len(t)] = t; arr[i].tgt, arr[i].n_tgt = host.data_ptr() + o * 16, len(t); o += len(t)
    h = capi.Handle(device=0, max_batch_slots=a.slots)
    out = np.zeros(a.pairs, capi.RESULT_DTYPE)
    for rep in range(3):
        print(f"--- batch {rep}", file=sys.stderr, flush=True)
        h.register_batch_raw(arr, a.pairs, p, capi.MEM_HOST, out)

if __name__ == "__main__":
    main()
