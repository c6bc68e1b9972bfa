"""Stage times of ONE pair through qb200_register_pair (host buffers): median of 10 calls, plus the wall latency.

  python tools/single_pair_stages.py            (under ncu --metrics gpu__time_duration.sum: the per-kernel list of one pair)
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quatro_b200 import capi, synth
import bench

def main():
    p = bench.scene_params("street")
    s, t = synth.outdoor_pair(1000)[:2]
    hs = torch.from_numpy(np.ascontiguousarray(s)).pin_memory(); ht = torch.from_numpy(np.ascontiguousarray(t)).pin_memory()
    h = capi.Handle(device=0, max_batch_slots=1)
    arr = (capi.Pair * 1)()
    arr[0].src, arr[0].n_src, arr[0].tgt, arr[0].n_tgt = hs.data_ptr(), len(s), ht.data_ptr(), len(t)
    out = np.zeros(1, capi.RESULT_DTYPE)
    st, lat = [], []
    n = int(os.environ.get("REPS", "12"))
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        h.register_batch_raw(arr, 1, p, capi.MEM_HOST, out)
        e1.record(); torch.cuda.synchronize()
        if i >= 2 or n < 3:
            st.append(h.stage_ms()); lat.append(e0.elapsed_time(e1))
    names = ["h2d", "voxel", "fpfh", "match", "graph", "clique", "pose", "d2h"]
    print(json.dumps({"latency_ms": float(np.median(lat)), "stages_ms": {k: round(float(v), 4) for k, v in zip(names, np.median(np.array(st), axis=0))},
                      "n_corr": int(out[0]["n_corr"]), "clique": int(out[0]["clique_size"]), "launches": h.launch_count()}))

if __name__ == "__main__":
    main()
