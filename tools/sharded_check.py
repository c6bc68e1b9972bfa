#!/usr/bin/env python3
"""torchrun helper: register N synthetic pairs sharded over WORLD_SIZE ranks and compare with rank 0's
single-GPU results.  Usage: torchrun --nproc-per-node 2 tools/sharded_check.py [n_pairs] [backend]
With backend=gloo the per-pair 'registration' is a deterministic fake record (host-logic test on CPU)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist
    from quatro_b200.capi import RESULT_DTYPE
    from quatro_b200.sharded import register_sharded, shard_indices

    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    backend = sys.argv[2] if len(sys.argv) > 2 else "nccl"
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        device = torch.device("cuda", local_rank)
        from quatro_b200 import synth
        from quatro_b200.capi import Handle, default_params
        handle = Handle(device=local_rank, max_batch_slots=4)
        p = default_params()

        def register_local(ids):
            return handle.register_batch([synth.outdoor_pair(int(i), rings=32, azimuths=900)[:2] for i in ids], p)
    else:
        dist.init_process_group("gloo")
        device = None

        def register_local(ids):
            out = np.zeros(len(ids), RESULT_DTYPE)
            out["valid"] = 1
            out["n_corr"] = 100 + np.asarray(ids)
            out["T"][:, 12] = np.asarray(ids) * 0.5
            return out

    if backend == "nccl" and len(sys.argv) > 3 and sys.argv[3] == "cabi":
        # the C-ABI path: qb200_comm_init_rank + qb200_register_batch_rank (ncclAllGather inside the library, deferred gather)
        from quatro_b200.capi import Pair, MEM_HOST, RESULT_DTYPE as RD
        uid = [Handle.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        handle.comm_init_rank(world, rank, uid[0])
        n_local = (n_pairs + world - 1) // world
        ids = [i * world + rank for i in range(n_local)]            # round-robin: global pair g = i * world + r (ids beyond n_pairs: extra work, ignored)
        prs = [synth.outdoor_pair(int(i), rings=32, azimuths=900)[:2] for i in ids]
        arr = (Pair * n_local)()
        for k, (s_, t_) in enumerate(prs):
            arr[k].src, arr[k].n_src, arr[k].tgt, arr[k].n_tgt = s_.ctypes.data, len(s_), t_.ctypes.data, len(t_)
        allrec = np.zeros(world * n_local, RD)
        handle.register_batch_rank_raw(arr, n_local, p, MEM_HOST, allrec, defer=True)
        handle.comm_wait()
        full = allrec[:n_pairs].copy()
    else:
        full = register_sharded(register_local, n_pairs, rank, world, device)
    assert len(full) == n_pairs
    ok = True
    if rank == 0:
        ref = register_local(np.arange(n_pairs))            # every pair on ONE rank
        ok = full.tobytes() == ref.tobytes()
        print(f"sharded_check world={world} n_pairs={n_pairs} backend={backend}: {'IDENTICAL' if ok else 'MISMATCH'}")
    sizes = [len(shard_indices(n_pairs, r, world)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
