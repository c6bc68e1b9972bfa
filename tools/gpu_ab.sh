#!/bin/bash
# full GPU suite with the default build, then the bench with the own voxel sort and with the library sort
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/ab_pytest.log
B="--no-dense --no-cpu-baseline --graph-L 0"
timeout 300 python bench.py $B > gpurun_out/ab_bench_own.json 2> gpurun_out/ab_bench_own.err; echo "own rc=$?"
QB200_VOXEL_SORT=cub timeout 300 python bench.py $B > gpurun_out/ab_bench_cub.json 2> gpurun_out/ab_bench_cub.err; echo "cub rc=$?"
python - <<'P'
import json
for t in ("own","cub"):
    try:
        d=json.loads(open(f"gpurun_out/ab_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "value %.0f e2e %.0f ms %.2f voxel-stage %.2f lat %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["stages_ms_per_step"]["voxel"], d["single_pair_latency_ms"]))
    except Exception as e:
        print(t, "failed", e)
P
