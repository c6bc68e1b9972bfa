#!/bin/bash
# final GPU pass of the round: full GPU suite, default bench line, launch lists, ncu --set full of the new K1 kernels, K8 at L = 3000 and
# the pre-processing kernels.  Everything lands in gpurun_out/ under the tag $1.
tag=${1:-fin}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; echo "bench rc=$?"
L="--pairs 64 --slots 64 --steps 1 --warmup 0 --no-cpu-baseline --graph-L 0 --sync-steps"
QB200_LANES=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_street.csv python bench.py $L --no-dense > /dev/null 2>&1; echo "launch street rc=$?"
QB200_LANES=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_dense.csv python bench.py $L --scene dense > /dev/null 2>&1; echo "launch dense rc=$?"
QB200_LANES=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"vsort_|voxel_pack|voxel_bbox|voxel_centroid|tim_graph" -c 40 -f -o gpurun_out/${tag}_prof_k1 python bench.py --pairs 64 --slots 64 --steps 1 --warmup 0 --no-cpu-baseline --no-dense --graph-L 3000 --sync-steps > gpurun_out/${tag}_prof_k1.json 2> /dev/null; echo "ncu k1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pw_|ip_" -c 30 -f -o gpurun_out/${tag}_prof_pre python tools/prof_preprocess.py > gpurun_out/${tag}_prof_pre.json 2> /dev/null; echo "ncu pre rc=$?"
head -c 400 gpurun_out/${tag}_bench_1gpu.json
