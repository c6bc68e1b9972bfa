#!/bin/bash
# One gpurun call of a round: GPU tests, the default bench line, ncu launch lists of one street and one dense wave,
# and one ncu --set full capture of K8 at L = 3000 x 32 sets.  Everything lands in gpurun_out/ under the tag $1.
tag=${1:-r}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; echo "bench rc=$?"
L="--pairs 64 --slots 64 --steps 1 --warmup 0 --no-cpu-baseline --graph-L 0 --sync-steps"
QB200_LANES=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_street.csv python bench.py $L --no-dense > /dev/null 2>&1; echo "launch street rc=$?"
QB200_LANES=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_dense.csv python bench.py $L --scene dense > /dev/null 2>&1; echo "launch dense rc=$?"
QB200_LANES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tim_graph -c 2 -f -o gpurun_out/${tag}_prof_k8 python bench.py --pairs 8 --slots 8 --steps 1 --warmup 0 --no-cpu-baseline --no-dense --graph-L 3000 > gpurun_out/${tag}_prof_k8.json 2> /dev/null; echo "ncu k8 rc=$?"
head -c 600 gpurun_out/${tag}_bench_1gpu.json
