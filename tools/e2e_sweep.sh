mkdir -p gpurun_out
run() { # plan lanes
  QB200_WAVE_PLAN="$1" QB200_LANES=$2 timeout 200 python bench.py --no-dense --no-cpu-baseline --graph-L 0 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('plan $1 lanes $2 value %.0f e2e %.0f ms %.2f e2e_ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
}
run "16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16" 8
run "16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16" 6
run "32,32,32,32,32,32,32,32" 8
run "32,32,32,32,32,32,32,32" 6
run "8,24,32,32,32,32,32,32,24,8" 8
run "24,24,24,24,24,24,24,24,24,24,16" 8
