B="--no-dense --no-cpu-baseline --graph-L 0"
for l in 3 4 5 6 8; do
QB200_LANES=$l timeout 200 python bench.py $B 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('lanes $l value %.0f e2e %.0f ms %.2f e2e_ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
done
