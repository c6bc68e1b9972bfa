for kb in 144 100 200; do
  QB200_KCORE_SMEM_KB=$kb timeout 200 python bench.py --no-dense --no-cpu-baseline --graph-L 0 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('kcore smem $kb KB: value %.0f e2e %.0f clique_stage %.2f' % (d['value'], d['e2e']['value'], d['stages_ms_per_step']['clique']))"
done
