#!/bin/bash
# quick A/B: FPFH parity tests + the street bench line (value, stage lane-times)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fpfh or register_pair or batch_matches" 2>&1 | tail -2
timeout 300 python bench.py --no-dense --no-cpu-baseline --graph-L 0 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v,2) for k,v in d['stages_ms_per_step'].items()})"
