#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "box" 2>&1 | tail -3
for th in 8 16; do for st in 1 4 8; do for g in 0 1; do
echo "TH=$th streams=$st graph=$g"
VPPB_BOX_TH=$th timeout 120 python bench.py --no-extras --cpu-budget 0.3 --streams $st --graph $g 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  value %.0f Mpix/s  ms/step %.4f  kernel us/launch %.2f frac %.3f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['us_per_launch'], d['roofline']['frac'], d['parity_checked']))
    elif 'rror' in l: print(l.strip()[:300])
"
done; done; done
for th in 8 16; do echo "4K TH=$th"; VPPB_BOX_TH=$th timeout 120 python bench.py --no-extras --cpu-budget 0.3 --workload 4k --streams 1 --graph 1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  value %.0f Mpix/s  ms/step %.4f  kernel us/launch %.2f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['us_per_launch'], d['roofline']['frac']))
"; done
