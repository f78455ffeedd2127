#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "lucas or pyrlk" 2>&1 | tail -3
timeout 300 python bench.py --cpu-budget 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps(d['extras'], indent=1)); print('value', d['value'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'])"
VPPB_LK_V1=1 timeout 300 python bench.py --cpu-budget 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('v1', json.dumps(d['extras']['pyrlk_1080p_10k']))"
