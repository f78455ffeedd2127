#!/bin/bash
# round 2, GPU call L: parity + semi-dense flow timing after the 64-bit cell records (one acquire load per hop)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -x -q -m gpu > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
tail -3 gpurun_out/l_pytest.log
cat > /tmp/sdof_t.py <<'PY'
import sys, json, ctypes as C, numpy as np
sys.path.insert(0, '.')
import torch, bench
import vpp_b200 as vpp
from vpp_b200 import capi
capi.check(capi.lib.vppb_init(0))
st = torch.cuda.current_stream(); sp = C.c_void_p(st.cuda_stream)
ex = bench.gpu_extras.__code__
# reuse the bench's own rows
import types
out = {}
full = bench.gpu_extras(vpp, capi, torch, st, sp, torch.device("cuda", 0))
for k in ("sdof_1080p", "sdof_8k", "fast9_4k"):
    out[k] = full[k]
print(json.dumps(out))
PY
timeout 600 python /tmp/sdof_t.py > gpurun_out/l_sdof.json 2> gpurun_out/l_sdof.err
cat gpurun_out/l_sdof.json | cut -c1-1200
tail -3 gpurun_out/l_sdof.err
