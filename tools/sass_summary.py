#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (TMA tensor / bulk loads, mbarrier waits, packed byte
ops, warp votes / shuffles, GPU-scope acquire / release) in vpp_b200/lib/libvppb.so.  Usage: python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "vpp_b200", "lib", "libvppb.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ["UTMALDG", "UBLKCP", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "VABSDIFF4", "VOTE", "SHFL", "LDG.E.STRONG.GPU", "ST.E.STRONG.GPU|STG.E.STRONG.GPU", "ATOM|RED", "LDS", "STS",
        "PRMT", "IADD3", "IMAD", "BAR.SYNC", "WARPSYNC"]
kern, counts, total = None, collections.OrderedDict(), {}
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?(\S+)", l)
    if m and kern:
        op = m.group(1).rstrip(";")
        total[kern] += 1
        for k in KEYS:
            if any(op.startswith(x) for x in k.split("|")):
                counts[kern][k] += 1
print("# SASS evidence (cuobjdump -sass vpp_b200/lib/libvppb.so), instruction counts per kernel\n")
print("`UTMALDG` = cp.async.bulk.tensor (TMA tile load), `UBLKCP` = cp.async.bulk (1-D bulk copy: the peer-memory halo rows), `SYNCS.*` = mbarrier arrive / try_wait, "
      "`VABSDIFF4` = packed byte |a-b|, `*.STRONG.GPU` = acquire / release at GPU scope (dataflow sweeps).\n")
show = [k for k in KEYS]
print("| kernel | instrs | " + " | ".join(k.split("|")[0] for k in show) + " |")
print("|---|---|" + "---|" * len(show))
for kname, c in counts.items():
    d = demangle(kname)
    d = re.sub(r"\(.*", "", d).replace("void vppb::", "").replace("vppb::", "")
    if not any(x in d for x in ("k_box5_stream", "k_box5_bytes_tma<", "k_fast9_band", "k_fast9_emit_bands", "k_sdof_sweep", "k_sdof_match", "k_lk_match_v2<7", "k_add_i32_vec",
                                "k_kpc_merge", "k_rgb_to_gray<3", "k_lowpass_sub2_u8_fast", "k_scharr_u8_v8<", "k_sdof_fused<2", "k_pyrlk_prepare", "k_lbp_u8<true", "k_local_maxima_filter<unsigned char",
                                "k_lk_match_oriented<2, true", "k_fast9_block_rank")):
        continue
    if "k_box5_stream" in d and not re.search(r"<3, 4, 0, (1|128), 0>", d):
        continue
    print("| `%s` | %d | " % (d, total[kname]) + " | ".join(str(c.get(k, 0)) for k in show) + " |")
