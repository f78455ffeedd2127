#!/bin/bash
for th in 16 8; do
VPPB_BOX_TH=$th timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_box5_bytes_tma -s 30 -c 2 -f -o gpurun_out/prof_box4k_th$th python bench.py --steps 2 --warmup 3 --no-extras --cpu-budget 0.2 --workload 4k --streams 1 --graph 0 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
