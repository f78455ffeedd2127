#!/bin/bash
# round 2, GPU call T: where the time of the one-launch vppb_pyrlk_prepare goes
mkdir -p gpurun_out
python tools/prep_bench.py
VPPB_PREPARE=streams python tools/prep_bench.py
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_pyrlk_prepare -s 3 -c 1 -f -o gpurun_out/t_prof_prep python tools/prep_bench.py > gpurun_out/t_ncu.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file gpurun_out/t_launches.csv python tools/prep_bench.py > /dev/null 2>&1
grep -E "k_pyrlk|memset|Memset" gpurun_out/t_launches.csv | awk -F'","' '{print $5, $NF}' | head -8
VPPB_PREPARE=streams timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 27 --csv --log-file gpurun_out/t_launches_streams.csv python tools/prep_bench.py > /dev/null 2>&1
grep -E "k_" gpurun_out/t_launches_streams.csv | awk -F'","' '{print $5, $NF}' | head -12
