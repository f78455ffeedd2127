#!/bin/bash
# One GPU visit: parity tests, bench, ncu launch list, ncu full capture of the box kernel.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 180 --timeout-method=thread 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout -k 10 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout -k 10 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err; tail -c 1200 gpurun_out/bench_ref.json
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras --cpu-budget 1 > /dev/null 2>&1
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:k_box5_bytes_tma -s 40 -c 3 -f -o gpurun_out/prof_box python bench.py --steps 2 --warmup 3 --no-extras --cpu-budget 1 > /dev/null 2>&1
ls -la gpurun_out | tail
cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os; print('affinity', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count())"
