#!/bin/bash
# First GPU visit of the next round: everything that was built on the CPU emulator after round 1's GPU budget ran out
# gets its first hardware run and its first timings.  ~6 GPU-minutes on one B200.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2.sh'
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 180 --timeout-method=thread 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout -k 10 300 python tools/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; tail -40 gpurun_out/kernel_bench.txt
timeout -k 10 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
# semi-dense flow: default anti-diagonal schedule vs the opt-in dependency-level schedule, 1080p and 8K
timeout -k 10 300 python tools/sdof_bench.py > gpurun_out/sdof_bench.txt 2>&1; cat gpurun_out/sdof_bench.txt
# PCIe legs of the e2e path: 2-D vs linear copies, duplex overlap, pipeline depth, staged upload + fused copy/mirror
timeout -k 10 200 python dbg/exp_e2e.py > gpurun_out/exp_e2e.txt 2>&1; cat gpurun_out/exp_e2e.txt
# launch list of the pyrLK extra (9 pyramid launches expected) and one full capture of the ingest kernel
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --cpu-budget 1 > /dev/null 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_rgb_to_gray -c 2 -f -o gpurun_out/prof_ingest python bench.py --steps 2 --warmup 3 --cpu-budget 1 > /dev/null 2>&1
# the kernels that have never run on hardware before, once under compute-sanitizer (memcheck, then racecheck on shared memory)
timeout -k 10 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "batch or rgb or fused_level or pyramid" > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -5 gpurun_out/sanitizer_memcheck.txt
timeout -k 10 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "box5x5_batch_equals_oracle and 270" > gpurun_out/sanitizer_racecheck.txt 2>&1; tail -5 gpurun_out/sanitizer_racecheck.txt
ls -la gpurun_out | tail
