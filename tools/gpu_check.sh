#!/bin/bash
# One GPU-box visit: parity tests, bench lines, per-kernel breakdown, ncu launch list and full captures.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python tools/profile_step.py bf16x3 > gpurun_out/breakdown_bf16x3.txt 2>&1
python tools/profile_step.py bf16 > gpurun_out/breakdown_bf16.txt 2>&1
if [ "$1" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 400 -c 12 -f -o gpurun_out/prof_gemm \
    python tools/profile_step.py bf16x3 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn -s 40 -c 4 -f -o gpurun_out/prof_attn \
    python tools/profile_step.py bf16x3 > gpurun_out/ncu_attn.log 2>&1
fi
cat gpurun_out/bench_bf16x3.json gpurun_out/bench_bf16.json
