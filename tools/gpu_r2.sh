#!/bin/bash
# Round-2 GPU evidence, one B200 (never multi-rank under ncu).  Usage: tools/gpu_r2.sh <tag> [stages...]
#   tests | bench | ref | launches | classes | sass
set -x
TAG=${1:-r2}; shift
STAGES=${@:-tests bench ref launches classes}
mkdir -p gpurun_out /tmp/ncu
for st in $STAGES; do case $st in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/${TAG}_tests.log ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_bf16x3.json 2> gpurun_out/${TAG}_bench_bf16x3.err ;;
bench_bf16)
  timeout 600 python bench.py --steps 20 --warmup 5 --precision bf16 --no-train > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err ;;
ref)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_ref.err ;;
launches)
  NAVC_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file /tmp/ncu/launches.csv python tools/profile_step.py bf16x3 128 range > gpurun_out/${TAG}_ncu_launch.log 2>&1
  python tools/ncu_launches.py /tmp/ncu/launches.csv 0 > gpurun_out/${TAG}_ncu_launches_bf16x3.txt ;;
classes)
  NAVC_GRAPHS=0 timeout 900 ncu --set full --clock-control none --profile-from-start off -c 80 -f -o /tmp/ncu/prof_layer \
      python tools/profile_step.py bf16x3 128 range > gpurun_out/${TAG}_ncu_layer.log 2>&1
  python tools/ncu_table.py /tmp/ncu/prof_layer.ncu-rep > gpurun_out/${TAG}_ncu_table_layer_bf16x3.txt
  python tools/ncu_classes.py /tmp/ncu/prof_layer.ncu-rep bf16x3 gpurun_out/${TAG}_ncu_classes_bf16x3.json > /dev/null ;;
gemmprof)
  # full capture (with source correlation) of the first decoder layer's GEMMs of one eager step; the report travels back
  NAVC_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm" -s 5 -c 6 -f \
      -o gpurun_out/${TAG}_prof_gemm python tools/profile_step.py bf16x3 128 range > gpurun_out/${TAG}_ncu_gemm.log 2>&1
  python tools/ncu_metrics.py gpurun_out/${TAG}_prof_gemm.ncu-rep > gpurun_out/${TAG}_ncu_gemm_metrics.txt ;;
trainprof)
  # ncu table of the attention kernels of one NACF training step (forward cores + tcgen05 / CUDA-core backward)
  timeout 900 ncu --set full --clock-control none -k regex:"attn" -s 48 -c 8 -f -o /tmp/ncu/prof_train_attn \
      python tools/train_bench.py --method NACF --batch 256 --steps 1 --warmup 2 > gpurun_out/${TAG}_ncu_train.log 2>&1
  python tools/ncu_table.py /tmp/ncu/prof_train_attn.ncu-rep > gpurun_out/${TAG}_ncu_table_train_attention.txt ;;
esac; done
ls -la gpurun_out | tail -20
