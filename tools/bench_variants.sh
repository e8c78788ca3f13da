#!/bin/bash
# A/B runs of bench.py (inference only) over environment variants: tools/bench_variants.sh <tag> "VAR=val ..." ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline > gpurun_out/${TAG}_v${i}.json 2> gpurun_out/${TAG}_v${i}.err
  echo "== $v" >> gpurun_out/${TAG}_summary.txt
  python tools/bench_summary.py gpurun_out/${TAG}_v${i}.json >> gpurun_out/${TAG}_summary.txt 2>&1
done
cat gpurun_out/${TAG}_summary.txt
