#!/bin/bash
# ncu evidence for the kernels added after round-1i: AR beam step kernels, packed attention backward, dgrad GEMM.
# One B200; reports stay in /tmp on the box, only the tables travel back.
set -x
mkdir -p gpurun_out /tmp/ncu
export NAVC_GRAPHS=0
timeout 300 ncu --set full --clock-control none -k regex:"self_attention_step|beam_topk|beam_advance" -s 30 -c 9 -f -o /tmp/ncu/prof_ar \
    python tools/ar_bench.py --batch 128 --steps 1 > gpurun_out/ncu_ar.log 2>&1
python tools/ncu_table.py /tmp/ncu/prof_ar.ncu-rep > gpurun_out/r1n_ncu_table_ar_step.txt
timeout 300 ncu --set full --clock-control none -k regex:"attn_bwd" -s 8 -c 8 -f -o /tmp/ncu/prof_ab \
    python tools/attn_bwd_bench.py > gpurun_out/ncu_ab.log 2>&1
python tools/ncu_table.py /tmp/ncu/prof_ab.ncu-rep > gpurun_out/r1n_ncu_table_attn_bwd.txt
timeout 300 ncu --set full --clock-control none -k regex:"gemm_tc_kernel<.*(4|5), (128|256)>" -s 200 -c 10 -f -o /tmp/ncu/prof_dg \
    python tools/train_bench.py --method NAB --batch 256 --steps 1 --warmup 3 > gpurun_out/ncu_dg.log 2>&1
python tools/ncu_table.py /tmp/ncu/prof_dg.ncu-rep > gpurun_out/r1n_ncu_table_dgrad_wgrad.txt
rm -f gpurun_out/ncu_*.log
ls -la gpurun_out
