#!/bin/bash
# ncu evidence for every kernel of the hot path (one B200, never multi-rank).  The .ncu-rep files are
# tabulated on the box (tools/ncu_table.py) and only the tables + one small report travel back.
set -x
mkdir -p gpurun_out /tmp/ncu
export NAVC_GRAPHS=0
# launch list of one eager step (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 340 --csv --log-file /tmp/ncu/launches.csv \
    python tools/profile_step.py bf16x3 > gpurun_out/ncu_launch.log 2>&1
python tools/ncu_launches.py /tmp/ncu/launches.csv 0 > gpurun_out/ncu_launches_bf16x3.txt
# full captures: ~2 decoder layers of the 4th step (embed, GEMMs, attention), then vocabulary GEMM + decode-step kernels
ncu --set full --clock-control none -s 1015 -c 30 -f -o /tmp/ncu/prof_layer python tools/profile_step.py bf16x3 > gpurun_out/ncu_layer.log 2>&1
python tools/ncu_table.py /tmp/ncu/prof_layer.ncu-rep > gpurun_out/ncu_table_layer_bf16x3.txt
ncu --set full --clock-control none -k regex:"gemm_tc_kernel<.*1, 256>|refine_step|select_best|length_head|highway_bn|length_beam|init_canvas|split_bf16" -s 9 -c 9 -f -o /tmp/ncu/prof_misc \
    python tools/profile_step.py bf16x3 > gpurun_out/ncu_misc.log 2>&1
python tools/ncu_table.py /tmp/ncu/prof_misc.ncu-rep > gpurun_out/ncu_table_misc_bf16x3.txt
ncu --set full --clock-control none -s 1015 -c 20 -f -o /tmp/ncu/prof_layer_bf16 python tools/profile_step.py bf16 > gpurun_out/ncu_layer_bf16.log 2>&1
python tools/ncu_table.py /tmp/ncu/prof_layer_bf16.ncu-rep > gpurun_out/ncu_table_layer_bf16.txt
# training kernels
ncu --set full --clock-control none -k regex:"attn_bwd|transpose_pack|layernorm_bwd|embed_ln_bwd|drop_add|act_drop|log_softmax|bn_|highway|mean_bwd|clip_adam" -s 300 -c 30 -f -o /tmp/ncu/prof_train \
    python tools/train_bench.py --method NACF --batch 256 --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
python tools/ncu_table.py /tmp/ncu/prof_train.ncu-rep > gpurun_out/ncu_table_train.txt
# one report with source correlation for the dominant kernel (kept small: 2 launches)
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel<.*2, 256>" -s 20 -c 2 -f -o gpurun_out/prof_gemm_top \
    python tools/profile_step.py bf16x3 > gpurun_out/ncu_gemm_top.log 2>&1
ls -la /tmp/ncu gpurun_out | head -40
