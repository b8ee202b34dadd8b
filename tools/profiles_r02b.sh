#!/bin/bash
# Round-2 final profile batch (run under gpurun), after the same-input fusion: launch list of one fused bench step (100 launches),
# ncu --set full of the B-stationary kernel on the fused q/k/v shape, of the skinny kernel on the grouped AdaLN launch, of the
# HBM-bound kernels (a)/(b) + qdm_geglu; per-shape tables of the fused launch inventories.  Every command runs once WITHOUT ncu first.
set -x
bash tools/step_durations.sh
NCU="ncu --set full --clock-control none --import-source on"
python tools/one_gemm.py w4ts 65536 960 320 3 > /dev/null 2>&1 && \
$NCU -k regex:qdm_gemm2_bstat --launch-skip 3 -c 1 -f -o gpurun_out/prof_bstat_65536x960x320 python tools/one_gemm.py w4ts 65536 960 320 3 > gpurun_out/ncu_bstat.log 2>&1
python tools/one_gemm.py w4 1 1104128 2432 3 > /dev/null 2>&1 && \
$NCU -k regex:skinny --launch-skip 3 -c 1 -f -o gpurun_out/prof_skinny_1x1104128x2432 python tools/one_gemm.py w4 1 1104128 2432 3 > gpurun_out/ncu_skinny.log 2>&1
python tools/one_kernel.py && \
$NCU -k regex:"awq_wsum_stage1|quant_pack_awq_kernel|dequant_awq_kernel|quant_group|geglu_kernel" --launch-skip 6 -c 5 -f -o gpurun_out/prof_kernels_ab python tools/one_kernel.py > gpurun_out/ncu_kernels_ab.log 2>&1
python bench.py --layers --fused --out gpurun_out/gemm_layers_fused_r02.json > /dev/null 2>&1
python bench.py --layers --fused --model sdxl --out gpurun_out/gemm_layers_fused_sdxl_r02.json > /dev/null 2>&1
python bench.py --layers --fused --model sd35 --out gpurun_out/gemm_layers_fused_sd35_r02.json > /dev/null 2>&1
python bench.py --mode kernels --out gpurun_out/kernels_ab_r02.json > gpurun_out/kernels_ab_r02.log 2>&1
ls -la gpurun_out | tail -12
