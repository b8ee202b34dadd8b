#!/bin/bash
# Round-2 profile batch (run under gpurun): launch list of one bench step, ncu --set full captures of the TS GEMM and the clip
# search kernels, per-shape tables of the three denoisers.  Every command runs once WITHOUT ncu first.
set -x
bash tools/step_durations.sh
NCU="ncu --set full --clock-control none --import-source on"
for s in "4096 1280 1280" "1232 1280 768" "4096 2432 2432"; do
  tag=$(echo $s | tr ' ' x)
  python tools/one_gemm.py w4ts $s 3 > /dev/null 2>&1 && \
  $NCU -k regex:qdm_w4ts_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_ts_$tag python tools/one_gemm.py w4ts $s 3 > gpurun_out/ncu_ts_$tag.log 2>&1
done
python tools/one_clip.py 9728 2432 512 2 && \
$NCU -k regex:"group_gram|awq_clip" --launch-skip 4 -c 2 -f -o gpurun_out/prof_clip_9728x2432 python tools/one_clip.py 9728 2432 512 2 > gpurun_out/ncu_clip.log 2>&1
python bench.py --layers --out gpurun_out/gemm_layers_r02.json > /dev/null 2>&1
python bench.py --layers --model sdxl --out gpurun_out/gemm_layers_sdxl_r02.json > /dev/null 2>&1
python bench.py --layers --model sd35 --out gpurun_out/gemm_layers_sd35_r02.json > /dev/null 2>&1
ls -la gpurun_out | tail -12
