#!/bin/bash
# GPU-side duration of every launch of one bench step (ncu, cold cache, serialised) -> gpurun_out/launches_step.csv
python bench.py --steps 1 --warmup 3 --no-graph > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qdm_gemm" --launch-skip 552 -c 184 --csv \
    --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log | cut -c1-200
