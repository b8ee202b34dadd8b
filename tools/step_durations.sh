#!/bin/bash
# GPU-side duration and DRAM bytes of every launch of one bench step (ncu, cold cache, serialised)
#   -> gpurun_out/launches_step.csv ; aggregate with tools/step_by_shape.py
python bench.py --steps 1 --warmup 3 --no-graph > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"qdm_gemm|smallm" \
    --launch-skip 552 -c 184 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-graph \
    > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log | cut -c1-200
