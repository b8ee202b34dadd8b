#!/bin/bash
# GPU-side duration and DRAM bytes of every launch of one bench step (ncu, cold cache, serialised)
#   -> gpurun_out/launches_step.csv ; aggregate with tools/step_by_shape.py
python bench.py --steps 1 --warmup 3 --no-graph --no-extras > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"qdm_gemm|smallm|qdm_w4ts|qdm_w4rp" \
    --launch-skip 736 -c 184 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-graph --no-extras \
    > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log | cut -c1-200
