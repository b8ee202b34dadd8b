#!/bin/bash
# GPU-side duration and DRAM bytes of every launch of one bench step (ncu, cold cache, serialised)
#   -> gpurun_out/launches_step.csv ; aggregate with tools/step_by_shape.py
# The step as bench.py launches it: 100 launches (same-input Linears fused); `UNFUSED=1` for the 184-launch per-Linear step.
# Launch order of `bench.py --steps 1 --warmup 3 --no-graph --no-extras`: build pass, 3 warm-up steps, THE TIMED STEP, ...
N=100; EXTRA=""
if [ -n "$UNFUSED" ]; then N=184; EXTRA="--no-fuse"; fi
python bench.py --steps 1 --warmup 3 --no-graph --no-extras $EXTRA > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"qdm_gemm|smallm|skinny|qdm_w4ts|qdm_w4rp" \
    --launch-skip $((4 * N)) -c $N --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-graph --no-extras $EXTRA \
    > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log | cut -c1-200
