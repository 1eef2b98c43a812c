#!/bin/bash
# pass 2 of tools/gpu_profile.sh alone: time, DRAM bytes and tensor-pipe activity of every convolution launch of ONE timed
# step (the first after the 3 warm-up steps).  NCONV = convolution launches per step (len of the engine's conv plans).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3"
NCONV=${NCONV:-69}
$CMD > gpurun_out/prof_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:conv_igemm|stem_pool|conv3x3_rows' -s $((3 * NCONV)) -c $NCONV --csv --log-file gpurun_out/conv_step.csv $CMD > gpurun_out/ncu_convstep.log 2>&1
echo "conv step exit $?"
