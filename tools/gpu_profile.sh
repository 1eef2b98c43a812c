#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): plain run first, then
#  (1) the launch list, (2) DRAM bytes + time of every conv launch of one timed step, (3) --set full on a few launches.
mkdir -p gpurun_out
CMD="python bench.py --no-extras --steps 2 --warmup 3"
NCONV=${NCONV:-63}     # convolution launches per step (engine plan list); the 3 warm-up steps are skipped
$CMD > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:conv_igemm|stem_pool|conv3x3_rows' -s $((3 * NCONV)) -c $NCONV --csv --log-file gpurun_out/conv_step.csv $CMD > gpurun_out/ncu_convstep.log 2>&1
echo "conv step exit $?"
$CMD > gpurun_out/prof_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:conv_igemm|stem_pool|conv3x3_rows' -s $((3 * NCONV)) -c 24 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | tail -12
