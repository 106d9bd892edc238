#!/bin/bash
# Profiling evidence of a round (one GPU): the launch list of a short eager bench run, and one `ncu --set full` capture of
# the hot kernels.  Each ncu pass only runs after the same command line has exited 0 without ncu.
#     gpurun -- tools/gpu_ncu.sh r02 [full|list|both]
TAG=${1:-r02}; WHAT=${2:-both}
export QF_GRAPH=0     # eager launches: kernel nodes inside a graph with a conditional WHILE node are invisible to ncu
CMD="python bench.py --n 2048 --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline"
if [ "$WHAT" != "full" ]; then
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_n2048.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
fi
if [ "$WHAT" != "list" ]; then
# Eager mode enqueues maxit = 10 gated iterations per step (5 matching kernels each: Poisson, GEMM 1, GEMM 2, tail,
# control).  Skip the warm-up step (50 launches), capture the first live iteration of the first timed step; the update
# kernel in a pass of its own.  (A capture of all 51 launches of a step exceeds the 64 MiB that come back from the box.)
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_zgemm3m_ws|k_poisson_band|k_post|k_control' -s 50 -c 5 -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_update' -s 1 -c 1 -f -o gpurun_out/${TAG}_full_update $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
fi
ls -la gpurun_out/${TAG}_* | head; tail -3 gpurun_out/${TAG}_ncu2.log
