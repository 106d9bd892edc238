#!/bin/bash
# Profiling evidence of a round (one GPU): the launch list of a short eager bench run, and one `ncu --set full` capture of
# the hot kernels.  Each ncu pass only runs after the same command line has exited 0 without ncu.
#     gpurun -- tools/gpu_ncu.sh r02
TAG=${1:-r02}
export QF_GRAPH=0     # eager launches: kernel nodes inside a graph with a conditional WHILE node are invisible to ncu
CMD="python bench.py --n 2048 --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_n2048.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_zgemm3m_ws|k_poisson_band|k_post|k_update|k_control' -s 12 -c 10 -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/${TAG}_* | head; tail -3 gpurun_out/${TAG}_ncu2.log
