#!/bin/bash
# Round-2 profiling pass on one B200 (run under gpurun from the repo root, after the same commands have exited 0 without ncu):
# the launch list of the default bench command and one `--set full` capture of each shipped hot kernel.  Outputs: gpurun_out/.
set -u
o=gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $o/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $o/r2_ncu_launches.log 2>&1
N="ncu --set full --clock-control none --import-source on -c 1 -f"
$N -k regex:k_global_mcmc -o $o/r2_k1_global $B > $o/r2_ncu_k1.log 2>&1
$N -k regex:k_isir -o $o/r2_k2_isir $B --sampler glmcmc > $o/r2_ncu_k2.log 2>&1
$N -k regex:k_mala_fast -o $o/r2_k3_mala_fast_262k $B --sampler glmala --chains 262144 --iters 1001 --layout none > $o/r2_ncu_k3.log 2>&1
$N -k k_flow -o $o/r2_k4_flow_fast python profiles/flow_ncu_target.py fast > $o/r2_ncu_k4a.log 2>&1
$N -k k_flow -o $o/r2_k4_flow_precise python profiles/flow_ncu_target.py precise > $o/r2_ncu_k4b.log 2>&1
$N -k regex:k_flow_bwd -o $o/r2_k4_flow_bwd python profiles/flow_ncu_target.py train > $o/r2_ncu_k4c.log 2>&1
$N -k regex:k_kde_logprob -o $o/r2_k5_kde python bench.py --sampler kde --steps 1 --warmup 0 --no-e2e --no-cpu > $o/r2_ncu_k5.log 2>&1
ls -la $o/r2_*.ncu-rep $o/r2_launches.csv
