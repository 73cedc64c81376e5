#!/bin/bash
# Round-2 final pass on one B200 (under gpurun from the repo root): the default bench line, the reference arm, the ncu launch
# list of the same bench command (each only after its command has exited 0 without ncu) and full captures of the two flow
# pipeline kernels.  Outputs: gpurun_out/.
set -u
o=gpurun_out
python bench.py > $o/r2_bench_final.json 2> $o/r2_bench_final.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $o/r2_bench_reference.json 2> $o/r2_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $o/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-ref-python > $o/r2_ncu_launches.log 2>&1
N="ncu --set full --clock-control none --import-source on -c 1 -f"
$N -k regex:k_flow_pipe_precise -o $o/r2_k4_flow_pipe_precise python profiles/flow_ncu_target.py precise > $o/r2_ncu_k4pp.log 2>&1
$N -k 'regex:k_flow_pipe$' -o $o/r2_k4_flow_pipe python profiles/flow_ncu_target.py fast > $o/r2_ncu_k4p.log 2>&1
ls -la $o/*.ncu-rep $o/r2_launches.csv
