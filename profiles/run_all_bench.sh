#!/bin/bash
# Round-1 measurement pass on one B200 (run under gpurun from the repo root): every bench line, the ncu launch list of the
# default command, and one full-size ncu capture of the iSIR kernel for its DRAM traffic.  Outputs land in gpurun_out/.
set -u
o=gpurun_out
python bench.py > $o/fin_bench.json 2> $o/fin_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $o/fin_bench_reference.json 2> $o/fin_bench_reference.err
for s in glmcmc glmala aglmcmc kde; do
  python bench.py --sampler $s --steps 10 --warmup 3 > $o/fin_bench_$s.json 2> $o/fin_bench_$s.err
done
for s in glmcmc_nf aglmcmc_pooled; do
  python bench.py --sampler $s --steps 2 --warmup 1 > $o/fin_bench_$s.json 2> $o/fin_bench_$s.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/fin_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $o/fin_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_isir -c 1 -o $o/k2_isir_full -f \
    python bench.py --sampler glmcmc --steps 1 --warmup 0 --no-e2e --no-cpu --no-extra > $o/fin_ncu_k2.log 2>&1
ls -la $o/fin_* | head -30
