# K3 tuning sweep (round 2): builds made with gl-abc-mcmc_b200/build.py build(variant=..., extra_flags=...)
for v in "" .c32 .c32k1 .k1 .c16; do
  for cfg in "32768 1001" "65536 1001" "262144 1001"; do
    set -- $cfg
    GLABC_LIB=$PWD/gl-abc-mcmc_b200/csrc/libglabc$v.so python bench.py --sampler glmala --chains $1 --iters $2 --layout none --steps 8 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', $1, d['value'], d['roofline']['kernel_ms'])"
  done
done
