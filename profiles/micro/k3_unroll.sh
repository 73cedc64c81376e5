# K3 gradient-loop unroll 1 (default) against 2 with one-warp CTAs
for v in "" .u2; do
  for cfg in "32768 1001" "65536 1001" "262144 1001"; do
    set -- $cfg
    GLABC_LIB=$PWD/gl-abc-mcmc_b200/csrc/libglabc$v.so python bench.py --sampler glmala --chains $1 --iters $2 --layout none --steps 8 --warmup 3 --no-cpu --no-e2e --no-extra --no-other-configs --no-ref-python 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant[$v]', $1, d['value'], d['roofline']['kernel_ms'])"
  done
done
