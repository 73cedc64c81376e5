# K3 CTA size: 64-thread blocks (512 CTAs at 32,768 chains = 3.46 per SM: a 4 : 3 imbalance) against 32-thread blocks (6.9 per SM)
for v in "" .b32; do
  for cfg in "32768 1001" "65536 1001" "262144 1001"; do
    set -- $cfg
    GLABC_LIB=$PWD/gl-abc-mcmc_b200/csrc/libglabc$v.so python bench.py --sampler glmala --chains $1 --iters $2 --layout none --steps 8 --warmup 3 --no-cpu --no-e2e --no-extra --no-other-configs --no-ref-python 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant[$v]', $1, d['value'], d['roofline']['kernel_ms'])"
  done
done
