# K1 / K2: CTA size vs the 6-vs-7 CTAs-per-SM tail
for s in global glmcmc; do
  for b in 32 64 128; do
    python bench.py --sampler $s --block $b --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$s', $b, d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
  done
done
