#!/bin/bash
# sweep the CTA size of kron_rows_kernel on cfg 2 (tuning knob PETIGA_KRON_THREADS); prints threads, kernel ms, frac, SM MHz
for t in "$@"; do
  PETIGA_KRON_THREADS=$t python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-quad 2>/dev/null > /tmp/sweep.json
  python -c "import json; d=json.load(open('/tmp/sweep.json')); print($t, d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
