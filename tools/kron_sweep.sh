#!/bin/bash
# sweep the CTA size of kron_rows_kernel (tuning knob PETIGA_KRON_THREADS): tools/kron_sweep.sh MESH T1 T2 ...
# prints threads, ms per step (async loop), frac of HBM copy peak, SM MHz
mesh=$1; shift
for t in "$@"; do
  PETIGA_KRON_THREADS=$t python bench.py --mesh $mesh --steps 200 --warmup 3 --no-e2e --no-cpu-baseline --no-quad 2>/dev/null > /tmp/sweep.json
  python -c "import json; d=json.load(open('/tmp/sweep.json')); print($t, d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
