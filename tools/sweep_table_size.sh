#!/bin/bash
# kernel time of the DeepFM packed path against the table size (L2-resident ... DRAM-resident)
for r in 2560 25600 256000 5128192; do
  python bench.py --no-cpu-baseline --narrow-threads 0 --rows-per-field $r 2>/dev/null > /tmp/sweep.json
  python -c "import json; d=json.load(open('/tmp/sweep.json')); print('rows_per_field', $r, 'us_per_step', round(d['ms_per_step']*1e3, 2))"
done
