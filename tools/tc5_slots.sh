#!/bin/bash
# cross_tc5.cu: tiles in flight per CTA (slots) -- tcgen05 chain against the mma.sync chain, same inputs (bench_ops)
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader
for n in 1 2 3 4 5; do
  echo "slots $n"
  TRS_TC5_SLOTS=$n python tools/bench_ops.py --only cross_layer 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-62s %8.1f us  %6.1f M samples/s' % (d['op'][:62], d['us'], d['samples_per_s'] / 1e6))"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader
