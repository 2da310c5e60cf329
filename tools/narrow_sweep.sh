#!/bin/bash
# e2e (pipelined session) against chunks per batch and host narrowing threads
for c in 1 2 4; do for t in 8 16; do
  python bench.py --no-cpu-baseline --steps 30 --narrow-threads $t --e2e-chunks $c 2>/dev/null > /tmp/ns.json
  python -c "import json; d=json.load(open('/tmp/ns.json'))['e2e']; print('chunks', $c, 'threads', $t, 'narrowing e2e', round(d['pipelined_host_narrowing_value']/1e6,1), 'plain e2e', round(d['pipelined_plain_value']/1e6,1), 'int32 e2e', round(d['int32_indices_value']/1e6,1), 'sync', round(d['sync_call_value']/1e6,1))"
done; done
