#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/r2_deepfm_time.py --extra --reps 5 > gpurun_out/r2_deepfm_time2.jsonl 2> gpurun_out/r2_deepfm_time2.err
echo "time rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_deepfm_time2.jsonl'):
    d=json.loads(l)
    if 'int32' in d['config'] or 'L2' in d['config']: continue
    print(d['config'][:42], d['kernel'], d['variant'], 'pdl' if d['pdl_overlap'] else 'nopdl', d['us_per_launch_median'])
PY
