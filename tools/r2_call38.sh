#!/bin/bash
for st in 2 3 4; do
  echo "== TRS_DENSE_A_STAGES=$st"
  TRS_DENSE_A_STAGES=$st timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>/dev/null | grep -o '"op": "[^"]*", "batch": [0-9]*, "us": [0-9.]*' | sed 's/"op": "\(.\{12\}\)[^"]*"/\1/' | head -1
done
