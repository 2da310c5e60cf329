#!/bin/bash
for blk in 0 128 112 176; do
  echo "== TRS_DENSE_BLOCK=$blk"
  TRS_DENSE_BLOCK=$blk timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>/dev/null | grep -o '"op": "[^"]*", "batch": [0-9]*, "us": [0-9.]*' | sed 's/"op": "\(.\{12\}\)[^"]*"/\1/'
done
TRS_DENSE_BLOCK=112 TRS_DENSE_TRACE=1 timeout 200 python tools/bench_ops.py --only deepfm_generic_mlp400 2>&1 | grep "dense trace" | awk '{k=$4" "$6" "$12" "$14; if (n[k]++ == 3) print}'
