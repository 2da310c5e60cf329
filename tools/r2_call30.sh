#!/bin/bash
mkdir -p gpurun_out
TRS_CIN_TRACE=1 timeout 120 python tools/cin_profile_driver.py 65536 2>&1 | grep -E "cin trace|Msamp" | sort | uniq -c | sort -rn | awk '{$1=""; print}' | awk '{k=$4" "$6; if (n[k]++ < 2) print}'
timeout 120 python tools/cin_profile_driver.py 65536 2>&1 | tail -1
