#!/bin/bash
# ncu --set full of two kernels added in round 1g (one launch each), raw pages exported as CSV on the box
mkdir -p gpurun_out
timeout 120 ncu --set full --clock-control none --import-source on -k regex:tall_dense_tc -s 3 -c 1 -f -o gpurun_out/r1g_tall_dense_tc python tools/bench_ops.py --only tall_mlp --rows-per-field 65536 > gpurun_out/r1g_ncu_tall.log 2>&1
echo "tall rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:afm_backward -s 3 -c 1 -f -o gpurun_out/r1g_afm_backward python tools/bench_ops.py --only afm_backward --rows-per-field 65536 > gpurun_out/r1g_ncu_afm_bwd.log 2>&1
echo "afm rc=$?"
for n in r1g_tall_dense_tc r1g_afm_backward; do timeout 60 ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep | tail -3
