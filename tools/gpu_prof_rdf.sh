#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"rdf_pairs_fast" -c 2 -f -o gpurun_out/rdf_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_rdf.log 2>&1
tail -2 gpurun_out/ncu_rdf.log
