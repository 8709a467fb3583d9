#!/bin/bash
# round-end evidence: plain bench, launch list, and one ncu --set full capture per dominant kernel
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --frames 1 --no-cpu"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:rdf_pairs_fast -s 1 -c 1 -f -o gpurun_out/prof_rdf_fast $B > gpurun_out/ncu_rdf.log 2>&1; echo "rdf exit $?"
ncu --set full --clock-control none --import-source on -k regex:sq_lattice_factored -s 1 -c 1 -f -o gpurun_out/prof_sq_factored $B > gpurun_out/ncu_sqf.log 2>&1; echo "sqf exit $?"
$B --sq-algo direct > gpurun_out/plain_direct.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sq_lattice_direct -s 1 -c 1 -f -o gpurun_out/prof_sq_direct $B --sq-algo direct > gpurun_out/ncu_sqd.log 2>&1; echo "sqd exit $?"
