#!/bin/bash
# g(r) alone: timings of library variants / switches, then one ncu --set full capture of the pair kernel
mkdir -p gpurun_out
echo "default";  python tools/rdf_time.py 8 4
echo "L32 off"; MDC_RDF_L32=0 python tools/rdf_time.py 8 4
echo "REL off"; MDC_RDF_REL=0 python tools/rdf_time.py 8 4
echo "REL off, L32 off"; MDC_RDF_REL=0 MDC_RDF_L32=0 python tools/rdf_time.py 8 4
for v in mdcraft_b200/_lib/v_*.so; do
  [ -e $v ] || continue
  echo "$v";  MDCRAFT_B200_LIB=$v python tools/rdf_time.py 8 4
done
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:rdf_pairs_fast -s 1 -c 1 -f -o gpurun_out/rdf_rel python tools/rdf_time.py 2 2 > gpurun_out/ncu_rdf_rel.log 2>&1
echo "ncu exit $?"
