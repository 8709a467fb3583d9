#!/bin/bash
# kernel experiments: tests (main lib) + bench with alternative builds of the library
python -m pytest tests -q -m gpu -k rdf 2>&1 | tail -3
for L in "$@"; do
  echo "== $L"
  MDCRAFT_B200_LIB=$PWD/mdcraft_b200/_lib/$L python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:round(d[k],3) for k in ('value','sq_frames_per_s','gr_frames_per_s')}, round(d['roofline']['frac'],3))"
done
