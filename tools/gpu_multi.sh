#!/bin/bash
# usage: gpu_multi.sh N [bench args...]   -- the bench over N GPUs of one box, as the driver launches it
N=$1; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
  > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
tail -5 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/bench_n{n}.json'))
except Exception as e:
    print("no line:", e); raise SystemExit
print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'scaling', 'gpu_launches', 'parity_check')})
print('e2e', d['e2e']['value'], d['e2e']['seconds'], 'capi', d['e2e_capi']['value'])
print('timing', {k: v for k, v in d['timing'].items() if k != 'how'})
print({k: v for k, v in d.items() if k.startswith(('sq_max', 'gr_slab', 'gr_in', 'classes', 'allreduce', 'parity'))})
print(d.get('clocks'))
PY
