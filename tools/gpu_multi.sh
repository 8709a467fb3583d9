#!/bin/bash
# multi-GPU check: smoke on rank 0, then the bench under torchrun with $1 ranks
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
tail -c 1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
