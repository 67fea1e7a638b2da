#!/bin/bash
# multi-GPU: the driver's launch line (torchrun, one rank per GPU), ours + reference arm
mkdir -p gpurun_out
G=${G:-2}
T=${TAG:-r02m}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/${T}_bench_g$G.json 2> gpurun_out/${T}_bench_g$G.err; echo "bench x$G rc=$?"
cat gpurun_out/${T}_bench_g$G.json; tail -5 gpurun_out/${T}_bench_g$G.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $G --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference_g$G.json 2>> gpurun_out/${T}_bench_g$G.err; echo "ref x$G rc=$?"
cat gpurun_out/${T}_bench_reference_g$G.json
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"; cat /sys/fs/cgroup/cpu.max 2>/dev/null
