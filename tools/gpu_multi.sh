#!/bin/bash
# multi-GPU weak-scaling check (torchrun, NCCL stats all-gather) + new config tests
mkdir -p gpurun_out
G=${G:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $G --steps 1200 --warmup 20 --no-cpu --e2e-steps 20 > gpurun_out/bench_g$G.json 2> gpurun_out/bench_g$G.err; echo "bench x$G rc=$?"
cat gpurun_out/bench_g$G.json; tail -3 gpurun_out/bench_g$G.err
timeout 900 python -m pytest tests -m gpu -x -q -k "config2 or config5 or step_sample" 2>&1 | tail -3
