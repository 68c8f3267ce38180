#!/bin/bash
# 2-GPU weak-scaling check of both bench arms, launched the way the driver does.
out=gpurun_out/${1:-scale2}
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/smi.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 ) > $out/bench_n2.json 2> $out/bench_n2.err
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > $out/bench_ref_n2.json 2> $out/bench_ref_n2.err
( timeout 300 python -m pytest tests/test_dist_cpu.py -q ) > $out/pytest_dist.log 2>&1
cat $out/bench_n2.json; tail -3 $out/bench_n2.err; cat $out/bench_ref_n2.json; tail -2 $out/pytest_dist.log
