#!/bin/bash
# N-GPU comparison of the exchange variants in one gpurun call: tools/scale_probe.sh N out_prefix
N=$1; P=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
$TR tools/exchange_bench.py > ${P}_ex.log 2>&1
$TR bench.py --gpus $N --steps 60 --warmup 5 --chunks 0 > ${P}_plain.log 2> ${P}_plain.err
$TR bench.py --gpus $N --steps 60 --warmup 5 --chunks 4 > ${P}_pipe.log 2> ${P}_pipe.err
$TR bench.py --gpus $N --steps 60 --warmup 5 --chunks 4 --exchange multimem > ${P}_mm.log 2> ${P}_mm.err
$TR bench.py --gpus $N --steps 60 --warmup 5 --chunks 0 --exchange multimem > ${P}_mm0.log 2> ${P}_mm0.err
