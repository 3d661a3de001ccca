#!/bin/bash
N=$1; P=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
for d in 1 2; do
$TR bench.py --gpus $N --steps 60 --warmup 5 --chunks 4 --exchange multimem --deferred-views $d > ${P}_mm_d$d.log 2> ${P}_mm_d$d.err
done
$TR bench.py --gpus $N --steps 60 --warmup 5 --chunks 2 --exchange multimem --deferred-views 1 > ${P}_mm_d1c2.log 2> ${P}_mm_d1c2.err
