#!/bin/bash
N=$1; P=$2; shift 2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
i=0
for spec in "$@"; do
  i=$((i+1))
  $TR bench.py --gpus $N --steps 60 --warmup 5 $spec > ${P}_v$i.log 2> ${P}_v$i.err
  echo "v$i: $spec" >> ${P}_specs.txt
done
