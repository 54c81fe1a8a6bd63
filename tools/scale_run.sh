#!/bin/bash
# Scaling record on one 8-GPU box: the reference arm, bench.py at N = 1/2/4/8 (weak) and N = 8 (strong), and the
# product CLI on the 64-clip batch at 1/2/4/8 GPUs. Outputs under gpurun_out/r03_scale_*.
o=gpurun_out
nproc > $o/r03_scale_host.log; nvidia-smi -L >> $o/r03_scale_host.log
python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > $o/r03_scale_ref.json 2> $o/r03_scale_ref.err
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $o/r03_scale_n1.json 2> $o/r03_scale_n1.err
port=29600
for n in 2 4 8; do
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 > $o/r03_scale_n$n.json 2> $o/r03_scale_n$n.err
done
port=$((port+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 10 --warmup 3 --scaling strong > $o/r03_scale_n8_strong.json 2> $o/r03_scale_n8_strong.err
python tools/batch64_cli.py --gpus 1,2,4,8 --threads 0 --chunk 10 --modes nopin > $o/r03_scale_batch_cli.log 2>&1
for f in $o/r03_scale_*.json; do echo "== $f"; tail -c 400 $f; echo; done
cat $o/r03_scale_batch_cli.log
