#!/bin/bash
# usage: tools/scaling_sweep.sh N  -- train (weak, strong), ensemble at N GPUs of one box; one JSON line each
N=$1; O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
if [ "$N" = "1" ]; then T="python"; fi
timeout 150 $T bench.py --gpus $N --steps 12 --warmup 4 --no-aux > $O/r02_train_weak_n$N.json 2> $O/r02_sweep_n$N.err
timeout 150 $T bench.py --gpus $N --steps 12 --warmup 4 --no-aux --strong > $O/r02_train_strong_n$N.json 2>> $O/r02_sweep_n$N.err
timeout 150 $T bench.py --gpus $N --steps 12 --warmup 4 --no-aux --strong --graph > $O/r02_train_strong_graph_n$N.json 2>> $O/r02_sweep_n$N.err
timeout 150 $T bench.py --gpus $N --workload ensemble --steps 4 --warmup 3 > $O/r02_ensemble_n$N.json 2>> $O/r02_sweep_n$N.err
for f in $O/r02_train_weak_n$N.json $O/r02_train_strong_n$N.json $O/r02_train_strong_graph_n$N.json $O/r02_ensemble_n$N.json; do
  python -c "import sys,json; L=[l for l in open(sys.argv[1]) if l.startswith(chr(123))]; d=json.loads(L[-1]) if L else {}; print(sys.argv[1], d.get('value'), d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('scaling'))" $f
done
tail -c 300 $O/r02_sweep_n$N.err
