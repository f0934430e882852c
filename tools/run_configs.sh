#!/bin/bash
# BASELINE.json configs[2..4] (and with `all` configs[1]) through bench.py at N GPUs of this box:
#   tools/run_configs.sh N [tag] [all]
# Writes gpurun_out/bench_<tag>_c{3,4,5}_<N>gpu.json (one JSON line each).
N=${1:-1}
TAG=${2:-r2}
C3="--hidden-dim 128 --processor-layers 8 --batch 8 --graph multiscale"
C4="--model hi_lam --graph hierarchical --ar-steps 3"
C5="--model hi_lam_parallel --graph hierarchical --hidden-dim 128 --processor-layers 2 --scale 2 --batch 2"
run() {
  name=$1; shift
  out=gpurun_out/bench_${TAG}_${name}_${N}gpu
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-fp32-line "$@" > $out.json 2> $out.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" \
      > $out.json 2> $out.err
  fi
  echo "$name N=$N rc=$? $(python -c "
import json,sys
try:
    d=json.load(open('$out.json')); print(round(d['value'],1),'samples/s', round(d['ms_per_step'],2),'ms/step e2e',round(d['e2e']['value'],1))
except Exception as e: print('no json', e)")"
}
if [ "$3" = "all" ]; then run c2; fi  # configs[1], the bench default
run c3 $C3
run c4 $C4
run c5 $C5
