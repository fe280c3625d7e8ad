#!/bin/bash
# Scratch: C2 weak at N ranks under different NCCL / graph settings.  bash scripts/n8_variants.sh N
N=${1:-8}
run() {
  echo "== $*"
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 20 --warmup 5 --no-extra 2>/dev/null | tail -n 1 | \
      python -c "import json,sys; l=json.loads(sys.stdin.read()); print(round(l['value']/1e6,2), 'M stages/s', round(l['ms_per_step'],4), 'ms', 'e2e', round(l['e2e']['pinned']/1e6,2))"
}
run A=1
run HQPCU_GRAPHS=0
run NCCL_NVLS_ENABLE=0
run NCCL_ALGO=Ring NCCL_PROTO=LL
run NCCL_GRAPH_REGISTER=0
run NCCL_PROTO=LL128
