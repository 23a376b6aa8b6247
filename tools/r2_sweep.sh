#!/bin/bash
# usage: r2_sweep.sh <tag> <n_gpus> "<people list>" [extra bench args]  -- bench lines over total populations (BASELINE.json configs[4])
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; N=$2; L=$3; shift; shift; shift
for P in $L; do
  if [ "$N" = "1" ]; then
    timeout 1200 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --people $P "$@" > $O/${T}_n${N}_$P.json 2> $O/${T}_n${N}_$P.err; rc=$?
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 --cpu-seconds 0 --people $P "$@" > $O/${T}_n${N}_$P.json 2> $O/${T}_n${N}_$P.err; rc=$?
  fi
  echo "== N=$N people=$P $* exit $rc"; [ $rc -ne 0 ] && tail -4 $O/${T}_n${N}_$P.err
  tail -1 $O/${T}_n${N}_$P.json > $O/${T}_n${N}_$P.line; python tools/bench_brief.py $O/${T}_n${N}_$P.line
done
