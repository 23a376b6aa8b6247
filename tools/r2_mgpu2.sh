#!/bin/bash
# usage: r2_mgpu2.sh <tag> <n_gpus>  -- multi-GPU parity tests (pytest) + the driver's bench command at N GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; N=$2
export JB_MGPU_LOG=$O/${T}_mgpu_parity.log; rm -f $JB_MGPU_LOG
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/${T}_pytest_mgpu.log 2>&1; echo "pytest exit $?"; tail -5 $O/${T}_pytest_mgpu.log; cat $JB_MGPU_LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --cpu-seconds 0 > $O/${T}_bench_n$N.json 2> $O/${T}_bench_n$N.err; echo "bench exit $?"; tail -2 $O/${T}_bench_n$N.err
tail -1 $O/${T}_bench_n$N.json > $O/${T}_bench_n$N.line; python tools/bench_brief.py $O/${T}_bench_n$N.line
