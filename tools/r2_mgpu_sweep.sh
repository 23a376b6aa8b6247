#!/bin/bash
# usage: r2_mgpu_sweep.sh <tag> <n_gpus> "<people list>" [parity] [config4]  -- N-GPU bench lines (+ parity tests, + config4 at the first size)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; N=$2; L=$3
if [ "$4" = "parity" ]; then
  export JB_MGPU_LOG=$O/${T}_mgpu_parity.log; rm -f $JB_MGPU_LOG
  timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/${T}_pytest_mgpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/${T}_pytest_mgpu.log; cat $JB_MGPU_LOG
fi
bash tools/r2_sweep.sh $T $N "$L"
if [ "$5" = "config4" ]; then
  bash tools/r2_sweep.sh ${T}_c4 $N "$(echo $L | cut -d' ' -f1)" --config4
fi
