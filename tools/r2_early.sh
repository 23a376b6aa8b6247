#!/bin/bash
# usage: r2_early.sh <tag> -- parity tests, then the step timeline (tools/stamps.py) with and without the early infection consumers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1
timeout 900 python -m pytest tests/test_gpu_vs_oracle.py tests/test_full_size.py tests/test_golden_parity.py tests/test_variants.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${T}_pytest.log; tail -5 $O/${T}_pytest.log
for P in 56000000 14000000; do
  for E in 1 0; do
    echo "== people $P early $E"
    JB_NEWINF_EARLY=$E JB_STAMPS=1 timeout 600 python tools/stamps.py $P 2>&1 | tail -14
  done
done
