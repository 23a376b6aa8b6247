# usage: mgpu.sh <n_gpus> [ENV=VAL ...] -- 2-GPU correctness check + bench line, bounded by timeouts
n=$1; shift
env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -E "mgpu_check|Error|error|Traceback" | head -5
env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $n --steps 45 --warmup 9 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n GPUs $*', 'value', round(d['value']/1e9,2), 'G  ms/step', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))"
