#!/bin/bash
# N-GPU session: the default bench line at N GPUs (C2 weak + exchange check + C4 strong + -bb on pattern shards)
N=${1:-2}
TAG=${2:-r02k}
mkdir -p gpurun_out/$TAG
cd "$(dirname "$0")/.."
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/$TAG/bench_n$N.json 2> gpurun_out/$TAG/bench_n$N.err ) 2>&1 | grep real
grep -v "^W\|^\[W\|^\*\|OMP_NUM" gpurun_out/$TAG/bench_n$N.err | tail -8
python - <<PY
import json
l=json.loads(open("gpurun_out/$TAG/bench_n$N.json").read().strip().splitlines()[-1])
print("N=%d value %.3e ms %.4f ins/s %.1fM e2e ms %.4f" % (l["n_gpus"], l["value"], l["ms_per_step"], l["insertions_per_s"]/1e6, l["e2e"]["ms_per_step"]))
print("exchange", json.dumps(l.get("exchange")))
print("c4", json.dumps(l.get("c4_strong"))[:600])
print("bb", json.dumps(l.get("bb"))[:900])
PY
