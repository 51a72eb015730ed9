#!/bin/bash
# r02e GPU session (N GPUs of one box, default 2): the in-library exchange step over NVLink peer memory -- parity of sharded
# contexts against unsharded ones (peer and NCCL exchange), then the bench line at N GPUs (C2 weak + C4 strong + checks).
N=${1:-2}
mkdir -p gpurun_out/r02e
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -14
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02e/bench_n$N.json 2> gpurun_out/r02e/bench_n$N.err ) 2>&1 | grep real
tail -c 1500 gpurun_out/r02e/bench_n$N.err | grep -v "^W\|^\[W" | tail -8
python - <<PY
import json
l=json.loads(open("gpurun_out/r02e/bench_n$N.json").read().strip().splitlines()[-1])
print("N=%d value %.3e ms %.4f ins/s %.1fM e2e ms %.4f" % (l["n_gpus"], l["value"], l["ms_per_step"], l["insertions_per_s"]/1e6, l["e2e"]["ms_per_step"]))
print("exchange", json.dumps(l.get("exchange")))
print("c4", json.dumps(l.get("c4_strong"))[:900])
PY
