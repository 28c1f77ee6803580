#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-eager-gpu 2> gpurun_out/r02c_bench_n2.err | tail -n 1 > gpurun_out/r02c_bench_n2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c_bench_n2.json").read())
print(round(d["value"]), d["ms_per_step"], d["roofline"]["frac"], d["sustained"]["value"], d["e2e"], d["e2e_fp32_upload"]["value"], d["e2e_track_u8"]["value"], d.get("config5"))
PY
tail -3 gpurun_out/r02c_bench_n2.err
