#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --no-eager-gpu 2> gpurun_out/r02e_bench_n$N.err | tail -n 1 > gpurun_out/r02e_bench_n$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/r02e_bench_n$N.json").read())
print("N=$N value", round(d["value"]), d["ms_per_step"], "frac", round(d["roofline"]["frac"],3), "sustained", round(d["sustained"]["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["transport"], "fp32", round(d["e2e_fp32_upload"]["value"]), "track_u8", round(d["e2e_track_u8"]["value"]), "config5", d.get("config5"))
PY
tail -2 gpurun_out/r02e_bench_n$N.err
