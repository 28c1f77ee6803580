#!/bin/bash
# diagnostic: which of the untested changes hangs the bf16 forward (every step under a short timeout)
mkdir -p gpurun_out
L=gpurun_out/r2e_diag.log; : > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "rc=$?" >> $L; }
cat > /tmp/fwd.py <<'PY'
import sys, os
sys.path.insert(0, '/root/repo')
import torch
import lipsync_b200 as lb
B = int(sys.argv[1])
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to('cuda').eval(); m.compute_precision='bf16'
v, a = lb.synthetic_windows(1, min(B, 4))
v = v.repeat((B + 3) // 4, 1, 1, 1, 1)[:B].cuda(); a = a.repeat((B + 3) // 4, 1, 1, 1)[:B].cuda()
print("start", flush=True)
out = m(v, a); torch.cuda.synchronize()
print("logits", out[:4].tolist(), flush=True)
PY
cat > /tmp/tok.py <<'PY'
import sys, os
sys.path.insert(0, '/root/repo')
import torch
import lipsync_b200 as lb
B = int(sys.argv[1])
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to('cuda').eval(); m.compute_precision='bf16'
g = torch.Generator().manual_seed(4)
v = torch.randn(B, 32, 256, generator=g).cuda(); a = torch.randn(B, 16, 256, generator=g).cuda()
print("start", flush=True)
f, c = m.fuse_tokens(v, a); torch.cuda.synchronize()
print("cls", c[0, :4].tolist(), flush=True)
PY
LSD_TOK_FUSED=0 LSD_STEM_CHUNK=0 run timeout 90 python -u /tmp/fwd.py 4
LSD_TOK_FUSED=0 LSD_STEM_CHUNK=0 run timeout 60 python -u /tmp/tok.py 4
LSD_STEM_CHUNK=0 run timeout 60 python -u /tmp/tok.py 1
LSD_STEM_CHUNK=0 run timeout 60 python -u /tmp/tok.py 4
LSD_TOK_FUSED=0 LSD_STEM_CHUNK=7 run timeout 60 python -u /tmp/fwd.py 4
LSD_TOK_FUSED=0 LSD_STEM_CHUNK=7 run timeout 60 python -u /tmp/fwd.py 16
LSD_TOK_FUSED=0 LSD_STEM_CHUNK=7 run timeout 60 python -u /tmp/fwd.py 64
# host pack rate + u8 transport (safe configuration)
export LSD_TOK_FUSED=0 LSD_STEM_CHUNK=0
run timeout 120 python -u - <<'PY'
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, os
from lipsync_b200 import _cabi
L = _cabi.lib()
print("cpus", os.cpu_count())
n = 64 * 3 * 32 * 96 * 96
x = (torch.randint(0, 256, (n,), dtype=torch.uint8).float() / 255.0).pin_memory()
d = torch.empty(n, dtype=torch.uint8).pin_memory()
for th in (1, 2, 4, 8, 16, 32):
    L.lsd_host_pack_u8_exact(x.data_ptr(), d.data_ptr(), n, th)
    t = time.perf_counter()
    for _ in range(5): r = L.lsd_host_pack_u8_exact(x.data_ptr(), d.data_ptr(), n, th)
    dt = (time.perf_counter() - t) / 5
    print(f"pack threads={th} ok={r} {dt*1e3:.2f} ms {n*4/dt/1e9:.1f} GB/s", flush=True)
PY
run timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "u8_transport or cuda_graph"
run timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5
tail -c 6000 $L
