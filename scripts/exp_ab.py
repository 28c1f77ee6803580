"""A/B of scheduling knobs at B=64: ms per forward over back-to-back forwards (burst: 40 steps; sustained: 400 steps), fresh process per variant."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys; sys.path.insert(0, %r)
import torch, lipsync_b200 as lb
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to('cuda').eval(); m.compute_precision = 'bf16'
v, a = lb.synthetic_windows(1, 4)
v = v.repeat(16, 1, 1, 1, 1).cuda(); a = a.repeat(16, 1, 1, 1).cuda()
for _ in range(5): m(v, a)
torch.cuda.synchronize()
def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): m(v, a)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
b = [run(40) for _ in range(3)]
s = run(400)
print("RES burst %%.4f %%.4f %%.4f sustained %%.4f" %% (b[0], b[1], b[2], s))
''' % ROOT
variants = [a.split(",") if a else [] for a in sys.argv[1:]] or [[]]
for rep in range(2):
    for var in variants:
        env = dict(os.environ)
        for kv in var:
            if kv:
                k, v = kv.split("=")
                env[k] = v
        r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith("RES")]
        print(",".join(var) or "default", line[-1] if line else r.stderr[-300:], flush=True)
