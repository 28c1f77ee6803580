"""SM clock / power while the B=64 forward runs back to back for a few seconds (nvidia-smi sampled every 100 ms)."""
import sys, os, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lipsync_b200 as lb
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to('cuda').eval(); m.compute_precision = 'bf16'
v, a = lb.synthetic_windows(1, 4)
v = v.repeat(16, 1, 1, 1, 1).cuda(); a = a.repeat(16, 1, 1, 1).cuda()
for _ in range(5): m(v, a)
torch.cuda.synchronize()
samples, stop = [], False
def poll():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active,temperature.gpu", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.1)
t = threading.Thread(target=poll); t.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(1200):
    m(v, a)
    if i % 50 == 49: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
stop = True; t.join()
print(f"{e0.elapsed_time(e1) / 1200:.3f} ms per step over 1200 steps")
for s in samples[::3]: print(s)
