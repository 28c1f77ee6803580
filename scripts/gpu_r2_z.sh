#!/bin/bash
cat > /tmp/dig.py <<'PY'
import sys; sys.path.insert(0, '/root/repo')
import torch, hashlib, lipsync_b200 as lb
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to('cuda:0').eval(); m.compute_precision = 'bf16'
v, a = lb.synthetic_windows(33, 5)
out = m(v.cuda(), a.cuda()).float().cpu()
torch.cuda.synchronize()
print('LOGITS', [round(x, 6) for x in out.tolist()])
PY
echo "--- ring off"; LSD_STEM_RING=0 timeout 120 python /tmp/dig.py 2>&1 | tail -2
echo "--- ring on";  timeout 120 python /tmp/dig.py 2>&1 | tail -4
echo "rc=$?"
