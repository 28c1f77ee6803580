#!/bin/bash
cat > /tmp/dig.py <<'PY'
import sys; sys.path.insert(0, '/root/repo')
import torch, lipsync_b200 as lb
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to('cuda:0').eval(); m.compute_precision = 'bf16'
v, a = lb.synthetic_windows(33, 5)
out = m(v.cuda(), a.cuda()).float().cpu()
torch.cuda.synchronize()
print('LOGITS', [round(x, 6) for x in out.tolist()])
PY
echo "--- flat layer1"; timeout 120 python /tmp/dig.py 2>&1 | tail -1
echo "--- ring layer1"; LSD_L1_RING=1 timeout 120 python /tmp/dig.py 2>&1 | tail -3
echo "rc=$?"
LSD_L1_RING=1 LSD_CR_TRACE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep "\[cr\]" | tail -2
LSD_L1_RING=1 LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep -i "timeline" | tail -1 | cut -c1-220
LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep -i "timeline" | tail -1 | cut -c1-220
