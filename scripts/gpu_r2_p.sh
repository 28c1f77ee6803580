#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-gpu --no-config5"
show() { python -c "import sys,json; d=json.loads(open('$1').read()); r=d['roofline']; print(d['value'], d['ms_per_step'], 'enc_ms', r['kernel_ms_per_step'], 'frac', r['frac'], 'sus', d['sustained']['value'], 'e2e', d['e2e']['value'], d['e2e_track_u8']['value'], d['latency']['infer_confidence_ms'])"; }
run() { tag=$1; shift; echo "$tag: $*"; env "$@" timeout 200 $B 2>/dev/null | tail -n 1 > gpurun_out/r2p_$tag.json; show gpurun_out/r2p_$tag.json; env "$@" LSD_TIMELINE=1 timeout 120 python scripts/run_forward_b64.py 2>&1 | grep timeline | tail -1; }
run base LSD_X=0
run hfearly LSD_HF_EARLY=1
run afterrows LSD_AUDIO_AFTER_ROWS=1
run afterrows_dyn LSD_AUDIO_AFTER_ROWS=1 LSD_UMMA_DYNAMIC=1
# latency probe inside a bench-like process: does an earlier score_batches call (pack threads + helper thread) change the graph replays?
timeout 200 python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import lipsync_b200 as lb, bench
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0), strict=True); m.to("cuda:0").eval(); m.compute_precision = "bf16"
p = lb.Predictor(m, batch_size=64)
print("fresh", bench.latency_probe(p, lb, 30)["infer_confidence_ms"], flush=True)
v, a = lb.synthetic_windows(3, 8)
vh = (torch.randint(0, 256, (8, 3, 32, 96, 96), dtype=torch.uint8).float() / 255.0).pin_memory()
p.score_batches((vh, a) for _ in range(4))
print("after score_batches", bench.latency_probe(p, lb, 30)["infer_confidence_ms"], p.graph_captures, flush=True)
vd, ad = lb.synthetic_windows(3, 64); vd, ad = vd.cuda(), ad.cuda()
for _ in range(200): m(vd, ad)
torch.cuda.synchronize()
print("after 200 B=64 forwards", bench.latency_probe(p, lb, 30)["infer_confidence_ms"], p.graph_captures, flush=True)
import time; time.sleep(2.0)
print("after 2 s idle", bench.latency_probe(p, lb, 30)["infer_confidence_ms"], p.graph_captures, flush=True)
PY
