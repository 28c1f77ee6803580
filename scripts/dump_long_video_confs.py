"""B200: per-window confidences of the CUDA scorer on the seeded multi-track scenarios of tests/golden/long_video_synth.py.
    python scripts/dump_long_video_confs.py > gpurun_out/long_video_cuda_confs.json
Feed the result to tests/golden/make_long_video_golden.py (this container, real reference) to generate `cuda_scored/*` goldens."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
import lipsync_b200 as lb
from tests.golden.long_video_synth import SCENARIOS, make_audio, make_tracks

model = lb.LipSyncModel(); model.load_state_dict(lb.make_synthetic_state_dict(0), strict=True); model.to("cuda:0").eval()
model.compute_precision = "bf16"
pred = lb.Predictor(model, batch_size=16)
out = {}
for name in ("two_tracks_clear_winner", "three_tracks_turn_taking"):
    spec = SCENARIOS[name]
    tracks, n_frames = make_tracks(spec)
    mel, vad = make_audio(spec, n_frames)
    res = lb.predict_long_video_from_tracks(pred, tracks, mel, vad, spec["fps"], n_frames)
    out[name] = {str(t["track_id"]): t["window_confidences"] for t in res["tracks"]}
    print(name, res["verdict"], res["selected_track_id"], file=sys.stderr)
print(json.dumps(out))
