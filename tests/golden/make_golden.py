"""Generate golden vectors from the REAL reference (imported from /root/reference, this container only).

    python tests/golden/make_golden.py

The reference ships no tests, fixtures or known-answer vectors, so the oracle is pinned
against outputs of the reference itself run here on the seeded weights/inputs of `state_spec.py`.
Outputs (committed): tests/golden/model_golden.npz  — logits (full) + per-stage fingerprints
                     tests/golden/scoring_golden.json — Predictor helper known answers
Nothing at test time reads /root/reference; tests regenerate the same weights/inputs from the seeds.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import lipsync_b200 as lb  # noqa: E402
from tests.golden.fingerprint import fingerprint  # noqa: E402

CASES = {
    # name: (weight seed, rescale_head, input seed, B, T, H, W, F, Ta)
    "canonical": (0, True, 1, 4, 32, 96, 96, 80, 128),
    "canonical_unscaled": (0, False, 1, 4, 32, 96, 96, 80, 128),
    "half_window": (0, True, 2, 2, 16, 96, 96, 80, 64),      # _temporal_smoothed_confidence (predictor.py:295-331)
    "odd_shapes": (0, True, 3, 1, 8, 64, 64, 64, 100),       # audio tokens 13 -> lerp down to 8
    "single_frame": (0, True, 4, 1, 1, 96, 96, 80, 128),     # delta branch zeros (artifact_detector.py:168-171)
    "batch_tail": (0, True, 5, 3, 32, 96, 96, 80, 128),
}


def _hooks(model, store):
    names = {
        "visual_encoder.stem": "v_stem", "visual_encoder.layer1": "v_layer1", "visual_encoder.layer2": "v_layer2",
        "visual_encoder.layer3": "v_layer3", "visual_encoder.layer4": "v_layer4",
        "audio_encoder.stem": "a_stem", "audio_encoder.layer1": "a_layer1", "audio_encoder.layer2": "a_layer2",
        "audio_encoder.layer3": "a_layer3", "audio_encoder.layer4": "a_layer4",
        "cross_modal": "fused", "temporal": "cls", "artifact_detector": "artifact",
        "artifact_detector.high_freq_detector": "art_hf",
        "temporal.transformer.layers.0": "t_layer0", "temporal.transformer.layers.3": "t_layer3",
    }
    mods = dict(model.named_modules())
    hs = []
    for mname, key in names.items():
        def fn(_m, _i, out, key=key):
            store[key] = out.detach().clone()
        hs.append(mods[mname].register_forward_hook(fn))
    return hs


def main():
    from app.models.lip_sync_model import LipSyncModel

    out = {}
    torch.set_num_threads(os.cpu_count() or 1)
    for cname, (wseed, rescale, iseed, b, t, h, w, f, ta) in CASES.items():
        sd = lb.make_synthetic_state_dict(wseed, rescale_head=rescale)
        model = LipSyncModel().eval()
        model.load_state_dict(sd, strict=True)
        # slow path on purpose: hooks on the encoder layers need the python-level layer forward
        video, audio = lb.synthetic_windows(iseed, b, t, h, w, f, ta)
        store = {}
        hs = _hooks(model, store)
        with torch.no_grad():
            logits, aux = model(video, audio, return_aux=True)
        for hh in hs:
            hh.remove()
        with torch.inference_mode():
            logits_fast = model(video, audio)  # transformer fast path, as the Predictor runs it
        assert float((logits - logits_fast).abs().max()) < 1e-5
        out[f"{cname}/logits"] = logits_fast.numpy().astype(np.float32)
        store.update({"v_emb": aux["visual_tokens"], "a_emb": aux["audio_tokens"],
                      "fused": aux["fused_tokens"], "cls": aux["cls_output"]})
        for k, v in store.items():
            out[f"{cname}/fp/{k}"] = fingerprint(v)
        if cname == "canonical":
            out[f"{cname}/cls"] = aux["cls_output"].numpy().astype(np.float32)
        print(cname, logits_fast.numpy())
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **out)
    print("wrote model_golden.npz with", len(out), "arrays")

    # ---- Predictor helper known answers (a11-a15): import with a stub librosa (audio.py:6 imports it at top)
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    from app.inference.predictor import Predictor

    p = Predictor.__new__(Predictor)
    scoring = {"align": [], "robust": [], "weighted": [], "spans": []}
    rng = np.random.RandomState(0)
    for total_a, total_v, v_start in [(5335, 800, 0), (5335, 800, 8), (5335, 800, 768), (5335, 800, 792),
                                      (100, 40, 0), (100, 40, 8), (130, 33, 1), (128, 32, 0), (1000, 157, 120)]:
        full = np.arange(total_a, dtype=np.float32)[None, None, :].repeat(2, axis=1)
        chunk = p._align_audio_chunk(full, v_start, total_v)
        scoring["align"].append({"total_a": total_a, "total_v": total_v, "v_start": v_start,
                                 "cols": chunk[0, 0].astype(int).tolist()})
    for mode in ("none", "median", "trimmed_mean"):
        for n in (0, 1, 2, 5, 10, 23):
            confs = rng.rand(n).astype(np.float64).tolist()
            p.confidence_smoothing = mode
            p.trim_ratio = 0.1
            scoring["robust"].append({"mode": mode, "confs": confs, "out": p._robust_confidence(confs)})
    p.confidence_smoothing = "median"
    for n, with_vad in [(0, False), (4, False), (4, True), (11, True), (11, False)]:
        confs = rng.rand(n).tolist()
        speak = (rng.rand(n) * 1.4 - 0.2).tolist()
        vad = rng.rand(n).tolist() if with_vad else None
        scoring["weighted"].append({"confs": confs, "speak": speak, "vad": vad,
                                    "out": p._speech_weighted_confidence(confs, speak, vad_weights=vad)})
    # _temporal_smoothed_confidence window spans (predictor.py:302-325) with a stubbed _infer_confidence
    for t_v, t_a in [(32, 128), (16, 64), (24, 96), (32, 100), (12, 48), (40, 128)]:
        seen = []
        p._infer_confidence = lambda v, a, seen=seen: (seen.append((v.shape[1], a.shape[2])) or 0.5)
        _, confs, spans = p._temporal_smoothed_confidence(np.zeros((3, t_v, 4, 4), np.float32), np.zeros((1, 80, t_a), np.float32))
        scoring["spans"].append({"t_v": t_v, "t_a": t_a, "spans": [list(s) for s in spans], "shapes": [list(s) for s in seen]})
    with open(os.path.join(HERE, "scoring_golden.json"), "w") as fh:
        json.dump(scoring, fh)
    print("wrote scoring_golden.json")


if __name__ == "__main__":
    main()
