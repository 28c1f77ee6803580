"""Seeded synthetic multi-track long videos shared by the golden generator (real reference, this container) and the tests
(CPU replay and the B200 pipeline): uint8 mouth crops per track, absolute chunk starts, clip log-mel, VAD mask."""
import numpy as np


def _confs(seed, n, mean, spread, lo=None):
    r = np.random.RandomState(seed)
    c = np.clip(mean + spread * r.randn(n), 0.01, 0.99)
    if lo is not None:
        a, b, v = lo
        c[a:b] = v
    return [float(x) for x in c]


def _spec(seed, spans, stab, confs, fps=15.0, motion=(6.0, 6.0), mel=(0.0, 80.0)):
    """mel: the clip log-mel is `-(mel[0] + mel[1] * U(0,1))` dB."""
    return {"seed": seed, "spans": spans, "stability": stab, "confs": confs, "fps": fps, "motion": motion, "mel": mel}


def _n_chunks(span):
    return (span[1] - span[0] - 32) // 8 + 1


_S = {}
_sp = [(0, 136), (16, 120)]
_S["two_tracks_clear_winner"] = _spec(11, _sp, [0.9, 0.6], {"0": _confs(1, _n_chunks(_sp[0]), 0.85, 0.04), "1": _confs(2, _n_chunks(_sp[1]), 0.30, 0.05)})
_S["two_tracks_close_scores"] = _spec(12, _sp, [0.8, 0.8], {"0": _confs(3, _n_chunks(_sp[0]), 0.74, 0.01), "1": _confs(4, _n_chunks(_sp[1]), 0.72, 0.01)})
_S["two_tracks_conf_gap_small_selection_clear"] = _spec(13, _sp, [0.95, 0.30], {"0": _confs(5, _n_chunks(_sp[0]), 0.66, 0.01), "1": _confs(6, _n_chunks(_sp[1]), 0.62, 0.01)})
_sp3 = [(0, 200), (40, 176), (96, 200)]
_S["three_tracks_turn_taking"] = _spec(14, _sp3, [0.7, 0.5, 0.9], {"0": _confs(7, _n_chunks(_sp3[0]), 0.55, 0.15), "1": _confs(8, _n_chunks(_sp3[1]), 0.60, 0.15),
                                                                     "2": _confs(9, _n_chunks(_sp3[2]), 0.58, 0.15)})
_S["two_tracks_fake_winner"] = _spec(15, _sp, [0.9, 0.5], {"0": _confs(10, _n_chunks(_sp[0]), 0.12, 0.03), "1": _confs(11, _n_chunks(_sp[1]), 0.20, 0.03)})
_S["two_tracks_static_mouth"] = _spec(16, _sp, [0.9, 0.5], {"0": _confs(12, _n_chunks(_sp[0]), 0.40, 0.03), "1": _confs(13, _n_chunks(_sp[1]), 0.20, 0.03)}, motion=(0.0, 6.0))
_S["two_tracks_selection_uncertain"] = _spec(18, _sp, [0.80, 0.75], {"0": _confs(15, _n_chunks(_sp[0]), 0.74, 0.005), "1": _confs(16, _n_chunks(_sp[1]), 0.72, 0.005)})
_S["two_tracks_quiet_static_mouth"] = _spec(19, _sp, [0.9, 0.5], {"0": _confs(17, _n_chunks(_sp[0]), 0.40, 0.03), "1": _confs(18, _n_chunks(_sp[1]), 0.20, 0.03)},
                                            motion=(0.0, 6.0), mel=(55.0, 25.0))
_S["two_tracks_loud_static_mouth"] = _spec(20, _sp, [0.9, 0.5], {"0": _confs(19, _n_chunks(_sp[0]), 0.56, 0.02), "1": _confs(20, _n_chunks(_sp[1]), 0.20, 0.03)},
                                           motion=(0.0, 6.0), mel=(0.0, 20.0))
_sp1 = [(0, 360)]
_S["one_track_localised_fake"] = _spec(17, _sp1, [0.9], {"0": _confs(14, _n_chunks(_sp1[0]), 0.85, 0.04, lo=(20, 32, 0.05))})
SCENARIOS = _S


def make_tracks(spec):
    """-> (tracks, n_frames).  Crops: smooth random base image + per-frame random walk of amplitude `motion` (uint8)."""
    n_frames = max(b for _, b in spec["spans"])
    tracks = []
    for tid, ((a, b), stab) in enumerate(zip(spec["spans"], spec["stability"])):
        r = np.random.RandomState(1000 * spec["seed"] + tid)
        base = r.randint(40, 216, size=(1, 96, 96, 3)).astype(np.float32)
        walk = np.cumsum(spec["motion"][min(tid, len(spec["motion"]) - 1)] * r.randn(b - a, 96, 96, 3).astype(np.float32), axis=0)
        crops = np.clip(np.rint(base + walk), 0, 255).astype(np.uint8)
        starts = list(range(a, b - 32 + 1, 8))
        tracks.append({"track_id": tid, "crops_u8": crops, "chunk_starts": starts, "stability": float(stab), "hits": int(b - a),
                       "consecutive_miss_max": 0, "track_start_frame": int(a), "track_end_frame": int(b - 1)})
    return tracks, n_frames


def make_audio(spec, n_frames):
    r = np.random.RandomState(77 + spec["seed"])
    ta = int(n_frames / spec["fps"] * 100.0)
    mel = (-(spec["mel"][0] + spec["mel"][1] * r.rand(1, 80, ta))).astype(np.float32)
    vad = r.rand(ta) > 0.3
    return mel, vad
