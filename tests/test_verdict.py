"""Decision parity: the aggregation / gate mirror against decisions of the REAL reference `_predict_long_video`
(tests/golden/make_verdict_golden.py).  real / fake / uncertain must be identical.  CPU only."""
import json
import os

import pytest

import lipsync_b200 as lb
from tests.conftest import ROOT

with open(os.path.join(ROOT, "tests", "golden", "verdict_golden.json")) as _fh:
    GOLD = json.load(_fh)

EXACT = ["verdict", "is_real", "is_fake", "window_consensus_uncertain", "strict_fake_evidence", "sparse_real_guard_applied",
         "mouth_motion_override_applied", "override_reason", "temporal_confidence_drop"]
CLOSE = ["confidence", "window_weighted_confidence", "window_fake_vote_ratio", "temporal_drift", "first_half_avg_confidence",
         "second_half_avg_confidence"]


@pytest.mark.parametrize("name", sorted(GOLD))
def test_decision_matches_reference(name):
    g = GOLD[name]
    i = g["inputs"]
    st = dict(i["settings"])
    out = lb.aggregate_long_video(i["window_confs"], i["window_speaking"], i["window_vad"], i["mouth_check_result"], **st)
    for k in EXACT:
        assert out[k] == g[k], (name, k, out[k], g[k])
    for k in CLOSE:
        assert abs(out[k] - g[k]) <= 1e-6, (name, k, out[k], g[k])


def test_all_three_verdicts_are_covered():
    assert {g["verdict"] for g in GOLD.values()} == {"real", "fake", "uncertain"}
    assert {str(g["override_reason"]) for g in GOLD.values()} >= {"None", "window_consensus_mixed", "sparse_real_signal", "mouth_motion_uncertain"}


def test_select_windows_by_time():
    tracks = [{"track_id": 0, "stability": 0.9, "window_confidences": [0.2, 0.8, 0.5], "window_spans": [(0, 32), (8, 40), (16, 48)]},
              {"track_id": 1, "stability": 0.1, "window_confidences": [0.9, 0.1], "window_spans": [(8, 40), (24, 56)]}]
    out = lb.select_windows_by_time(tracks)
    assert [w["frame_start"] for w in out] == [0, 8, 16, 24]
    assert [w["selected_track_id"] for w in out] == [0, 0, 0, 1]   # 0.75*0.8+0.25*0.9 > 0.75*0.9+0.25*0.1
