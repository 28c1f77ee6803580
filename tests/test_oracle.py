"""Pin the oracle (oracle/lipsync_oracle.py) against golden vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import lipsync_b200 as lb
from oracle import lipsync_oracle as orc
from tests.golden.fingerprint import fingerprint
from tests.golden.make_golden import CASES

LOGIT_TOL = 2e-5  # fp32 CPU, different op order (manual MHA / LN vs fused torch modules)


def test_state_spec_has_270_entries():
    spec = lb.state_spec()
    assert len(spec) == 270
    n_params = sum(int(np.prod(s)) for k, s in spec.items() if not k.endswith(lb.state_spec.__globals__["BUFFER_SUFFIXES"]))
    assert n_params == 16_248_275


@pytest.mark.parametrize("case", ["canonical", "half_window", "odd_shapes", "single_frame"])
def test_oracle_matches_reference_golden(golden, case):
    wseed, rescale, iseed, b, t, h, w, f, ta = CASES[case]
    sd = lb.make_synthetic_state_dict(wseed, rescale_head=rescale)
    video, audio = lb.synthetic_windows(iseed, b, t, h, w, f, ta)
    inter = {}
    logits = orc.forward(sd, video, audio, inter=inter)
    ref = golden[f"{case}/logits"]
    assert logits.shape == (b,)
    assert np.abs(logits.numpy() - ref).max() <= LOGIT_TOL
    for key in golden.files:
        if not key.startswith(f"{case}/fp/"):
            continue
        name = key.split("/")[-1]
        got = fingerprint(inter[name])
        exp = golden[key]
        assert got[0] == exp[0], name
        scale = max(1.0, float(np.abs(exp[3:]).max()))
        assert np.abs(got[3:] - exp[3:]).max() <= 5e-5 * scale, name
        assert abs(got[2] - exp[2]) <= 1e-5 * max(1.0, exp[2]), name


def test_oracle_unscaled_head(golden):
    wseed, rescale, iseed, b, t, h, w, f, ta = CASES["canonical_unscaled"]
    sd = lb.make_synthetic_state_dict(wseed, rescale_head=False)
    video, audio = lb.synthetic_windows(iseed, b, t, h, w, f, ta)
    logits = orc.forward(sd, video, audio)
    assert np.abs(logits.numpy() - golden["canonical_unscaled/logits"]).max() <= LOGIT_TOL


def test_oracle_batch_independence():
    """Per-window results do not depend on batch composition (SURVEY.md §7.3 determinism)."""
    sd = lb.make_synthetic_state_dict(0)
    video, audio = lb.synthetic_windows(5, 3)
    full = orc.forward(sd, video, audio)
    one = orc.forward(sd, video[1:2], audio[1:2])
    assert abs(float(full[1] - one[0])) < 1e-5


def test_oracle_shape_errors():
    sd = lb.make_synthetic_state_dict(0)
    with pytest.raises(ValueError):
        orc.visual_encoder(sd, torch.zeros(3, 8, 96, 96))
    with pytest.raises(ValueError):
        orc.audio_encoder(sd, torch.zeros(1, 80, 128))


def test_lerp_matches_documented_weights():
    a = torch.arange(16, dtype=torch.float32).view(1, 16, 1)
    out = orc.lerp_tokens(a, 32)[0, :, 0]
    assert out[0] == 0 and out[31] == 15
    assert abs(float(out[1]) - 0.25) < 1e-6 and abs(float(out[2]) - 0.75) < 1e-6
