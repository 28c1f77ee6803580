"""C-ABI boundary on a CPU box: the library builds, loads, and exports every symbol include/lsd_b200.h declares;
the product path fails loudly (no fallback) without an sm_100 device.  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

from tests.conftest import ROOT


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from lipsync_b200 import _cabi
    return _cabi


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "lsd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lsd_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = built.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/lsd_b200.h but not exported"
    assert declared == set(built.EXPORTS)


def test_version_and_pure_helpers(built):
    L = built.lib()
    assert L.lsd_version() >= 100
    assert L.lsd_audio_tokens(128) == 16 and L.lsd_audio_tokens(64) == 8 and L.lsd_audio_tokens(100) == 13
    assert L.lsd_logmel_frames(20480) == 129 and L.lsd_logmel_frames(0) == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-box behaviour")
def test_create_fails_loudly_without_gpu(built):
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback|sm_"):
        built.Handle(0)


def test_model_refuses_cpu():
    import lipsync_b200 as lb
    m = lb.LipSyncModel()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 8, 96, 96), torch.zeros(1, 1, 80, 128))
    with pytest.raises(ValueError):
        m(torch.zeros(3, 8, 96, 96), torch.zeros(1, 1, 80, 128))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 8, 96, 96), torch.zeros(1, 80, 128))
    with pytest.raises(NotImplementedError):
        lb.LipSyncModel(detect_artifacts=False)


def test_state_dict_boundary(seed0_sd):
    import lipsync_b200 as lb
    m = lb.LipSyncModel()
    assert list(sorted(m.state_dict().keys())) == sorted(seed0_sd.keys())
    m.load_state_dict(seed0_sd, strict=True)
    bad = dict(seed0_sd)
    bad.pop("classifier.net.4.bias")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)
    bad = dict(seed0_sd)
    bad["classifier.net.4.weight"] = torch.zeros(2, 128)
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)
    # checkpoint wrappers are unwrapped by the caller exactly as predictor.py:188-192 does
    m.half()
    assert m.state_dict()["classifier.net.4.weight"].dtype == torch.float16
    assert m.state_dict()["visual_encoder.stem.1.num_batches_tracked"].dtype == torch.int64


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multimodal-lip-sync-deepfake-detection-system_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"
