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


def test_geometry_magic_division_is_exact(tmp_path):
    """The tcgen05 epilogue and the planar glue kernels decode flat padded positions by multiply + shift with per-geometry magic
    numbers (csrc/umma_conv.cuh: uc_magic; device side uc_div = (n * m) >> (31 + s)).  The host part is compiled with g++ and held
    to the integer division for every divisor up to 2^16 (and a sample of large ones) over boundary and pseudo-random numerators
    below 2^31 — the range the host enforces for a launch."""
    import os, shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    if shutil.which("g++") is None or not os.path.isdir(cuda_inc):
        pytest.skip("g++ / CUDA headers not available")
    src = tmp_path / "magic.cpp"
    src.write_text(r'''
#include <cstdint>
#include <cstdio>
#include "umma_conv.cuh"
static uint32_t uc_div(uint32_t n, uint32_t m, int s) { return (uint32_t)(((uint64_t)n * m) >> (31 + s)); }
int main() {
  uint64_t lcg = 12345;
  long bad = 0, checked = 0;
  auto check = [&](uint32_t d) {
    uint32_t m; int s;
    lsd::uc_magic(d, m, s);
    const uint32_t edge[] = {0u, 1u, d - 1, d, d + 1, 2 * d - 1, 2 * d, 0x7fffffffu, 0x7ffffffeu, 0x7fffffffu / d * d, 0x7fffffffu / d * d - 1};
    for (uint32_t n : edge) { if (n < 0x80000000u) { ++checked; bad += uc_div(n, m, s) != n / d; } }
    for (int i = 0; i < 64; ++i) {
      lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
      const uint32_t n = (uint32_t)(lcg >> 33);   // < 2^31
      ++checked; bad += uc_div(n, m, s) != n / d;
    }
  };
  for (uint32_t d = 1; d <= 65536; ++d) check(d);
  for (uint32_t d = 65537; d < 0x7fffffffu; d += 1000003u) check(d);
  // the geometry constructor fills the magic numbers of SL, RW and TS
  const lsd::UcGeom g = lsd::make_geom(64, 32, 48, 48);
  bad += uc_div(123456789u, g.mSL, g.sSL) != 123456789u / (uint32_t)g.SL;
  bad += uc_div(54321u, g.mRW, g.sRW) != 54321u / (uint32_t)g.RW;
  bad += uc_div(2111u, g.mTS, g.sTS) != 2111u / (uint32_t)g.TS;
  printf("checked %ld bad %ld\n", checked, bad);
  return bad != 0;
}
''')
    exe = tmp_path / "magic"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", cuda_inc, "-I", os.path.join(root, "multimodal-lip-sync-deepfake-detection-system_b200", "csrc"),
                           str(src), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
