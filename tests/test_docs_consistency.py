"""The README's table of environment knobs and the sources agree: every `LSD_*` variable the library reads is documented, and the
table lists nothing that no longer exists."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multimodal-lip-sync-deepfake-detection-system_b200")


def _sources() -> str:
    files = glob.glob(os.path.join(PKG, "csrc", "*.cu")) + glob.glob(os.path.join(PKG, "csrc", "*.cpp")) + \
        glob.glob(os.path.join(PKG, "csrc", "*.h")) + glob.glob(os.path.join(PKG, "csrc", "*.cuh")) + glob.glob(os.path.join(PKG, "*.py"))
    return "\n".join(open(f).read() for f in files)


def test_readme_knobs_match_sources():
    readme = open(os.path.join(ROOT, "README.md")).read()
    src = _sources()
    documented = set(re.findall(r"LSD_[A-Z0-9_]+", readme))
    read = set(re.findall(r'(?:getenv|LSD_ENV)\("(LSD_[A-Z0-9_]+)"\)', src))
    assert not sorted(read - documented), f"read by the library but not in README.md: {sorted(read - documented)}"
    stale = sorted(k for k in documented if k not in src)
    assert not stale, f"documented in README.md but gone from the sources: {stale}"


def test_header_cites_reference_lines():
    """include/lsd_b200.h names the reference interface (file:line) every entry point replaces."""
    hdr = open(os.path.join(ROOT, "include", "lsd_b200.h")).read()
    for needle in ("lip_sync_model.py", "predictor.py", "audio.py", "video.py"):
        assert needle in hdr, needle
    assert len(re.findall(r"\.py:\d+", hdr)) >= 8
