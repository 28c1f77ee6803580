"""Host-side logic against known answers produced by the real reference (tests/golden/make_golden.py), the log-mel
oracle against its independent fixture, and the world_size-2 window sharding over gloo.  CPU only."""
import json
import os
import socket

import numpy as np
import pytest
import torch

import lipsync_b200 as lb
from oracle import logmel_oracle as lmo
from tests.conftest import ROOT
from tests.golden.make_logmel_golden import CASES as MEL_CASES, make_pcm


@pytest.fixture(scope="module")
def scoring():
    with open(os.path.join(ROOT, "tests", "golden", "scoring_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture()
def pred():
    return lb.Predictor(model=None, device=torch.device("cpu"))


def test_align_audio_chunk(scoring, pred):
    assert scoring["align"]
    for case in scoring["align"]:
        full = np.arange(case["total_a"], dtype=np.float32)[None, None, :].repeat(2, axis=1)
        chunk = pred._align_audio_chunk(full, case["v_start"], case["total_v"])
        assert chunk.shape == (1, 2, 128)
        assert chunk[0, 0].astype(int).tolist() == case["cols"]
        # the device gather uses (a_start, clamp-to-last-column): same columns
        a0 = pred._audio_start(case["v_start"], case["total_a"], case["total_v"])
        cols = np.minimum(a0 + np.arange(128), case["total_a"] - 1)
        assert cols.tolist() == case["cols"]


def test_robust_confidence(scoring, pred):
    for case in scoring["robust"]:
        pred.confidence_smoothing = case["mode"]
        pred.trim_ratio = 0.1
        assert pred._robust_confidence(case["confs"]) == case["out"]


def test_speech_weighted_confidence(scoring, pred):
    for case in scoring["weighted"]:
        assert pred._speech_weighted_confidence(case["confs"], case["speak"], vad_weights=case["vad"]) == case["out"]


def test_temporal_smoothing_spans(scoring, pred):
    for case in scoring["spans"]:
        wins, spans = pred._smoothing_windows(case["t_v"], case["t_a"])
        assert [list(s) for s in spans] == case["spans"]
        assert [[v1 - v0, a1 - a0] for (v0, v1, a0, a1) in wins] == case["shapes"]


def test_calibration_matches_reference_formulas():
    p = lb.Predictor(model=None, device=torch.device("cpu"), calibration_method="temperature", calibration_temperature=2.0)
    assert p._calibrate(1.0) == float(torch.sigmoid(torch.tensor(0.5)).item())
    p = lb.Predictor(model=None, device=torch.device("cpu"), calibration_method="platt", calibration_platt_a=1.5, calibration_platt_b=-0.2)
    assert p._calibrate(0.4) == float(torch.sigmoid(torch.tensor(1.5 * 0.4 - 0.2)).item())
    p = lb.Predictor(model=None, device=torch.device("cpu"))
    assert p._calibrate(0.0) == 0.5


def test_partition_covers_everything():
    for n in (0, 1, 7, 10, 64, 10000):
        for ws in (1, 2, 3, 4, 8):
            spans = [lb.partition_windows(n, ws, r) for r in range(ws)]
            got = [i for lo, hi in spans for i in range(lo, hi)]
            assert got == list(range(n))
            assert max(hi - lo for lo, hi in spans) <= -(-n // ws) if n else True


def test_fit_frames():
    x = np.arange(2 * 3 * 5, dtype=np.float32).reshape(1, 6, 5)
    assert lb.fit_frames(x, None) is x
    assert lb.fit_frames(x, 3).shape == (1, 6, 3)
    y = lb.fit_frames(x, 8)
    assert y.shape == (1, 6, 8) and (y[:, :, 5:] == x[:, :, -1:]).all()


@pytest.mark.parametrize("name", list(MEL_CASES))
def test_logmel_oracle_matches_independent_fixture(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "logmel_golden.npz"))
    db = lmo.preprocess_audio_pcm(make_pcm(name))[0]
    assert db.shape == tuple(g[name + "/shape"])
    assert db.max() == 0.0 and db.min() >= -80.0
    assert np.abs(db.reshape(-1)[g[name + "/idx"]] - g[name + "/val"]).max() <= 1e-3  # dB


def test_logmel_filterbank_properties():
    fb = lmo.mel_filterbank()
    assert fb.shape == (80, 201) and fb.dtype == np.float32
    assert int((fb != 0).sum()) == 391
    assert (fb >= 0).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _shard_worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = lb.Predictor(model=None, device=torch.device("cpu"))
    # stub scorer: logit of window i is a deterministic function of i (the CUDA scorer is tested with -m gpu)
    out = p.score_windows_sharded(n, lambda lo, hi: torch.arange(lo, hi, dtype=torch.float32) * 0.5 - 3.0, world, rank)
    q.put((rank, out.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7, 1])
def test_sharded_scoring_gloo_world2(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    expect = (np.arange(n, dtype=np.float32) * 0.5 - 3.0).tolist()
    assert res[0] == expect and res[1] == expect


def test_mouth_motion_decision_and_aggregate_mirror_reference_golden():
    """`mouth_motion_checks` (predictor.py:403-419) and `aggregate_mouth_motion_checks` (:463-523) on the statistics the real
    reference computed (tests/golden/speech_golden.json): same per-window results, same aggregate."""
    import json
    import os
    import lipsync_b200 as lb
    with open(os.path.join(os.path.dirname(__file__), "golden", "speech_golden.json")) as fh:
        gold = json.load(fh)
    pred = lb.Predictor(None, batch_size=4)
    for name, g in gold.items():
        checks = pred.mouth_motion_checks([c["mouth_motion_energy"] for c in g["mouth"]], [c["audio_energy"] for c in g["mouth"]])
        assert [c["check_result"] for c in checks] == [c["check_result"] for c in g["mouth"]], name
        agg = lb.Predictor.aggregate_mouth_motion_checks(checks)
        assert agg["check_result"] == g["aggregate"]["check_result"], name
        assert agg["counts"] == g["aggregate"]["counts"] and agg["samples_checked"] == g["aggregate"]["samples_checked"]
        assert abs(agg["audio_energy"] - g["aggregate"]["audio_energy"]) < 1e-9
    assert lb.Predictor.aggregate_mouth_motion_checks([])["check_result"] == "no_data"


def test_u8_normalisation_formula_is_exact():
    """video_rows normalises uint8 pixels as q = i * (1/255); q += (i - 255 q) * (1/255) with fused multiply-adds
    (csrc/umma_conv.cu: vr_u8_norm).  That must equal the reference's astype(float32) / 255.0 (video.py:552-556) bit for bit
    for every byte value; the fused operations are emulated in float64, which is exact for these magnitudes."""
    import numpy as np
    i = np.arange(256, dtype=np.float32)
    ref = i / np.float32(255.0)
    c = np.float32(1.0) / np.float32(255.0)
    q = (i * c).astype(np.float32)
    r = (np.float64(-255.0) * q.astype(np.float64) + i.astype(np.float64)).astype(np.float32)          # fmaf(-255, q, i)
    out = (r.astype(np.float64) * np.float64(c) + q.astype(np.float64)).astype(np.float32)              # fmaf(r, c, q)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_host_pack_u8_exact():
    """`lsd_host_pack_u8_exact` (host threads, no GPU): packs fp32 pixels that are exactly uint8 / 255.0 (video.py:552-556) to
    bytes and refuses anything else — including -0.0, NaN, out-of-range and one-ulp-off values, at any position (vector body,
    scalar tail, any work item)."""
    from lipsync_b200 import _cabi
    L = _cabi.lib()
    rng = np.random.default_rng(0)
    for n, threads in ((0, 1), (1, 1), (31, 2), (256, 0), (1237, 3), (3 * (1 << 18) + 77, 4)):
        k = rng.integers(0, 256, n, dtype=np.uint8)
        if n >= 256:
            k[:256] = np.arange(256)
        x = k.astype(np.float32) / 255.0
        d = np.full(n, 7, np.uint8)
        assert L.lsd_host_pack_u8_exact(x.ctypes.data, d.ctypes.data, n, threads) == 1
        assert np.array_equal(d, k)
        if n == 0:
            continue
        for bad in (0.5, float(np.nextafter(np.float32(1 / 255), np.float32(1))), -1 / 255, 256 / 255, float("nan"), float("inf"), -0.0, 1e30, -1e30):
            for pos in {0, n // 2, n - 1}:
                y = x.copy()
                y[pos] = bad
                assert L.lsd_host_pack_u8_exact(y.ctypes.data, d.ctypes.data, n, threads) == 0, (n, bad, pos)
    assert L.lsd_host_pack_u8_exact(None, None, 5, 1) == _cabi.LSD_ERR_ARG
    # destination alignment selects the store flavour of the AVX-512 loop (64-byte aligned: non-temporal): both give the same bytes,
    # and a misaligned source is fine
    n = (1 << 18) + 200
    k = rng.integers(0, 256, n, dtype=np.uint8)
    xbuf = np.zeros(n + 16, np.float32)
    dbuf = np.zeros(n + 128, np.uint8)
    a0 = (-dbuf.ctypes.data) % 64
    for soff in (0, 3):
        xs = xbuf[soff:soff + n]
        xs[:] = k.astype(np.float32) / 255.0
        for doff in (a0, a0 + 1, a0 + 32):
            dbuf[:] = 9
            d = dbuf[doff:doff + n]
            assert L.lsd_host_pack_u8_exact(xs.ctypes.data, d.ctypes.data, n, 2) == 1
            assert np.array_equal(d, k) and (dbuf[:doff] == 9).all() and (dbuf[doff + n:] == 9).all()


def test_host_pack_u8_begin_end():
    """Two-step form of the pack (the scoring loop enqueues batch k while the host threads pack batch k+1): one job at a time,
    same results as the blocking call."""
    from lipsync_b200 import _cabi
    L = _cabi.lib()
    rng = np.random.default_rng(1)
    n = 2 * (1 << 18) + 13
    k = rng.integers(0, 256, n, dtype=np.uint8)
    x = k.astype(np.float32) / 255.0
    d = np.zeros(n, np.uint8)
    assert L.lsd_host_pack_u8_end() == 0                                               # nothing in flight
    # two jobs may be in flight (the scoring loop queues the pack of batch k+2 behind the one of batch k+1); _end returns them in order
    k2 = rng.integers(0, 256, n, dtype=np.uint8)
    x2 = k2.astype(np.float32) / 255.0
    x2[5] = 0.3                                                                        # the second job fails, the first does not
    d2 = np.zeros(n, np.uint8)
    assert L.lsd_host_pack_u8_begin(x.ctypes.data, d.ctypes.data, n, 3) == _cabi.LSD_OK
    assert L.lsd_host_pack_u8_begin(x2.ctypes.data, d2.ctypes.data, n, 2) == _cabi.LSD_OK
    assert L.lsd_host_pack_u8_begin(x.ctypes.data, d.ctypes.data, n, 3) == _cabi.LSD_ERR_ARG      # two in flight: busy
    assert L.lsd_host_pack_u8_exact(x.ctypes.data, d.ctypes.data, n, 3) == _cabi.LSD_ERR_ARG      # busy
    assert L.lsd_host_pack_u8_end() == 1 and np.array_equal(d, k)
    assert L.lsd_host_pack_last_ms() > 0.0
    assert L.lsd_host_pack_u8_end() == 0
    assert L.lsd_host_pack_u8_end() == 0                                               # nothing in flight
    for rep in range(20):                                                              # back-to-back queued jobs with changing thread counts
        ka = rng.integers(0, 256, n, dtype=np.uint8); kb = rng.integers(0, 256, n, dtype=np.uint8)
        xa = ka.astype(np.float32) / 255.0; xb = kb.astype(np.float32) / 255.0
        da = np.zeros(n, np.uint8); db = np.zeros(n, np.uint8)
        assert L.lsd_host_pack_u8_begin(xa.ctypes.data, da.ctypes.data, n, 1 + rep % 4) == _cabi.LSD_OK
        assert L.lsd_host_pack_u8_begin(xb.ctypes.data, db.ctypes.data, n, 1 + (rep * 3) % 5) == _cabi.LSD_OK
        assert L.lsd_host_pack_u8_end() == 1 and np.array_equal(da, ka)
        assert L.lsd_host_pack_u8_end() == 1 and np.array_equal(db, kb)
    x[n - 2] = 0.3
    assert L.lsd_host_pack_u8_begin(x.ctypes.data, d.ctypes.data, n, 0) == _cabi.LSD_OK and L.lsd_host_pack_u8_end() == 0
    assert L.lsd_host_pack_u8_begin(None, d.ctypes.data, n, 1) == _cabi.LSD_ERR_ARG
    assert L.lsd_host_pack_u8_exact(x.ctypes.data, d.ctypes.data, n, 2) == 0           # the pool is free again


def test_auto_transport_decision():
    """`host_transport="auto"` of `Predictor.score_batches` (the `_run_chunked_inference` contract, predictor.py:554-580): the
    decision to stop packing is a pure function of the measured pack times (measured cases: DESIGN.md §6/§7)."""
    from lipsync_b200.inference import auto_transport_gives_up_packing as gives_up
    t32 = 64 * 3 * 32 * 96 * 96 * 4 / 52e9                       # 4.36 ms: a 64-window fp32 batch over PCIe gen5 x16
    assert not gives_up([], [], t32)                             # nothing measured: keep packing
    assert not gives_up([2.2, 2.3, 2.5, 2.2], [2.2] * 5, t32)    # one GPU, 16 threads: packing wins
    assert not gives_up([3.3] * 6, [3.3] * 6, t32)               # two GPUs packing at once, 12 threads each: packing still wins
    assert gives_up([4.5, 4.6, 4.4, 4.5], [4.5] * 4, t32)        # marginal rule: the fastest of the last three is slower than the copy
    assert not gives_up([4.5, 4.6, 3.4, 4.5], [4.5] * 4, t32)    # ... one fast pack among them keeps the transport
    assert not gives_up([4.5, 4.6, 4.4], [4.5] * 3, t32)         # ... and it needs four measurements
    assert not gives_up([], [12.0], t32)                         # clearly-slower rule needs two packs in a row
    assert gives_up([], [12.0, 11.5], t32)                       # eight GPUs fed at once: decided inside a three-batch warm-up call
    assert not gives_up([], [12.0, 3.0], t32)                    # a single outlier does not flip the transport
