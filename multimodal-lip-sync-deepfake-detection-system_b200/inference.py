"""Window-scoring API: the scoring/aggregation subset of the reference `Predictor`
(app/inference/predictor.py), re-hosted on the batched B200 forward.

Mirrors, with the reference's names, argument meaning and return values:
  `_infer_confidence`              predictor.py:212-244   (+ batched `_infer_confidences`)
  `_robust_confidence`             predictor.py:246-260
  `_speech_weighted_confidence`    predictor.py:262-293
  `_temporal_smoothed_confidence`  predictor.py:295-331
  `_align_audio_chunk`             predictor.py:525-552
  `_run_chunked_inference`         predictor.py:554-580   (serial B=1 loop -> batches of `batch_size` windows)
New: `score_track` (uint8 track -> device-built windows, C `lsd_score_windows`) and
`score_windows_sharded` (contiguous block partition over ranks + one all-gather of fp32 logits).
Video decode, face tracking, VAD and the gate/verdict block of `_predict_long_video` stay reference Python.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from . import aggregate as _agg
from .model import LipSyncModel


def partition_windows(n_windows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of [0, n_windows): rank r owns [lo, hi) with ceil(n/R) windows per rank
    (contiguous so each rank reads one contiguous span of the track)."""
    per = -(-n_windows // world_size) if world_size > 0 else n_windows
    lo = min(n_windows, rank * per)
    hi = min(n_windows, lo + per)
    return lo, hi


def gather_logits(local: torch.Tensor, n_windows: int, world_size: int, rank: int) -> torch.Tensor:
    """All-gather per-rank fp32 logits (padded to ceil(n/R)) and strip the padding: the only collective on the path."""
    import torch.distributed as dist

    per = -(-n_windows // world_size)
    buf = torch.zeros(per, dtype=torch.float32, device=local.device)
    buf[: local.numel()] = local.to(torch.float32)
    if world_size == 1:
        return buf[:n_windows]
    out = torch.empty(world_size * per, dtype=torch.float32, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, buf)  # NCCL over NVLink: <= 4 B per window
    else:
        dist.all_gather(list(out.view(world_size, per).unbind(0)), buf)  # gloo (CPU tests of the host logic)
    return out[:n_windows]


def auto_transport_gives_up_packing(steady_ms: Sequence[float], all_ms: Sequence[float], t32_s: float) -> bool:
    """The decision of `host_transport="auto"` (a pure function of the measured pack times, unit-tested on CPU): stop packing when
    packing is the slower way to feed the GPU.  `steady_ms`: whole-batch pack times in ms, oldest first, without the first batch of a
    call and without first packs into freshly pinned staging buffers; `all_ms`: every pack but the very first of the predictor;
    `t32_s`: the fp32 bytes of a batch at ~52 GB/s of PCIe gen5 x16, in seconds.
      * marginal: the fastest of the last three steady packs (once four are measured) + 0.1 ms is slower than the fp32 copy — a slow
        pack or two on a busy host must not flip the transport for good;
      * clearly slower: the last two packs of any kind both took more than 1.5 x the fp32 copy (every GPU of a box fed at once: the
        host DRAM is the limit and the pack moves more of it than the copy it saves — 12 ms against 4.4 ms with eight GPUs) — decided
        within a three-batch warm-up call instead of five batches into the next one."""
    if len(steady_ms) >= 4 and min(steady_ms[-3:]) * 1e-3 + 0.1e-3 > t32_s:
        return True
    return len(all_ms) >= 2 and min(all_ms[-2:]) * 1e-3 > 1.5 * t32_s


class Predictor:
    """Scoring half of the reference Predictor.  `model` is a `lipsync_b200.LipSyncModel` already on a CUDA device."""

    def __init__(
        self,
        model: LipSyncModel,
        device: Optional[torch.device] = None,
        use_half_precision: bool = False,
        confidence_smoothing: str = "median",
        trim_ratio: float = 0.1,
        calibration_method: str = "none",
        calibration_temperature: float = 1.0,
        calibration_platt_a: float = 1.0,
        calibration_platt_b: float = 0.0,
        isotonic_calibrator=None,
        chunk_size: int = 32,
        chunk_stride: int = 8,
        batch_size: int = 64,
        use_cuda_graphs: bool = True,
        graph_max_batch: int = 8,
        host_transport: str = "auto",
        host_pack_threads: Optional[int] = None,
        host_split: bool = False,
    ) -> None:
        self.model = model
        self.device = device if device is not None else (model._device() if model is not None else torch.device("cpu"))
        self.use_half_precision = bool(use_half_precision and self.device.type == "cuda")
        allowed = {"none", "median", "trimmed_mean"}
        self.confidence_smoothing = confidence_smoothing if confidence_smoothing in allowed else "median"
        self.trim_ratio = float(min(max(trim_ratio, 0.0), 0.49))
        _cal_allowed = {"none", "temperature", "platt", "isotonic"}
        self._cal_method = calibration_method if calibration_method in _cal_allowed else "none"
        self._cal_temperature = float(max(1e-3, calibration_temperature))
        self._cal_platt_a = float(calibration_platt_a)
        self._cal_platt_b = float(calibration_platt_b)
        self._isotonic_cal = isotonic_calibrator
        if self._cal_method == "isotonic" and self._isotonic_cal is None:
            self._cal_method = "none"
        self.chunk_size = int(chunk_size)
        self.chunk_stride = int(chunk_stride)
        self.batch_size = int(max(1, batch_size))
        #: small host batches (the B=1 calls of `_infer_confidence`, the 1+3 windows of `_temporal_smoothed_confidence`) replay a
        #: captured CUDA graph (H2D copies + the ~60 launches of the forward + D2H) instead of enqueueing launch by launch
        self.use_cuda_graphs = bool(use_cuda_graphs)
        self.graph_max_batch = int(graph_max_batch)
        self._graphs = {}
        self.graph_captures = 0      # how many graphs this predictor captured (a rising count under steady shapes = re-capture churn)
        #: how fp32 host windows cross PCIe in `score_batches`: "u8" packs windows whose pixels are exactly k/255 (what
        #: video.py:552-556 produces) to one byte per pixel on the host threads (`lsd_host_pack_u8_exact`: verified value by value,
        #: logits bit-identical), "fp32" copies them as they are, "auto" packs when the first batch qualifies and keeps packing
        #: unless the pack turns out slower than PCIe would move the fp32 bytes (fastest of the last three packs once four are measured, remembered across calls, see finish())
        if host_transport not in ("auto", "u8", "fp32"):
            raise ValueError(f"host_transport must be 'auto', 'u8' or 'fp32', got {host_transport!r}")
        self.host_transport = host_transport
        #: split transport of `score_batches` (opt-in): part of every batch crosses PCIe as fp32 while the host threads pack the rest
        #: (True: the share follows the measured pack rate; a float in (0, 1) fixes it; logits unchanged, `lsd_expand_u8`).  Off by
        #: default: measured slower than packing whole batches on one GPU (17.1 k against 22 k windows/s end to end, and unstable —
        #: the fp32 part and the pack threads read the same host DRAM at once).
        self.host_split = host_split if isinstance(host_split, float) else bool(host_split)
        if host_pack_threads is None:
            local_ws = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
            try:
                cpus = len(os.sched_getaffinity(0))
            except (AttributeError, OSError):
                cpus = os.cpu_count() or 1
            host_pack_threads = max(1, min(32, cpus // local_ws))
        self.host_pack_threads = int(host_pack_threads)
        self.last_transport = "fp32"
        self.last_h2d_bytes_per_batch = 0

    # ------------------------------------------------------------------ calibration (predictor.py:226-244)
    def _calibrate(self, logit_val: float) -> float:
        if self._cal_method == "temperature":
            return float(torch.sigmoid(torch.tensor(logit_val / self._cal_temperature)).item())
        if self._cal_method == "platt":
            return float(torch.sigmoid(torch.tensor(self._cal_platt_a * logit_val + self._cal_platt_b)).item())
        raw_prob = float(torch.sigmoid(torch.tensor(logit_val, dtype=torch.float32)).item())
        if self._cal_method == "isotonic" and self._isotonic_cal is not None:
            cal_prob = float(self._isotonic_cal.predict([[raw_prob]])[0])
            return float(np.clip(cal_prob, 0.0, 1.0))
        return raw_prob

    # ------------------------------------------------------------------ scoring
    def _infer_logits(self, visuals: Sequence[np.ndarray], audios: Sequence[np.ndarray]) -> List[float]:
        """Batched forward over host windows of one common shape, `batch_size` windows per launch."""
        n = len(visuals)
        if (self.use_cuda_graphs and 0 < n <= self.graph_max_batch and self.model is not None and self.device.type == "cuda"
                and self.model._precision() == _cabi.LSD_PREC_BF16):
            return self._infer_logits_graph(visuals, audios)

        def gen():
            for i0 in range(0, n, self.batch_size):
                v = torch.from_numpy(np.stack(visuals[i0:i0 + self.batch_size])).pin_memory()
                a = torch.from_numpy(np.stack(audios[i0:i0 + self.batch_size])).pin_memory()
                yield v, a

        out: List[float] = []
        for t in self.score_batches(gen()):
            out.extend(float(x) for x in t.tolist())
        return out

    # ------------------------------------------------------------------ CUDA-graph path for small host batches
    def _infer_logits_graph(self, visuals: Sequence[np.ndarray], audios: Sequence[np.ndarray]) -> List[float]:
        """`len(visuals) <= graph_max_batch` host windows of one shape -> logits, through a CUDA graph captured once per
        (batch, shapes, dtype): pinned staging -> H2D -> `lsd_forward` (all of its streams) -> D2H.  Same kernels, same
        numerics as the launch-by-launch path (tests compare them bitwise)."""
        m = self.model
        dev = self.device
        v = np.stack(visuals)
        a = np.stack(audios)
        dt = torch.float16 if self.use_half_precision else torch.float32
        with m._lsd_lock:
            key = (v.shape, a.shape, dt, m.state_generation())
            ent = self._graphs.get(key)
            if ent is None:
                for k in [k for k in self._graphs if k[3] != key[3]]:      # stale generations
                    del self._graphs[k]
                ent = self._graphs[key] = self._capture_graph(v.shape, a.shape, dt)
                self.graph_captures += 1
            vh, ah, oh, graph = ent["vh"], ent["ah"], ent["oh"], ent["graph"]
            vh.copy_(torch.from_numpy(v))
            ah.copy_(torch.from_numpy(a))
            graph.replay()
            torch.cuda.current_stream(dev).synchronize()
            return [float(x) for x in oh.tolist()]

    def _capture_graph(self, vshape, ashape, dt):
        m = self.model
        dev = self.device
        B, _, T, H, W = vshape
        F_, Ta = ashape[2], ashape[3]
        ent = {
            "vh": torch.empty(vshape, dtype=dt).pin_memory(), "ah": torch.empty(ashape, dtype=dt).pin_memory(),
            "vd": torch.empty(vshape, dtype=dt, device=dev), "ad": torch.empty(ashape, dtype=dt, device=dev),
            "od": torch.empty(B, dtype=torch.float32, device=dev), "oh": torch.empty(B, dtype=torch.float32).pin_memory(),
            "ws": torch.empty(m.workspace_bytes(B, T, H, W, F_, Ta) + 1024, dtype=torch.uint8, device=dev),   # private: its padding persists
        }
        ent["vh"].zero_()
        ent["ah"].zero_()

        def body():
            ent["vd"].copy_(ent["vh"], non_blocking=True)
            ent["ad"].copy_(ent["ah"], non_blocking=True)
            m.forward_into(ent["vd"], ent["ad"], ent["od"], ent["ws"])
            ent["oh"].copy_(ent["od"], non_blocking=True)

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):       # the first call of a shape uploads stage programs and zeroes the padding: not capturable
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            body()
        ent["graph"] = g
        return ent

    def score_batches(self, batches) -> List[torch.Tensor]:
        """Pipelined scoring of a sequence of HOST batches `(visual (B,3,T,H,W), audio (B,1,F,Ta))` (torch CPU tensors, ideally
        pinned): the H2D copies of the next batches run on a copy stream while batch k is scored, logits come back through
        pinned memory.  Returns one fp32 CPU tensor of logits per batch.  Same numerics as calling the model batch by batch."""
        with self.model._lsd_lock:      # the cached device slots make this call non-re-entrant; serialise like forward()
            return self._score_batches_locked(batches)

    def _score_batches_locked(self, batches) -> List[torch.Tensor]:
        dev = self.device
        m = self.model
        comp = torch.cuda.current_stream(dev)
        copy = getattr(self, "_copy_stream", None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(dev)
        # three device slots: the copy stream stays busy back to back (copy and forward take about the same time at B=64 —
        # 4.1 ms of PCIe vs 3.9 ms of compute — so two slots leave bubbles whenever either one jitters)
        NS = 3
        if getattr(self, "_slot_cache", None) is None:      # device slots and events are kept across calls
            self._slot_cache = ([None] * NS, [torch.cuda.Event() for _ in range(NS)], [torch.cuda.Event() for _ in range(NS)])
        slots, ready, free = self._slot_cache
        outs: List[torch.Tensor] = []
        pool, pool_off = None, 0
        stage = getattr(self, "_u8_stage", None)
        if stage is None:
            stage = self._u8_stage = [None] * NS
        L = _cabi.lib()
        state = {"mode": "fp32" if self.use_half_precision else self.host_transport}
        if state["mode"] == "auto" and getattr(self, "_auto_prefers_fp32", False):
            state["mode"] = "fp32"       # an earlier call on this predictor measured the pack slower than the copy (busy host / several GPUs)

        def split_count(B):
            """Windows of a batch that cross PCIe as fp32 while the host threads pack the rest (split transport): a batch costs the
            pipeline max(PCIe time, pack time), so the fp32 share f equalises  f*t32 + (1-f)*t32/4  and  (1-f)*T + 0.33 ms  (t32: the
            fp32 bytes of the whole batch at ~52 GB/s, T: the measured pack time of a whole batch).  0 until a pack has been measured."""
            if not self.host_split or B < 2:
                return 0
            if isinstance(self.host_split, float):
                return int(round(self.host_split * B))
            return int(round(getattr(self, "_split_frac", 0.0) * B)) if B >= 8 else 0

        def start(k, vh, ah):
            """Host side of batch k, first half: starts the uint8 pack on the library's host threads (the call returns at once;
            this thread goes on to enqueue batch k-1).  No Python helper thread: a second Python thread that makes CUDA calls was
            measured to slow every later single-window call of the process by ~1 ms."""
            s = k % NS
            if self.use_half_precision:
                vh, ah = vh.half(), ah.half()
            if state["mode"] != "fp32" and vh.dtype == torch.float32 and vh.device.type == "cpu" and vh.is_contiguous() and vh.numel() > 0:
                fresh = stage[s] is None or stage[s].shape != vh.shape
                if fresh:
                    stage[s] = torch.empty(vh.shape, dtype=torch.uint8).pin_memory()
                elif k >= NS:
                    ready[s].synchronize()    # the H2D copy that last read this staging buffer (batch k - NS) has finished
                B = int(vh.shape[0])
                wbytes = vh[0].numel()
                na = min(split_count(B), B - 1) if wbytes % 16 == 0 else 0       # (lsd_expand_u8 works in 16-byte pieces)
                if L.lsd_host_pack_u8_begin(vh.data_ptr() + 4 * na * wbytes, stage[s].data_ptr(), (B - na) * wbytes, self.host_pack_threads) == _cabi.LSD_OK:
                    return (k, vh, ah, s, fresh, na)
            return (k, vh, ah, None, False, 0)

        def finish(st):
            """-> (video to ship, audio, transport label, na): na > 0 = split transport (video is then (fp32 windows, packed rest))."""
            k, vh, ah, s, fresh, na = st
            if s is None:
                return vh, ah, "fp32", 0
            ok = L.lsd_host_pack_u8_end()
            if state["mode"] == "fp32":          # the transport was given up while this batch's pack was queued: drained, shipped as fp32
                return vh, ah, "fp32", 0
            if ok == 1:
                B = int(vh.shape[0])
                d = self.__dict__
                first_ever = d.setdefault("_pack_jobs_seen", 0) == 0     # the very first pack of this predictor also starts the pack threads
                d["_pack_jobs_seen"] += 1
                T_ms = L.lsd_host_pack_last_ms() * B / max(1, B - na)     # pack time of a whole batch, from this (possibly partial) pack
                t32 = vh.numel() * 4 / 52e9                               # the fp32 bytes at ~52 GB/s of PCIe gen5 x16
                steady = k >= 1 and not fresh      # not the first batch of a call, not the first pack into a freshly pinned staging buffer
                if steady and self.host_split:
                    f = (T_ms * 1e-3 + 0.33e-3 - t32 / 4) / (0.75 * t32 + T_ms * 1e-3)
                    f = min(0.6, max(0.0, f))
                    if f < 0.05:
                        f = 0.0
                    self._split_frac = f if not hasattr(self, "_split_frac") else 0.5 * self._split_frac + 0.5 * f
                if state["mode"] == "auto" and not self.host_split:
                    # "auto" stops packing only when packing is the slower way (auto_transport_gives_up_packing; histories are kept
                    # across calls, so a three-batch warm-up call can decide for the next one).  Measured (AVX-512 loop, non-temporal
                    # stores, queued jobs): 16 host threads for one GPU pack a 64-window batch in 2.2 ms against >= 4.4 ms of copy;
                    # 12 threads per GPU with two GPUs packing at once need 3.3 ms (4.3 - 4.7 ms per step for the fp32 copy there);
                    # eight GPUs fed at once: ~12 ms per pack whatever the thread count against 6.5 - 9.8 ms per step for the copy —
                    # the pack moves more host-DRAM bytes than the copy it saves, and there the host memory, not PCIe, is the limit.
                    every, recent = d.setdefault("_pack_ms_all", []), d.setdefault("_pack_ms_hist", [])
                    if not first_ever:
                        every.append(T_ms)
                        del every[:-8]
                    if steady:
                        recent.append(T_ms)
                        del recent[:-8]
                    if auto_transport_gives_up_packing(recent, every, t32):
                        state["mode"] = "fp32"             # this batch is packed and exact: it still ships packed, the next ones as fp32
                        self._auto_prefers_fp32 = True     # remembered for the later calls of this predictor
                if na > 0:
                    return (vh, stage[s]), ah, "u8 + fp32 split (host-packed part exact)", na
                return stage[s], ah, "u8 (host-packed, exact)", 0
            state["mode"] = "fp32"               # not k/255 data: stop checking for the rest of this call
            return vh, ah, "fp32", 0

        it = iter(batches)
        pend = []                       # host side started, not yet shipped: batches k, k+1 (the library queues two pack jobs)
        started = [0]

        def fill():
            while len(pend) < 2:
                nxt = next(it, None)
                if nxt is None:
                    return
                pend.append(start(started[0], *nxt))
                started[0] += 1

        fill()
        k = 0
        split_slots = getattr(self, "_split_slots", None)
        if split_slots is None:
            split_slots = self._split_slots = [None] * NS
        try:
            while pend:
                vh, ah, transport, na = finish(pend.pop(0))                     # batch k is packed (or goes as it is)
                fill()          # batch k+1 is being packed while batch k is enqueued, the pack of batch k+2 is queued behind it
                s = k % NS
                self.last_transport = transport        # of the last batch shipped
                if na > 0:
                    # split transport: windows [0, na) cross PCIe as fp32, the rest packed; the packed part is expanded next to the
                    # fp32 part on the device (lsd_expand_u8: the reference's own astype(float32) / 255.0, bit for bit) and ONE fp32
                    # batch is scored
                    v32, v8 = vh
                    B = int(v32.shape[0])
                    wbytes = v32[0].numel()
                    self.last_h2d_bytes_per_batch = 4 * na * wbytes + (B - na) * wbytes + ah.numel() * ah.element_size()
                    ss = split_slots[s]
                    if ss is None or ss[0].shape != v32.shape or ss[2].shape != ah.shape:
                        ss = split_slots[s] = (torch.empty(v32.shape, dtype=torch.float32, device=dev), torch.empty(v32.shape, dtype=torch.uint8, device=dev),
                                               torch.empty(ah.shape, dtype=ah.dtype, device=dev))
                        free[s].record(comp)
                    with torch.cuda.stream(copy):
                        copy.wait_event(free[s])          # the forward that last read this slot has finished
                        ss[0][:na].copy_(v32[:na], non_blocking=True)
                        ss[1][:B - na].copy_(v8[:B - na], non_blocking=True)
                        ss[2].copy_(ah, non_blocking=True)
                        ready[s].record(copy)
                    comp.wait_event(ready[s])
                    rc = L.lsd_expand_u8(ss[1].data_ptr(), ss[0].data_ptr() + 4 * na * wbytes, (B - na) * wbytes, comp.cuda_stream)
                    if rc != _cabi.LSD_OK:
                        raise RuntimeError(f"lsd_expand_u8 failed ({rc})")
                    logits = m(ss[0], ss[2])
                else:
                    self.last_h2d_bytes_per_batch = vh.numel() * vh.element_size() + ah.numel() * ah.element_size()
                    if slots[s] is None or slots[s][0].shape != vh.shape or slots[s][0].dtype != vh.dtype or slots[s][1].shape != ah.shape:
                        slots[s] = (torch.empty(vh.shape, dtype=vh.dtype, device=dev), torch.empty(ah.shape, dtype=ah.dtype, device=dev))
                        free[s].record(comp)
                    with torch.cuda.stream(copy):
                        copy.wait_event(free[s])          # the forward that last read this slot has finished
                        slots[s][0].copy_(vh, non_blocking=True)
                        slots[s][1].copy_(ah, non_blocking=True)
                        ready[s].record(copy)
                    comp.wait_event(ready[s])
                    logits = m(slots[s][0], slots[s][1])
                free[s].record(comp)
                # logits come back through one pinned block per 256 batches (a pinned allocation per step costs ~0.1-0.4 ms)
                nb = int(logits.numel())
                if pool is None or pool_off + nb > pool.numel():
                    pool = torch.empty(max(256 * nb, 4096), dtype=torch.float32, pin_memory=True)
                    pool_off = 0
                host = pool[pool_off:pool_off + nb].view(logits.shape)
                pool_off += nb
                host.copy_(logits.float(), non_blocking=True)
                outs.append(host)
                k += 1
        except BaseException:
            for st in pend:
                if st[3] is not None:
                    L.lsd_host_pack_u8_end()     # never leave a pack job in flight behind an error
            raise
        comp.synchronize()
        return outs

    def _infer_confidences(self, visuals: Sequence[np.ndarray], audios: Sequence[np.ndarray]) -> List[float]:
        return [self._calibrate(l) for l in self._infer_logits(visuals, audios)]

    def _infer_confidence(self, visual_np: np.ndarray, audio_np: np.ndarray) -> float:
        """Run a single forward pass, apply output calibration, return P(REAL)."""
        return self._infer_confidences([visual_np], [audio_np])[0]

    # ------------------------------------------------------------------ aggregation helpers (host numpy, as the reference)
    def _robust_confidence(self, confidences: List[float]) -> float:
        """predictor.py:246-260 (single implementation: `lipsync_b200.aggregate._robust`)."""
        return _agg._robust(confidences, self.confidence_smoothing, self.trim_ratio)

    def _speech_weighted_confidence(self, confidences: List[float], speaking_scores: List[float],
                                    vad_weights: Optional[List[float]] = None) -> float:
        """predictor.py:262-293 (single implementation: `lipsync_b200.aggregate._speech_weighted`)."""
        return _agg._speech_weighted(confidences, speaking_scores, vad_weights, self.confidence_smoothing, self.trim_ratio)

    @staticmethod
    def _smoothing_windows(t_v: int, t_a: int) -> Tuple[List[Tuple[int, int, int, int]], List[Tuple[int, int]]]:
        """Window spans of `_temporal_smoothed_confidence` (predictor.py:302-325): (v0, v1, a0, a1) + reported spans."""
        wins = [(0, t_v, 0, t_a)]
        spans = [(0, max(1, t_v))]
        win_v = max(12, t_v // 2)
        win_a = max(48, t_a // 2)
        if t_v >= win_v and t_a >= win_a:
            v_starts = [0, max(0, (t_v - win_v) // 2), max(0, t_v - win_v)]
            for v_start in v_starts:
                v_end = min(t_v, v_start + win_v)
                a_start = int(round(v_start * (t_a / max(1, t_v))))
                a_end = min(t_a, a_start + win_a)
                if (v_end - v_start) >= 16 and (a_end - a_start) >= 64:
                    wins.append((v_start, v_end, a_start, a_end))
                    spans.append((v_start, v_end))
        return wins, spans

    def _temporal_smoothed_confidence(self, visual_np: np.ndarray, audio_np: np.ndarray):
        t_v = int(visual_np.shape[1])
        t_a = int(audio_np.shape[2])
        wins, spans = self._smoothing_windows(t_v, t_a)
        confidences: List[float] = []
        # full window and the half windows have different shapes: one batch per distinct shape, order preserved
        by_shape = {}
        for i, (v0, v1, a0, a1) in enumerate(wins):
            by_shape.setdefault((v1 - v0, a1 - a0), []).append(i)
        res = [0.0] * len(wins)
        for idxs in by_shape.values():
            vs = [np.ascontiguousarray(visual_np[:, wins[i][0]:wins[i][1]]) for i in idxs]
            as_ = [np.ascontiguousarray(audio_np[:, :, wins[i][2]:wins[i][3]]) for i in idxs]
            for i, c in zip(idxs, self._infer_confidences(vs, as_)):
                res[i] = c
        confidences = res
        return self._robust_confidence(confidences), confidences, spans

    @staticmethod
    def _audio_start(v_start: int, total_a: int, total_v_frames: int, chunk_a_size: int = 128) -> int:
        a_ratio = total_a / max(1, total_v_frames)
        a_start = int(round(v_start * a_ratio))
        if a_start + chunk_a_size > total_a:
            a_start = max(0, total_a - chunk_a_size)
        return a_start

    def _align_audio_chunk(self, audio_np_full: np.ndarray, v_start: int, total_v_frames: int, chunk_a_size: int = 128) -> np.ndarray:
        total_a = int(audio_np_full.shape[2])
        a_start = self._audio_start(v_start, total_a, total_v_frames, chunk_a_size)
        a_end = min(total_a, a_start + chunk_a_size)
        chunk = audio_np_full[:, :, a_start:a_end]
        if chunk.shape[2] < chunk_a_size:
            pad = np.repeat(chunk[:, :, -1:], chunk_a_size - chunk.shape[2], axis=2)
            chunk = np.concatenate([chunk, pad], axis=2)
        return chunk

    def _run_chunked_inference(self, chunks: List[np.ndarray], chunk_starts: List[int], audio_np_full: np.ndarray,
                               total_v_frames: int) -> Tuple[float, List[float]]:
        """Score every chunk of a track (batched) and aggregate; returns (aggregated_confidence, per_chunk_confidences)."""
        audios = [self._align_audio_chunk(audio_np_full, v_start, total_v_frames) for v_start in chunk_starts]
        chunk_confs = self._infer_confidences(list(chunks), audios) if len(chunks) else []
        return self._robust_confidence(chunk_confs), chunk_confs

    # ------------------------------------------------------------------ device-side window builder
    def score_track_logits(self, track_u8: torch.Tensor, starts: Sequence[int], mel_full: torch.Tensor, total_v_frames: int,
                           chunk_a_size: int = 128, audio_starts_from: Optional[Sequence[int]] = None) -> torch.Tensor:
        """uint8 track `(n_frames,H,W,3)` + window start frames + clip log-mel `(1,F,Ta_full)` (both on the device)
        -> fp32 logits `(n_windows,)` on the device.  Windows are built on the GPU (`/255`, audio alignment).
        `audio_starts_from`: absolute start frames used for the audio alignment when `track_u8` is only a rank-local span
        of the full track and `starts` are relative to that span (sharded long videos)."""
        m = self.model
        dev = m._device()
        if track_u8.dtype != torch.uint8 or track_u8.dim() != 4 or track_u8.shape[3] != 3:
            raise ValueError(f"track must be uint8 (n_frames, H, W, 3), got {track_u8.dtype} {tuple(track_u8.shape)}")
        if mel_full.dim() != 3 or mel_full.shape[0] != 1:
            raise ValueError(f"mel_full must be (1, F, T_full), got {tuple(mel_full.shape)}")
        n = len(starts)
        logits = torch.empty(n, dtype=torch.float32, device=dev)
        if n == 0:
            return logits
        track_u8 = track_u8.contiguous()
        mel_full = mel_full.to(torch.float32).contiguous()
        n_frames, H, W = int(track_u8.shape[0]), int(track_u8.shape[1]), int(track_u8.shape[2])
        F_, Ta_full = int(mel_full.shape[1]), int(mel_full.shape[2])
        with m._lsd_lock:
            h = m._ensure_handle(dev)
            L = _cabi.lib()
            prec = m._precision()
            batch = min(self.batch_size, n)
            need = L.lsd_score_workspace_bytes(h.ptr, batch, self.chunk_size, H, W, F_, chunk_a_size, prec)
            if need == 0:
                _cabi.check(h.ptr, _cabi.LSD_ERR_SHAPE)
            if n > batch and prec == _cabi.LSD_PREC_BF16:
                need = 2 * need      # a double workspace lets lsd_score_windows overlap the tail of batch k with the encoder of batch k+1
            ws = m._workspace(need, dev)
            st = (C.c_int32 * n)(*[int(s) for s in starts])
            ast = None
            if audio_starts_from is not None:
                if len(audio_starts_from) != n:
                    raise ValueError("audio_starts_from must have one entry per window")
                ast = (C.c_int32 * n)(*[self._audio_start(int(v), Ta_full, int(total_v_frames), chunk_a_size) for v in audio_starts_from])
            rc = L.lsd_score_windows(h.ptr, track_u8.data_ptr(), n_frames, H, W, st, ast, n, self.chunk_size, mel_full.data_ptr(),
                                     F_, Ta_full, int(total_v_frames), chunk_a_size, prec, batch, logits.data_ptr(),
                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(h.ptr, rc)
        return logits

    def score_track(self, track_u8: torch.Tensor, starts: Sequence[int], mel_full: torch.Tensor, total_v_frames: int):
        logits = self.score_track_logits(track_u8, starts, mel_full, total_v_frames)
        confs = [self._calibrate(float(x)) for x in logits.cpu().tolist()]
        return self._robust_confidence(confs), confs

    # ------------------------------------------------------------------ speaking alignment / mouth motion
    def window_speech_stats(self, track_u8: torch.Tensor, starts: Sequence[int], mel_full: torch.Tensor, total_v_frames: int,
                            chunk_a_size: int = 128, audio_starts_from: Optional[Sequence[int]] = None):
        """Per-window `(speaking_alignment, mouth_motion_energy, audio_energy)` fp32 device tensors for the windows of a
        uint8 track: the statistics `_predict_long_video` computes per window with `_speaking_alignment_score`
        (predictor.py:333-370) and `_mouth_motion_energy_check` (:374-419).  Frame-difference energies are computed once
        per frame pair of the track (`lsd_track_motion`), then combined per window (`lsd_speech_stats`)."""
        m = self.model
        dev = m._device()
        if track_u8.dtype != torch.uint8 or track_u8.dim() != 4 or track_u8.shape[3] != 3:
            raise ValueError(f"track must be uint8 (n_frames, H, W, 3), got {track_u8.dtype} {tuple(track_u8.shape)}")
        if mel_full.dim() != 3 or mel_full.shape[0] != 1:
            raise ValueError(f"mel_full must be (1, F, T_full), got {tuple(mel_full.shape)}")
        n = len(starts)
        out = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3)]
        if n == 0:
            return tuple(out)
        track_u8 = track_u8.contiguous()
        mel_full = mel_full.to(torch.float32).contiguous()
        n_frames, H, W = int(track_u8.shape[0]), int(track_u8.shape[1]), int(track_u8.shape[2])
        F_, Ta_full = int(mel_full.shape[1]), int(mel_full.shape[2])
        motion = torch.zeros(2, max(1, n_frames - 1), dtype=torch.float32, device=dev)
        idx = torch.empty(2 * n, dtype=torch.int32, device=dev)
        with m._lsd_lock:
            h = m._ensure_handle(dev)
            L = _cabi.lib()
            stream = torch.cuda.current_stream(dev).cuda_stream
            _cabi.check(h.ptr, L.lsd_track_motion(h.ptr, track_u8.data_ptr(), _cabi.LSD_U8, _cabi.LSD_NDHWC, n_frames, H, W,
                                                  motion[0].data_ptr(), motion[1].data_ptr(), stream))
            st = (C.c_int32 * n)(*[int(s) for s in starts])
            ast = None
            if audio_starts_from is not None:
                ast = (C.c_int32 * n)(*[self._audio_start(int(v), Ta_full, int(total_v_frames), chunk_a_size) for v in audio_starts_from])
            rc = L.lsd_speech_stats(h.ptr, motion[0].data_ptr(), motion[1].data_ptr(), n_frames, st, ast, n, self.chunk_size,
                                    mel_full.data_ptr(), F_, Ta_full, int(total_v_frames), chunk_a_size, out[0].data_ptr(),
                                    out[1].data_ptr(), out[2].data_ptr(), idx.data_ptr(), stream)
            _cabi.check(h.ptr, rc)
        return tuple(out)

    def _speaking_alignment_score(self, visual_np: np.ndarray, audio_np: np.ndarray) -> float:
        """Same contract as the reference static method (predictor.py:333-370): `(3,T,H,W)` float32 crop in [0,1] +
        `(1,F,T_a)` log-mel -> score in [0,1]; computed on the device."""
        m = self.model
        dev = m._device()
        v = torch.from_numpy(np.ascontiguousarray(visual_np, dtype=np.float32)).to(dev)
        a = torch.from_numpy(np.ascontiguousarray(audio_np, dtype=np.float32)).to(dev)
        T, H, W = int(v.shape[1]), int(v.shape[2]), int(v.shape[3])
        Ta = int(a.shape[2])
        if T < 2 or Ta < 2:
            return 0.5
        motion = torch.zeros(2, T - 1, dtype=torch.float32, device=dev)
        out = torch.empty(3, dtype=torch.float32, device=dev)
        idx = torch.empty(2, dtype=torch.int32, device=dev)
        with m._lsd_lock:
            h = m._ensure_handle(dev)
            L = _cabi.lib()
            stream = torch.cuda.current_stream(dev).cuda_stream
            _cabi.check(h.ptr, L.lsd_track_motion(h.ptr, v.data_ptr(), _cabi.LSD_F32, _cabi.LSD_NCDHW, T, H, W,
                                                  motion[0].data_ptr(), motion[1].data_ptr(), stream))
            st = (C.c_int32 * 1)(0)
            rc = L.lsd_speech_stats(h.ptr, motion[0].data_ptr(), motion[1].data_ptr(), T, st, st, 1, T, a.data_ptr(), int(a.shape[1]), Ta,
                                    T, Ta, out[0:1].data_ptr(), out[1:2].data_ptr(), out[2:3].data_ptr(), idx.data_ptr(), stream)
            _cabi.check(h.ptr, rc)
        return float(out[0].item())

    def mouth_motion_checks(self, mouth_motion: Sequence[float], audio_energy: Sequence[float], mouth_motion_low_threshold: float = 0.015,
                            audio_energy_high_threshold: float = -25.0, audio_energy_low_threshold: float = -50.0) -> List[dict]:
        """Decision logic of `_mouth_motion_energy_check` (predictor.py:403-419) on the per-window statistics."""
        res = []
        for motion, energy in zip(mouth_motion, audio_energy):
            motion, energy = float(motion), float(energy)
            if energy > audio_energy_high_threshold and motion < mouth_motion_low_threshold:
                r = "likely_fake"
            elif energy < audio_energy_low_threshold and motion < mouth_motion_low_threshold:
                r = "uncertain"
            else:
                r = "no_issue"
            res.append({"audio_energy": round(energy, 4), "mouth_motion_energy": round(motion, 6), "check_result": r})
        return res

    @staticmethod
    def aggregate_mouth_motion_checks(checks: Sequence[dict], max_samples: int = 5) -> dict:
        """`_aggregate_mouth_motion_check` (predictor.py:463-523): sample up to `max_samples` evenly spaced windows (always the
        last one), majority vote with the conservative `uncertain` rule."""
        n = len(checks)
        if n == 0:
            return {"check_result": "no_data", "audio_energy": 0.0, "mouth_motion_energy": 0.0, "samples_checked": 0}
        if n <= max_samples:
            indices = list(range(n))
        else:
            step = n / max_samples
            indices = [int(i * step) for i in range(max_samples)]
            if (n - 1) not in indices:
                indices[-1] = n - 1
        counts = {"likely_fake": 0, "uncertain": 0, "no_issue": 0}
        energies, motions = [], []
        for i in indices:
            c = checks[i]
            counts[c["check_result"]] = counts.get(c["check_result"], 0) + 1
            energies.append(float(c["audio_energy"]))
            motions.append(float(c["mouth_motion_energy"]))
        k = len(indices)
        if counts["uncertain"] > k // 2:
            agg = "uncertain"
        elif counts["likely_fake"] > counts["uncertain"] + counts["no_issue"]:
            agg = "likely_fake"
        else:
            agg = "no_issue"
        return {"check_result": agg, "audio_energy": round(float(np.median(energies)), 4),
                "mouth_motion_energy": round(float(np.median(motions)), 6), "samples_checked": k, "counts": counts}

    # ------------------------------------------------------------------ decision (predictor.py:856-1155, :1235)
    def aggregate_long_video(self, window_confs, window_speaking, window_vad_weights=None, mouth_check_result: str = "no_data", **gates):
        """Final `real` / `fake` / `uncertain` decision from per-window confidences; see `lipsync_b200.aggregate`."""
        from .aggregate import aggregate_long_video
        return aggregate_long_video(window_confs, window_speaking, window_vad_weights, mouth_check_result,
                                    confidence_smoothing=self.confidence_smoothing, trim_ratio=self.trim_ratio, **gates)

    # ------------------------------------------------------------------ multi-GPU
    def score_windows_sharded(self, n_windows: int, score_range: Callable[[int, int], torch.Tensor],
                              world_size: int, rank: int) -> torch.Tensor:
        """Each rank scores its contiguous block `[lo,hi)` with `score_range(lo, hi) -> logits`, then one
        all-gather returns all `n_windows` fp32 logits on every rank (rank 0 runs the host aggregation)."""
        lo, hi = partition_windows(n_windows, world_size, rank)
        local = score_range(lo, hi) if hi > lo else torch.empty(0, dtype=torch.float32, device=self.device)
        return gather_logits(local, n_windows, world_size, rank)
