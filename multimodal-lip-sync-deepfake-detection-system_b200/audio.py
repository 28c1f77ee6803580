"""Log-mel front end: host mirror of `preprocess_audio` (app/preprocessing/audio.py:47-102).

The STFT / mel / dB arithmetic the reference delegates to librosa (`melspectrogram(n_fft=400, hop=160,
win=400, n_mels=80, power=2)` + `power_to_db(ref=np.max)`) runs in `liblsd_b200.so` (`lsd_logmel`);
pad-by-repeat / truncate to `target_frames` follows audio.py:93-99.  Decoding is limited to what the
standard library can read (16 kHz mono PCM WAV — what the reference's ffmpeg step produces, audio.py:19-29).
"""
from __future__ import annotations

import ctypes as C
import wave
from pathlib import Path
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _cabi

N_MELS, HOP, N_FFT, SR = 80, 160, 400, 16000
_handles = {}


def _handle(dev: torch.device) -> _cabi.Handle:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _handles:
        _handles[idx] = _cabi.Handle(idx)
    return _handles[idx]


def logmel_db(pcm: Union[torch.Tensor, Sequence[torch.Tensor]]):
    """fp32 mono 16 kHz PCM on a CUDA device -> log-mel dB `(80, 1 + n//160)` in [-80, 0], `ref=max` per clip.
    A list of clips (or a 2-D `(B, n)` tensor of equal-length clips) is processed in one call; one maximum per clip."""
    if isinstance(pcm, torch.Tensor) and pcm.dim() == 2:
        return _logmel_db_equal(pcm)
    single = isinstance(pcm, torch.Tensor)
    clips = [pcm] if single else list(pcm)
    if not clips:
        return []
    dev = clips[0].device
    if dev.type != "cuda":
        raise RuntimeError("logmel_db needs PCM on an sm_100 CUDA device; there is no CPU fallback")
    for c in clips:
        if c.dim() != 1:
            raise ValueError(f"each clip must be a 1-D PCM tensor, got {tuple(c.shape)}")
        if c.numel() == 0:
            raise ValueError("Empty audio signal")
    flat = torch.cat([c.to(torch.float32) for c in clips]).contiguous()
    L = _cabi.lib()
    offs, moffs, frames = [0], [0], []
    for c in clips:
        offs.append(offs[-1] + c.numel())
        fr = L.lsd_logmel_frames(c.numel())
        frames.append(fr)
        moffs.append(moffs[-1] + N_MELS * fr)
    out = torch.empty(moffs[-1], dtype=torch.float32, device=dev)
    scratch = torch.empty(len(clips), dtype=torch.float32, device=dev)
    h = _handle(dev)
    with h.lock:
        co = (C.c_int64 * len(offs))(*offs)
        mo = (C.c_int64 * len(moffs))(*moffs)
        rc = L.lsd_logmel(h.ptr, flat.data_ptr(), co, len(clips), out.data_ptr(), mo, scratch.data_ptr(),
                          torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(h.ptr, rc)
    mels = [out[moffs[i]:moffs[i + 1]].view(N_MELS, frames[i]) for i in range(len(clips))]
    return mels[0] if single else mels


def _logmel_db_equal(pcm: torch.Tensor) -> torch.Tensor:
    """`(B, n)` equal-length clips -> `(B, 80, 1 + n//160)`; one launch pair for the whole batch, no per-clip host work
    beyond the offset table."""
    if pcm.device.type != "cuda":
        raise RuntimeError("logmel_db needs PCM on an sm_100 CUDA device; there is no CPU fallback")
    B, n = int(pcm.shape[0]), int(pcm.shape[1])
    if n == 0:
        raise ValueError("Empty audio signal")
    L = _cabi.lib()
    fr = L.lsd_logmel_frames(n)
    out = torch.empty(B, N_MELS, fr, dtype=torch.float32, device=pcm.device)
    if B == 0:
        return out
    flat = pcm.to(torch.float32).contiguous()
    scratch = torch.empty(B, dtype=torch.float32, device=pcm.device)
    co = (C.c_int64 * (B + 1))(*range(0, (B + 1) * n, n))
    mo = (C.c_int64 * (B + 1))(*range(0, (B + 1) * N_MELS * fr, N_MELS * fr))
    h = _handle(pcm.device)
    with h.lock:
        rc = L.lsd_logmel(h.ptr, flat.data_ptr(), co, B, out.data_ptr(), mo, scratch.data_ptr(),
                          torch.cuda.current_stream(pcm.device).cuda_stream)
        _cabi.check(h.ptr, rc)
    return out


def fit_frames(mel_db: np.ndarray, target_frames: Optional[int]) -> np.ndarray:
    """audio.py:93-99: repeat the last column to pad, or truncate, along the time axis of `(1, F, T)`."""
    if target_frames is None:
        return mel_db
    t_cur = mel_db.shape[2]
    if t_cur < target_frames:
        padding = np.repeat(mel_db[:, :, -1:], target_frames - t_cur, axis=2)
        mel_db = np.concatenate([mel_db, padding], axis=2)
    elif t_cur > target_frames:
        mel_db = mel_db[:, :, :target_frames]
    return mel_db


def preprocess_audio_pcm(y: np.ndarray, target_frames: Optional[int] = None, device: str = "cuda") -> np.ndarray:
    """float32 mono 16 kHz samples -> `(1, 80, T)` float32 log-mel dB (the array `preprocess_audio` returns)."""
    y = np.asarray(y, dtype=np.float32).reshape(-1)
    if y.size == 0:
        raise ValueError("Empty audio signal")
    mel = logmel_db(torch.from_numpy(y).to(device))
    mel_db = mel.cpu().numpy().astype("float32")[None]
    return fit_frames(mel_db, target_frames)


def preprocess_audio(path: Path, sr: int = 16000, n_mels: int = 80, hop_length: int = 160, win_length: int = 400,
                     target_frames: Optional[int] = None) -> np.ndarray:
    """Same signature as the reference.  Only the reference's own settings are implemented in the kernel."""
    if (sr, n_mels, hop_length, win_length) != (SR, N_MELS, HOP, N_FFT):
        raise NotImplementedError("the log-mel kernel implements the reference configuration (16 kHz, 80 mel, hop 160, win 400)")
    path = Path(path)
    with wave.open(str(path), "rb") as w:
        if w.getframerate() != sr or w.getnchannels() != 1 or w.getsampwidth() != 2:
            raise ValueError(f"{path}: expected 16 kHz mono s16 WAV (ffmpeg -ar 16000 -ac 1 pcm_s16le, audio.py:23-27)")
        raw = w.readframes(w.getnframes())
    y = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    if y.size == 0:
        raise ValueError(f"Empty audio signal for {path}")
    return preprocess_audio_pcm(y, target_frames)


def detect_voice_activity_pcm(y: np.ndarray, sr: int = 16000, trigger_level: float = 7.0, trigger_time: float = 0.25,
                              search_time: float = 1.0, allowed_gap: float = 0.25, device: str = "cuda"):
    """`detect_voice_activity` (app/preprocessing/audio.py:105-245) on decoded samples: `(voice_mask bool (ceil(n/160),),
    duration_sec)`.  Frame energies, thresholding and the 3-frame smoothing run on the device (`lsd_frame_energy`,
    `lsd_vad_mask`); the threshold follows audio.py:196-212 — numpy median / 20th percentile of the energies and, as in the
    reference, the energy of `torchaudio.functional.vad`'s trimmed waveform (library call on the host, used only as a cap)."""
    y = np.asarray(y, dtype=np.float32).reshape(-1)
    if y.size == 0:
        return np.ones(1, dtype=bool), 0.0
    duration_sec = len(y) / sr
    dev = torch.device(device)
    L = _cabi.lib()
    n_frames = L.lsd_vad_frames(y.size)
    pcm = torch.from_numpy(y).to(dev)
    energy = torch.empty(n_frames, dtype=torch.float32, device=dev)
    mask = torch.empty(n_frames, dtype=torch.uint8, device=dev)
    h = _handle(pcm.device)
    stream = torch.cuda.current_stream(pcm.device).cuda_stream
    with h.lock:
        _cabi.check(h.ptr, L.lsd_frame_energy(h.ptr, pcm.data_ptr(), y.size, energy.data_ptr(), stream))
    frame_energies = energy.cpu().numpy()
    energy_median = np.median(frame_energies)
    energy_p20 = np.percentile(frame_energies, 20)
    threshold = min(energy_p20, energy_median * 0.05)
    threshold = max(1e-8, threshold)
    import torchaudio.functional as AF
    vad_waveform = AF.vad(waveform=torch.from_numpy(y).float().unsqueeze(0), sample_rate=sr, trigger_level=trigger_level,
                          trigger_time=trigger_time, search_time=search_time, allowed_gap=allowed_gap)
    if vad_waveform.numel() > 0:
        vad_energy = float(torch.mean(vad_waveform ** 2))
        threshold = min(threshold, max(1e-8, vad_energy * 0.05))
    with h.lock:
        _cabi.check(h.ptr, L.lsd_vad_mask(h.ptr, energy.data_ptr(), n_frames, float(threshold), mask.data_ptr(), stream))
    return mask.cpu().numpy().astype(bool), duration_sec
