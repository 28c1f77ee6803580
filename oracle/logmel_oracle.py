"""ORACLE (test infrastructure, not product code): numpy restatement of the reference's log-mel front end.

Follows `preprocess_audio` (app/preprocessing/audio.py:80-99 of the reference) whose arithmetic lives in the
third-party, un-vendored dependency **librosa** (`requirements.txt:12`: `librosa>=0.10`, unpinned, absent from
/root/reference and not installable here).  This file restates librosa>=0.10's published algorithm
(`feature.melspectrogram` -> `stft(center=True, pad_mode="constant", window="hann")`, `filters.mel(htk=False,
norm="slaney")`, `power_to_db(ref=np.max, amin=1e-10, top_db=80)`); SURVEY.md App. D.

PARITY UNPINNED against real librosa (the reference has no golden vectors for this path and librosa cannot be
run here).  It is cross-checked against an independent implementation, `torchaudio.transforms.MelSpectrogram`
(`tests/golden/make_logmel_golden.py` commits that output as a fixture); the two agree to < 1e-3 dB.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

SR, N_FFT, HOP, N_MELS = 16000, 400, 160, 80


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank() -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=400, n_mels=80, fmin=0, fmax=sr/2, htk=False, norm='slaney') -> (80, 201) f32."""
    fftfreqs = np.linspace(0.0, SR / 2.0, 1 + N_FFT // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(SR / 2.0), N_MELS + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((N_MELS, 1 + N_FFT // 2), dtype=np.float32)
    for i in range(N_MELS):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: N_MELS + 2] - mel_f[:N_MELS])
    weights *= enorm[:, None]
    return weights


def mel_power(y: np.ndarray) -> np.ndarray:
    """|STFT|^2 projected on the mel filterbank: (80, 1 + len(y)//160) float32."""
    y = np.asarray(y, dtype=np.float32).reshape(-1)
    n = np.arange(N_FFT)
    window = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)).astype(np.float32)  # periodic Hann (fftbins=True)
    yp = np.pad(y, (N_FFT // 2, N_FFT // 2), mode="constant")                   # center=True, pad_mode="constant"
    n_frames = 1 + len(y) // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    frames = yp[idx] * window[None, :]
    spec = np.fft.rfft(frames.astype(np.float32), n=N_FFT, axis=1).astype(np.complex64)  # (frames, 201)
    power = (spec.real.astype(np.float32) ** 2 + spec.imag.astype(np.float32) ** 2).T     # (201, frames)
    return (mel_filterbank() @ power).astype(np.float32)


def power_to_db(S: np.ndarray, amin: float = 1e-10, top_db: float = 80.0) -> np.ndarray:
    """librosa.power_to_db(S, ref=np.max)."""
    S = np.asarray(S, dtype=np.float32)
    ref = np.max(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref))
    return np.maximum(log_spec, log_spec.max() - top_db).astype(np.float32)


def preprocess_audio_pcm(y: np.ndarray, target_frames=None) -> np.ndarray:
    """audio.py:80-99 on an in-memory signal: (1, 80, T) float32 log-mel dB."""
    mel_db = power_to_db(mel_power(y))[None]
    if target_frames is not None:
        t = mel_db.shape[2]
        if t < target_frames:
            mel_db = np.concatenate([mel_db, np.repeat(mel_db[:, :, -1:], target_frames - t, axis=2)], axis=2)
        elif t > target_frames:
            mel_db = mel_db[:, :, :target_frames]
    return mel_db
