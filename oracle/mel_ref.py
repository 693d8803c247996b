"""Log-mel oracle (test infrastructure; see oracle/__init__.py).

Restates what the reference's `FeatureExtractor.logMelSpectrogram`
(`Sources/WhisperKit/Core/FeatureExtractor.swift:40-56`) obtains from the opaque
`MelSpectrogram.mlmodelc`: input `"audio" [480000] f32`
(`Sources/WhisperKit/Core/Models.swift:848-870`), output
`"melspectrogram_features" [1, nMels, 1, 3000] f16` (`Models.swift:873-904`).
The arithmetic is the published OpenAI Whisper front end (whisper/audio.py):
n_fft 400, hop 160, periodic Hann, centred reflect-padded STFT, drop the last
frame, |X|^2, Slaney mel (80 or 128 bins), log10(max(., 1e-10)),
max(x, x.max() - 8), (x + 4) / 4.

`padOrTrimAudio` (`Sources/WhisperKit/Core/Audio/AudioProcessor.swift:151-174`)
is `pad_or_trim` below.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
WINDOW_SAMPLES = 480000  # Models.swift:1457 defaultWindowSamples
N_FRAMES = 3000
N_BINS = N_FFT // 2 + 1


def pad_or_trim(audio: np.ndarray, start: int = 0, n: int = WINDOW_SAMPLES) -> np.ndarray:
    """AudioProcessor.padOrTrimAudio (AudioProcessor.swift:151-174): copy
    [start, start+n) and zero-fill to n samples."""
    out = np.zeros(n, dtype=np.float32)
    seg = np.asarray(audio, dtype=np.float32)[start:start + n]
    out[: len(seg)] = seg
    return out


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filters(n_mels: int) -> np.ndarray:
    """Slaney-normalised triangular filterbank [201, n_mels] (float64),
    librosa.filters.mel(sr=16000, n_fft=400, n_mels) transposed."""
    fft_freqs = np.linspace(0, SAMPLE_RATE / 2, N_BINS)
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(SAMPLE_RATE / 2), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fft_freqs[None, :]  # [n_mels+2, 201]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))  # [n_mels, 201]
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    w = w * enorm[:, None]
    return w.T.copy()


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)


def log_mel(audio: np.ndarray, n_mels: int, dtype=np.float64) -> np.ndarray:
    """[480000] f32 -> [n_mels, 3000] (float64 math unless dtype given).
    Matches HF WhisperFeatureExtractor / openai-whisper to ~1e-6."""
    x = np.asarray(audio, dtype=dtype)
    assert x.ndim == 1
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n_frames = 1 + (len(xp) - N_FFT) // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    frames = xp[idx] * hann_periodic().astype(dtype)[None, :]
    spec = np.fft.rfft(frames, axis=1)  # [n_frames, 201]
    power = (spec.real ** 2 + spec.imag ** 2)[:-1]  # drop last frame
    mel = power @ mel_filters(n_mels).astype(dtype)  # [3000, n_mels]
    log_spec = np.log10(np.maximum(mel, 1e-10))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec.T.copy()


from whisperkit_b200.synthetic import synthetic_pcm  # noqa: E402,F401  (shared input generator)
