"""Deterministic synthetic 16 kHz PCM used by tests and bench.py: seeded noise through a
1-pole low-pass plus gated sinusoids, clipped to [-1, 1].  Pure input generation - no checking logic."""
import numpy as np

SAMPLE_RATE = 16000
WINDOW_SAMPLES = 480000


def synthetic_pcm(window_idx: int, n: int = WINDOW_SAMPLES) -> np.ndarray:
    """Deterministic synthetic 16 kHz PCM: seeded noise through a
    1-pole LPF plus gated sinusoids, clipped to [-1, 1]."""
    rng = np.random.default_rng(1234 + window_idx)
    noise = 0.1 * rng.standard_normal(n)
    # 1-pole LPF y[t] = a*y[t-1] + (1-a)*x[t], vectorised via lfilter-free recurrence
    a = 0.9
    try:
        from scipy.signal import lfilter
        y = lfilter([1 - a], [1, -a], noise)
    except Exception:  # pragma: no cover
        y = np.empty(n)
        acc = 0.0
        for i in range(n):
            acc = a * acc + (1 - a) * noise[i]
            y[i] = acc
    t = np.arange(n) / SAMPLE_RATE
    tones = 0.05 * (np.sin(2 * np.pi * 220 * t) + np.sin(2 * np.pi * 440 * t) + np.sin(2 * np.pi * 880 * t))
    phase = (t % 3.0) / 0.5
    gate = np.where(phase < 1.0, 0.5 - 0.5 * np.cos(2 * np.pi * phase), 0.0)
    return np.clip(y + tones * gate, -1.0, 1.0).astype(np.float32)
