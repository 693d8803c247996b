// Host-side constant tables for the log-mel kernel (window, twiddles, sparse Slaney filterbank).
// Shared by mel.cu and the CPU replay harness in tests/hostcheck.
#pragma once
#include <math.h>

#include <vector>

#include "mel_core.cuh"

namespace wk {
using namespace mel;
// Host: tables (double precision), Slaney mel filterbank as in librosa / openai-whisper mel_filters.npz
static inline double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static inline double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

inline void mel_host_tables(int n_mels, std::vector<float>& win, std::vector<cf>& tw400, std::vector<cf>& tw25,
                     std::vector<float>& wts, std::vector<int>& start) {
    const double PI = 3.14159265358979323846;
    win.resize(kNfft);
    for (int i = 0; i < kNfft; ++i) win[i] = (float)(0.5 - 0.5 * cos(2 * PI * i / kNfft));
    tw400.resize(25 * kK1);
    for (int n2 = 0; n2 < 25; ++n2)
        for (int k1 = 0; k1 < kK1; ++k1) {
            const double a = -2 * PI * n2 * k1 / 400.0;
            tw400[n2 * kK1 + k1] = {(float)cos(a), (float)sin(a)};
        }
    tw25.resize(25);
    for (int b = 0; b < 5; ++b)
        for (int c = 0; c < 5; ++c) {
            const double a = -2 * PI * b * c / 25.0;
            tw25[b * 5 + c] = {(float)cos(a), (float)sin(a)};
        }
    std::vector<double> hz(n_mels + 2);
    const double m0 = hz_to_mel(0.0), m1 = hz_to_mel(8000.0);
    for (int i = 0; i < n_mels + 2; ++i) hz[i] = mel_to_hz(m0 + (m1 - m0) * i / (n_mels + 1));
    wts.assign(kMaxTaps * kMelColsPad, 0.f);
    start.assign(kMelColsPad, 0);
    for (int m = 0; m < n_mels; ++m) {
        const double enorm = 2.0 / (hz[m + 2] - hz[m]);
        int first = -1, cnt = 0;
        for (int k = 0; k < kBins; ++k) {
            const double f = 8000.0 * k / (kBins - 1);
            const double lower = (f - hz[m]) / (hz[m + 1] - hz[m]);
            const double upper = (hz[m + 2] - f) / (hz[m + 2] - hz[m + 1]);
            double w = lower < upper ? lower : upper;
            if (w > 0) {
                if (first < 0) first = k;
                const int j = k - first;
                if (j < kMaxTaps) wts[j * kMelColsPad + m] = (float)(w * enorm);
                ++cnt;
            }
        }
        start[m] = first < 0 ? 0 : first;
        (void)cnt;
    }
}


}  // namespace wk
