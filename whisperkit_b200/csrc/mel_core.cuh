// Arithmetic core of the fused log-mel kernel (K1), written as __host__ __device__ task functions so
// the exact same code is (a) run by the CUDA kernel in mel.cu and (b) replayed task-by-task on the CPU
// by tests/hostcheck (no GPU needed to validate indexing and numerics).
//
// Replaces the opaque MelSpectrogram.mlmodelc call of the reference
// (Sources/WhisperKit/Core/FeatureExtractor.swift:40-56); the arithmetic is OpenAI Whisper's log-mel:
// 400-point periodic-Hann STFT, hop 160, centred with reflect padding, |X|^2 over 201 bins, Slaney mel
// filterbank, log10(max(.,1e-10)), then (in pass 2) max(x, window_max - 8), (x + 4) / 4.
//
// The 400-point real DFT is a 16 x 25 Cooley-Tukey split done in registers:
//   n = 25*n1 + n2, k = k1 + 16*k2
//   phase 1 (task = frame, n2):  16-point real DFT over n1, keep k1 = 0..8, multiply by W400^(n2*k1)
//   phase 3 (task = frame, k1):  25-point complex DFT over n2 -> bins k1 + 16*k2 (and, by Hermitian
//                                symmetry, bins 400 - k for the columns k1 = 9..15 that were skipped)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define WK_HD __host__ __device__ __forceinline__
#else
#define WK_HD inline
#endif

namespace wk {
namespace mel {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kFramesPerWindow = 3000;
constexpr int kWindowSamples = 480000;
constexpr int kF = 32;                      // frames per CTA
constexpr int kThreads = 288;               // 9 warps: phase 3 has exactly 9*kF tasks
constexpr int kSamplesPerCta = (kF - 1) * kHop + kNfft;  // 5360
constexpr int kK1 = 9;                      // k1 = 0..8
constexpr int kYPerFrame = kK1 * 25;        // complex values per frame after phase 1
constexpr int kMaxTaps = 16;                // max non-zero bins per mel filter (14 for 80 mels, 9 for 128)
constexpr int kMelColsPad = 128;            // mel channels padded to 128 in every table and in the output rows
constexpr float kQScale = 4096.0f;          // pass-1 fixed point: q = (log10 + 10) * 4096, u16
constexpr float kQOffset = 10.0f;

struct cf {
    float x, y;
};
WK_HD cf cmul(cf a, cf b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
WK_HD cf cadd(cf a, cf b) { return {a.x + b.x, a.y + b.y}; }
WK_HD cf csub(cf a, cf b) { return {a.x - b.x, a.y - b.y}; }
// multiply by -i
WK_HD cf cmul_mi(cf a) { return {a.y, -a.x}; }

// index into the (zero-padded to 480000, then reflect-padded by 200) signal
WK_HD int reflect_index(int i) {
    if (i < 0) return -i;
    if (i >= kWindowSamples) return 2 * kWindowSamples - 2 - i;
    return i;
}

// ---- phase 1: 16-point real DFT (k1 = 0..8) of x[n1] = s[25*n1 + n2] * win[25*n1 + n2], twiddled by W400^(n2*k1)
// s: this frame's 400 samples; win: Hann[400]; tw: [25][9] complex W400^(n2*k1); y_out: this frame's [9][25] complex
WK_HD void phase1_task(const float* s, const float* win, const cf* tw, int n2, cf* y_out) {
    float x[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) x[n1] = s[25 * n1 + n2] * win[25 * n1 + n2];
    // n1 = 4a + b, k1 = c + 4d.  Stage A: 4-point DFT over a (real inputs) for each b.
    cf z[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const float x0 = x[b], x1 = x[4 + b], x2 = x[8 + b], x3 = x[12 + b];
        z[b][0] = {x0 + x1 + x2 + x3, 0.f};
        z[b][1] = {x0 - x2, -(x1 - x3)};
        z[b][2] = {x0 - x1 + x2 - x3, 0.f};
        z[b][3] = {x0 - x2, (x1 - x3)};
    }
    // twiddle W16^(b*c) = exp(-2*pi*i*b*c/16)
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
    const cf w16[10] = {{1.f, 0.f}, {c1, -s1}, {r2, -r2}, {s1, -c1}, {0.f, -1.f},
                        {-s1, -c1}, {-r2, -r2}, {-c1, -s1}, {-1.f, 0.f}, {-c1, s1}};
#pragma unroll
    for (int b = 1; b < 4; ++b)
#pragma unroll
        for (int c = 1; c < 4; ++c) z[b][c] = cmul(z[b][c], w16[b * c]);
    // Stage B: 4-point DFT over b for each c; X[c + 4d]
    cf X[9];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const cf t0 = z[0][c], t1 = z[1][c], t2 = z[2][c], t3 = z[3][c];
        const cf e = cadd(t0, t2), o = cadd(t1, t3), f = csub(t0, t2), g = cmul_mi(csub(t1, t3));
        X[c] = cadd(e, o);          // d = 0
        X[c + 4] = cadd(f, g);      // d = 1
        if (c == 0) X[8] = csub(e, o);  // d = 2
    }
#pragma unroll
    for (int k1 = 0; k1 < kK1; ++k1) y_out[k1 * 25 + n2] = cmul(X[k1], tw[n2 * kK1 + k1]);
}

// 5-point DFT (forward, exp(-2*pi*i/5)) in place on v[0..4]
WK_HD void dft5(cf* v) {
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    const cf x0 = v[0];
    const cf t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    const cf a1 = {x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y};
    const cf a2 = {x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y};
    const cf b1 = {s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y};
    const cf b2 = {s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y};
    v[0] = {x0.x + t1.x + t2.x, x0.y + t1.y + t2.y};
    // y1 = a1 - i*b1, y4 = a1 + i*b1, y2 = a2 - i*b2, y3 = a2 + i*b2   (-i*b = (b.y, -b.x))
    v[1] = {a1.x + b1.y, a1.y - b1.x};
    v[4] = {a1.x - b1.y, a1.y + b1.x};
    v[2] = {a2.x + b2.y, a2.y - b2.x};
    v[3] = {a2.x - b2.y, a2.y + b2.x};
}

// ---- phase 3: 25-point complex DFT over n2 of y[k1][n2]; writes |X|^2 into p_out[201]
// tw25: [5][5] complex W25^(b*c)
WK_HD void phase3_task(const cf* y, const cf* tw25, int k1, float* p_out) {
    cf z[5][5];  // [b][a] then [b][c]
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        cf v[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) v[a] = y[k1 * 25 + 5 * a + b];
        dft5(v);
#pragma unroll
        for (int c = 0; c < 5; ++c) z[b][c] = (b > 0 && c > 0) ? cmul(v[c], tw25[b * 5 + c]) : v[c];
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        cf v[5];
#pragma unroll
        for (int b = 0; b < 5; ++b) v[b] = z[b][c];
        dft5(v);
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            const int k2 = c + 5 * d;
            const int k = k1 + 16 * k2;
            const float pw = v[d].x * v[d].x + v[d].y * v[d].y;
            if (k <= 200) p_out[k] = pw;
            else if (k1 >= 1 && k1 <= 7) p_out[400 - k] = pw;  // Hermitian mirror covers columns 9..15
        }
    }
}

// ---- mel + log10 for one (frame, mel) pair; returns the u16 fixed-point code
// wts: [kMaxTaps][n_mels_ld] (tap-major), start[m]
WK_HD uint32_t mel_task(const float* p, const float* wts, const int* start, int m, int n_mels_ld) {
    const int s0 = start[m];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxTaps; ++j) {
        int k = s0 + j;
        k = k > 200 ? 200 : k;
        acc = fmaf(wts[j * n_mels_ld + m], p[k], acc);
    }
    acc = acc < 1e-10f ? 1e-10f : acc;
#if defined(__CUDA_ARCH__)
    const float lg = __log2f(acc) * 0.30102999566398120f;
#else
    const float lg = log10f(acc);
#endif
    float q = (lg + kQOffset) * kQScale + 0.5f;
    q = q < 0.f ? 0.f : (q > 65535.f ? 65535.f : q);
    return (uint32_t)q;
}

// pass 2: code -> normalised log-mel
WK_HD float mel_normalise(uint32_t q, uint32_t qmax) {
    const float x = (float)q * (1.0f / kQScale) - kQOffset;
    const float xm = (float)qmax * (1.0f / kQScale) - kQOffset - 8.0f;
    const float v = x > xm ? x : xm;
    return (v + 4.0f) * 0.25f;
}

}  // namespace mel
}  // namespace wk
