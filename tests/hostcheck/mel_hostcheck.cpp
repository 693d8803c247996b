// CPU replay of the log-mel kernel's task functions (test infrastructure).
// Runs exactly the __host__ __device__ code of whisperkit_b200/csrc/mel_core.cuh in the same
// phase order and with the same shared-memory aliasing as mel_pass1_kernel / mel_pass2_kernel,
// so indexing and numerics can be validated against the oracle without a GPU.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../whisperkit_b200/csrc/mel_tables.h"

using namespace wk;
using namespace wk::mel;

extern "C" int wk_hostcheck_mel(const float* pcm, int n_valid, int n_mels, float* out /*[n_mels][3000]*/) {
    std::vector<float> win, wts;
    std::vector<cf> tw400, tw25;
    std::vector<int> start;
    mel_host_tables(n_mels, win, tw400, tw25, wts, start);
    std::vector<uint16_t> codes((size_t)kFramesPerWindow * 128, 0);
    int gmax = 0;
    const int region = std::max(kF * kBins, kSamplesPerCta);
    std::vector<float> a(region);
    std::vector<cf> y(kF * kYPerFrame);
    for (int f0 = 0; f0 < kFramesPerWindow; f0 += kF) {
        const int nf = std::min(kF, kFramesPerWindow - f0);
        const int i0 = f0 * kHop - kNfft / 2;
        const int ns = (nf - 1) * kHop + kNfft;
        for (int i = 0; i < ns; ++i) {
            const int j = reflect_index(i0 + i);
            a[i] = j < n_valid ? pcm[j] : 0.f;
        }
        for (int task = 0; task < nf * 25; ++task) {
            const int f = task / 25, n2 = task - f * 25;
            phase1_task(a.data() + f * kHop, win.data(), tw400.data(), n2, y.data() + f * kYPerFrame);
        }
        for (int task = 0; task < nf * kK1; ++task) {
            const int f = task / kK1, k1 = task - f * kK1;
            phase3_task(y.data() + f * kYPerFrame, tw25.data(), k1, a.data() + f * kBins);
        }
        for (int task = 0; task < nf * n_mels; ++task) {
            const int f = task / n_mels, m = task - f * n_mels;
            const uint32_t q = mel_task(a.data() + f * kBins, wts.data(), start.data(), m, 128);
            gmax = std::max(gmax, (int)q);
            codes[(size_t)(f0 + f) * 128 + m] = (uint16_t)q;
        }
    }
    for (int f = 0; f < kFramesPerWindow; ++f)
        for (int m = 0; m < n_mels; ++m) out[(size_t)m * kFramesPerWindow + f] = mel_normalise(codes[(size_t)f * 128 + m], gmax);
    return 0;
}
