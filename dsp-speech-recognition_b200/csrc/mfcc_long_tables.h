// Host-side constant tables of the long-frame MFCC kernel K1L (shared by libdspfe.so and the test emulator).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "mfcc_long_kernel.cuh"
#include "mfcc_tables.h"
#include "pitch_tables.h"

namespace dspfe {

// host-side tables of K1L (float64, rounded once)
inline std::vector<float> build_long_tables(const MfccConfig& c) {
    const double kPi = 3.14159265358979323846;
    std::vector<float> t(kLtTotal, 0.f);
    fill_fft512_twiddles(reinterpret_cast<float2*>(t.data() + kLtTw), reinterpret_cast<float2*>(t.data() + kLtW32));
    for (int k = 0; k < c.nfft; ++k) { t[kLtW1536 + 2 * k] = (float)cos(-2 * kPi * k / c.nfft); t[kLtW1536 + 2 * k + 1] = (float)sin(-2 * kPi * k / c.nfft); }
    for (int n = 0; n < c.frame_len; ++n) t[kLtWin + n] = c.window.empty() ? 1.f : (float)c.window[n];
    const std::vector<double> bins = mel_bin_edges(c);            // floor((nfft+1) * mel2hz(.) / samplerate), base.py:49
    int32_t* edge = reinterpret_cast<int32_t*>(t.data() + kLtEdge);
    for (int i = 0; i < c.nfilt + 2; ++i) edge[i] = std::min(std::max((int)bins[i], 0), c.nfft / 2 + 1);
    for (int j = 0; j < c.nfilt; ++j) {
        t[kLtInvUp + j] = edge[j + 1] > edge[j] ? (float)(1.0 / (bins[j + 1] - bins[j])) : 0.f;
        t[kLtInvDn + j] = edge[j + 2] > edge[j + 1] ? (float)(1.0 / (bins[j + 2] - bins[j + 1])) : 0.f;
    }
    for (int q = 0; q < c.numcep; ++q) {
        const double s = q == 0 ? sqrt(1.0 / c.nfilt) : sqrt(2.0 / c.nfilt);                       // DCT-II ortho (base.py:13)
        const double lift = c.ceplifter > 0 ? 1.0 + (c.ceplifter / 2.0) * sin(kPi * q / c.ceplifter) : 1.0;   // base.py:64-68
        for (int j = 0; j < c.nfilt; ++j) t[kLtDct + q * 48 + j] = (float)(lift * s * cos(kPi * q * (2 * j + 1) / (2.0 * c.nfilt)));
    }
    return t;
}

}  // namespace dspfe
