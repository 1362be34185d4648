// Host-side (double precision) construction of the MFCC kernel's constant tables and of its
// shared-memory / table-blob layout.  Pure C++ (no CUDA calls): used by libdspfe.so at plan creation
// and by the CPU-only tests.  Formulas follow reference base.py:34-58 (mel filterbank),
// base.py:12-13 (DCT-II ortho), base.py:60-68 (lifter), base.py:74 (delta denominator).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "dspfe_types.h"

namespace dspfe {

struct MfccConfig {
    int samplerate = 16000;
    int frame_len = 400;     // samples (caller applies round_half_up(winlen*samplerate), sigproc.py:77)
    int frame_step = 160;
    int nfft = 512;
    int nfilt = 26;
    int numcep = 13;
    int ceplifter = 22;
    int append_energy = 1;
    int delta_n = 2;
    int seg_frames = 256;    // tile length in output frames
    double preemph = 0.97;
    double lowfreq = 0.0;
    double highfreq = 0.0;   // <= 0 means samplerate/2
    std::vector<double> window;  // empty = rectangular (reference default winfunc)
};

inline double hz2mel(double hz) { return 2595.0 * std::log10(1.0 + hz / 700.0); }
inline double mel2hz(double mel) { return 700.0 * (std::pow(10.0, mel / 2595.0) - 1.0); }

// FFT-bin edges of the triangular filters: floor((nfft+1)*mel2hz(linspace)/samplerate), base.py:44-49
inline std::vector<double> mel_bin_edges(const MfccConfig& c) {
    const double high = c.highfreq > 0 ? c.highfreq : c.samplerate / 2.0;
    const double lowmel = hz2mel(c.lowfreq), highmel = hz2mel(high);
    const int n = c.nfilt + 2;
    std::vector<double> b(n);
    const double step = (highmel - lowmel) / (n - 1);  // numpy.linspace: arange(num)*step + start, last := stop
    for (int i = 0; i < n; ++i) {
        const double mel = (i == n - 1) ? highmel : i * step + lowmel;
        b[i] = std::floor((c.nfft + 1) * mel2hz(mel) / c.samplerate);
    }
    return b;
}

// nfft = 1536 (the long-frame kernel K1L, mfcc_long_kernel.cuh)
inline std::string mfcc_long_config_check(const MfccConfig& c) {
    if (c.nfft != 1536) return "nfft must be 512 or 1536 in this build";
    if (c.frame_len < 1 || c.frame_len > c.nfft) return "frame_len must be in [1, nfft] (longer frames are truncated by the reference with a warning; not supported)";
    if (c.frame_step < 1) return "frame_step must be >= 1";
    if (c.nfilt < 1 || c.nfilt > kMaxNfilt) return "nfilt must be in [1, 40]";
    if (c.numcep < 1 || c.numcep > kMaxNumcep || c.numcep > c.nfilt) return "numcep must be in [1, min(16, nfilt)]";
    if (c.delta_n < 1 || c.delta_n > kMaxDeltaN) return "delta N must be in [1, 4]";
    const double high = c.highfreq > 0 ? c.highfreq : c.samplerate / 2.0;
    if (high > c.samplerate / 2.0) return "highfreq is greater than samplerate/2";
    if (c.lowfreq < 0 || c.lowfreq >= high) return "lowfreq must be in [0, highfreq)";
    if (!c.window.empty() && (int)c.window.size() != c.frame_len) return "window length must equal frame_len";
    return "";
}

inline std::string mfcc_config_check(const MfccConfig& c) {
    if (c.nfft != kNfft) return "nfft must be 512 in this build (other sizes: SURVEY f-2)";
    if (c.frame_len < 1 || c.frame_len > c.nfft) return "frame_len must be in [1, nfft] (longer frames are truncated by the reference with a warning; not supported)";
    if (c.frame_step < 2 || (c.frame_step & 1)) return "frame_step must be even and >= 2";
    if (c.nfilt < 1 || c.nfilt > kMaxNfilt) return "nfilt must be in [1, 40]";
    if (c.numcep < 1 || c.numcep > kMaxNumcep || c.numcep > c.nfilt) return "numcep must be in [1, min(16, nfilt)]";
    if (c.delta_n < 1 || c.delta_n > kMaxDeltaN) return "delta N must be in [1, 4]";
    if (c.seg_frames < 16 || c.seg_frames > 1024) return "seg_frames must be in [16, 1024]";
    const double high = c.highfreq > 0 ? c.highfreq : c.samplerate / 2.0;
    if (high > c.samplerate / 2.0) return "highfreq is greater than samplerate/2";
    if (c.lowfreq < 0 || c.lowfreq >= high) return "lowfreq must be in [0, highfreq)";
    if (!c.window.empty() && (int)c.window.size() != c.frame_len) return "window length must equal frame_len";
    return "";
}

// Fills the layout fields of `p` (table offsets, shared-memory carve-up, scalars) and returns the
// table blob.  Pointer fields of `p` are left untouched.
// `in_f32` selects the shared-memory layout for float32 input samples (a 4-byte raw staging area).
inline std::vector<float> build_mfcc_tables(const MfccConfig& c, MfccParams& p, std::string& err, bool in_f32 = false) {
    err = mfcc_config_check(c);
    std::vector<float> blob;
    if (!err.empty()) return blob;
    const double kPi = 3.14159265358979323846;
    p.frame_len = c.frame_len; p.frame_step = c.frame_step; p.nfilt = c.nfilt; p.numcep = c.numcep;
    p.delta_n = c.delta_n; p.seg_frames = c.seg_frames; p.append_energy = c.append_energy;
    p.preemph = (float)c.preemph;
    int den = 0; for (int i = 1; i <= c.delta_n; ++i) den += i * i;
    p.delta_scale = (float)(1.0 / (2.0 * den));
    p.pow_scale = (float)(1.0 / (4.0 * c.nfft));
    p.nrange = c.nfilt + 3;

    auto align4 = [&]() { while (blob.size() & 3) blob.push_back(0.f); };
    // W256^{n2*k1} at [k1*16 + n2]
    p.o_twa = (int)blob.size();
    for (int k1 = 0; k1 < 16; ++k1)
        for (int n2 = 0; n2 < 16; ++n2) {
            const double a = -2.0 * kPi * (double)(n2 * k1) / 256.0;
            blob.push_back((float)std::cos(a)); blob.push_back((float)std::sin(a));
        }
    // W512^k for the split post-pass, k < 144
    p.o_twp = (int)blob.size();
    for (int k = 0; k < 144; ++k) {
        const double a = -2.0 * kPi * (double)k / 512.0;
        blob.push_back((float)std::cos(a)); blob.push_back((float)std::sin(a));
    }
    const std::vector<double> bins = mel_bin_edges(c);
    std::vector<int> edge(p.nrange + 1);
    edge[0] = 0;
    for (int i = 0; i < c.nfilt + 2; ++i) edge[i + 1] = std::min(std::max((int)bins[i], 0), kBins);
    edge[p.nrange] = kBins;
    for (int i = 1; i <= p.nrange; ++i) edge[i] = std::max(edge[i], edge[i - 1]);
    // Mel ranges = the intervals between consecutive filter centres (plus the two unfiltered ends).  Inside
    // range [lo, hi): rising weight of filter j = (k-lo)/(hi-lo), falling weight of filter j-1 = (hi-k)/(hi-lo)
    // (reference base.py:52-57).  Ranges longer than 16 bins are split into sub-ranges so that the 16 lanes of
    // a group get balanced work; each sub-range starts its weight counters at (sub_lo - lo, hi - sub_lo).
    struct Sub { int lo, len; float fi0, gi0, inv; int range; };
    std::vector<Sub> subs;
    std::vector<int> rsub(2 * kMaxRanges, 0);
    for (int i = 0; i < p.nrange; ++i) {
        const int lo = edge[i], hi = edge[i + 1];
        rsub[2 * i] = (int)subs.size();
        if (hi > lo) {
            if (i >= 1 && i <= c.nfilt + 1 && (bins[i - 1] != (double)lo || bins[i] != (double)hi)) {
                err = "mel bin edges fall outside [0, nfft/2]"; return {};
            }
            const int parts = (hi - lo + 15) / 16;
            for (int q = 0; q < parts; ++q) {
                const int a = lo + (int)((int64_t)(hi - lo) * q / parts), b = lo + (int)((int64_t)(hi - lo) * (q + 1) / parts);
                subs.push_back({a, b - a, (float)(a - lo), (float)(hi - a), (float)(1.0 / (hi - lo)), i});
            }
        }
        rsub[2 * i + 1] = (int)subs.size() - rsub[2 * i];
    }
    if ((int)subs.size() > kMaxSubs) { err = "too many mel sub-ranges"; return {}; }
    p.o_sub = (int)blob.size();
    for (int i = 0; i < kMaxSubs; ++i) {
        Sub s = i < (int)subs.size() ? subs[i] : Sub{0, 0, 0.f, 0.f, 0.f, 0};
        float f; std::memcpy(&f, &s.lo, 4); blob.push_back(f);
        std::memcpy(&f, &s.len, 4); blob.push_back(f);
        blob.push_back(s.fi0); blob.push_back(s.gi0); blob.push_back(s.inv);
    }
    p.o_rsub = (int)blob.size();
    for (int v : rsub) { float f; std::memcpy(&f, &v, 4); blob.push_back(f); }
    // longest-first assignment of the sub-ranges to the 16 lanes of a group
    p.o_task = (int)blob.size();
    {
        std::vector<int> order;
        for (int i = 0; i < (int)subs.size(); ++i) order.push_back(i);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return subs[a].len > subs[b].len; });
        std::vector<int> t(kGroupLanes * kMaxTasks, -1), load(kGroupLanes, 0), cnt(kGroupLanes, 0);
        for (int si : order) {
            int best = -1;
            for (int l = 0; l < kGroupLanes; ++l)
                if (cnt[l] < kMaxTasks && (best < 0 || load[l] < load[best])) best = l;
            if (best < 0) { err = "mel sub-range assignment overflow"; return {}; }
            t[best * kMaxTasks + cnt[best]++] = si;
            load[best] += subs[si].len;
        }
        for (int v : t) { float f; std::memcpy(&f, &v, 4); blob.push_back(f); }
    }
    align4();
    // DCT-II (ortho) rows premultiplied by the lifter
    p.o_dct = (int)blob.size();
    p.dct_stride = c.nfilt; while ((p.dct_stride & 3) != 2) ++p.dct_stride;   // float2 rows, conflict-free over lanes
    for (int k = 0; k < c.numcep; ++k) {
        const double s = std::sqrt((k == 0 ? 1.0 : 2.0) / c.nfilt);
        const double lift = c.ceplifter > 0 ? 1.0 + (c.ceplifter / 2.0) * std::sin(kPi * k / c.ceplifter) : 1.0;
        for (int m = 0; m < p.dct_stride; ++m)
            blob.push_back(m < c.nfilt ? (float)(lift * s * std::cos(kPi * k * (2 * m + 1) / (2.0 * c.nfilt))) : 0.f);
    }
    align4();
    p.o_win = (int)blob.size();
    if (!c.window.empty()) {   // two zero-padded planes: even-index and odd-index window samples
        for (int q = 0; q < 2; ++q)
            for (int n = 0; n < 256; ++n) { const int i = 2 * n + q; blob.push_back(i < c.frame_len ? (float)c.window[i] : 0.f); }
    }
    align4();
    p.tbl_floats = (int)blob.size();

    // shared-memory carve-up
    auto up16 = [](int v) { return (v + 15) & ~15; };
    int off = up16(p.tbl_floats * 4);
    p.sm_mbar = off; off += 16;
    // FFT scratch; the epilogue reuses it for the delta-delta rows of the tile
    p.sm_scratch = off; off += std::max(kMfccGroups * kScratchUnits * 8, up16(c.seg_frames * c.numcep * 4));
    p.sm_mfcc = off; off += up16((c.seg_frames + 4 * c.delta_n) * c.numcep * 4);
    p.fbuf_floats = (kFramesPerPass - 1) * c.frame_step + c.frame_len;
    // raw staging: the chunk's samples + one history sample + up to 7 samples of 16-byte alignment slack either side;
    // fbuf mirrors raw index for index
    p.fbuf_vecs = (p.fbuf_floats + 1 + 14 + 7) / 8;
    const int fbuf_bytes = p.fbuf_vecs * 8 * 4;
    const int raw_bytes = p.fbuf_vecs * 8 * (in_f32 ? 4 : 2);
    const int dbuf_bytes = up16((c.seg_frames + 2 * c.delta_n) * c.numcep * 4);
    p.sm_fbuf = off;
    p.sm_raw = off + fbuf_bytes;
    off += std::max(fbuf_bytes + raw_bytes, dbuf_bytes);
    p.sm_total = off;
    if (p.sm_total > 227 * 1024) err = "configuration needs more than 227 KB of shared memory per CTA";
    return blob;
}

// Upper bound on the number of tiles for a packed batch (grid size of the MFCC kernel).
inline int64_t mfcc_max_tiles(int64_t total_samples, int64_t n_utt, int frame_step, int seg_frames) {
    const int64_t max_frames = total_samples / frame_step + n_utt;
    return n_utt + max_frames / seg_frames + 1;
}

}  // namespace dspfe
