// Host-side (double precision) construction of the MFCC kernel's constant tables and of its
// shared-memory / table-blob layout.  Pure C++ (no CUDA calls): used by libdspfe.so at plan creation
// and by the CPU-only tests.  Formulas follow reference base.py:34-58 (mel filterbank),
// base.py:12-13 (DCT-II ortho), base.py:60-68 (lifter), base.py:74 (delta denominator).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "dspfe_types.h"

namespace dspfe {

struct MfccConfig {
    int samplerate = 16000;
    int frame_len = 400;     // samples (caller applies round_half_up(winlen*samplerate), sigproc.py:77)
    int frame_step = 160;
    int count_len = 0;       // frame length framesig counts and pads with (sigproc.py:79-87); > frame_len when the caller's frames are longer
                             // than nfft: the reference then transforms the first nfft samples of each frame (sigproc.py:143-147)
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

// every transform size but K1's (the long-frame / general kernel K1L, mfcc_long_kernel.cuh)
inline std::string mfcc_long_config_check(const MfccConfig& c) {
    if (!(c.nfft == 32 || c.nfft == 64 || c.nfft == 128 || c.nfft == 256 || c.nfft == 512 || c.nfft == 1024 || c.nfft == 1536 || c.nfft == 2048))
        return "nfft must be a power of two in [32, 2048] or 1536 in this build";
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
    if (c.nfft != kNfft && c.nfft != kTriNfft) return "the tiled kernel is built for nfft 512 and 1536 (other sizes: the general kernel)";
    if (c.frame_len < 1 || c.frame_len > c.nfft) return "frame_len must be in [1, nfft] (longer frames are truncated by the reference with a warning; not supported)";
    if (c.frame_step < 1) return "frame_step must be >= 1";
    if (c.frame_step > (c.count_len > 0 ? c.count_len : c.frame_len)) return "frame_step greater than frame_len (gaps between frames) is not built: the row bounds assume overlapping or abutting frames";
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
    // K1T input twiddles W768^{m r} (decimation in frequency by three) and the split twiddles of the class-2 bins
    const bool tri = c.nfft == kTriNfft;
    p.tri = tri ? 1 : 0;
    p.grp_units = kScratchUnits;
    if (tri) {
        align4();
        p.o_tw3 = (int)blob.size();      // [m] (cos a1, cos a2, sin a1, sin a2), a_r = -2 pi m r / 768: the classes 1 and 2 ride in one packed transform
        for (int m = 0; m < 256; ++m) {
            const double a1 = -2.0 * kPi * (double)m / 768.0, a2 = -2.0 * kPi * (double)(2 * m) / 768.0;
            blob.push_back((float)std::cos(a1)); blob.push_back((float)std::cos(a2));
            blob.push_back((float)std::sin(a1)); blob.push_back((float)std::sin(a2));
        }
        p.o_tws2 = (int)blob.size();
        for (int q = 0; q < 256; ++q) {
            const double a = -2.0 * kPi * (double)(3 * q + 2) / 1536.0;
            blob.push_back((float)std::cos(a)); blob.push_back((float)std::sin(a));
        }
    }
    // Mel filterbank (reference base.py:40-58) as balanced *pieces*: filter j has non-zero weights on one contiguous
    // run of bins; the runs are cut into pieces and every (slot, lane) of a 16-lane group owns at most one piece.
    // All lanes run the same mel_T[slot] iterations per slot (zero weights pad short pieces), so the kernel's loop
    // has a uniform trip count, no per-lane branching and reads its weights from a [iteration][lane] table.
    // K1T keeps the power bins by residue class, bins 3q + c indexed by q: class 0 as one array of (frame A, frame B) values, the
    // classes 1 and 2 of ONE frame as an array of (class 1, class 2) values.  A filter's weights on a class are again a contiguous
    // run in q, so class 0 gets a piece table like K1's and the classes 1 / 2 a joint one with a weight PAIR per entry; the partial
    // sums meet in the combine step.
    const int nbins = c.nfft / 2 + 1, ncls = tri ? 2 : 1, nres = tri ? 3 : 1;
    const std::vector<double> bins = mel_bin_edges(c);
    for (int i = 0; i < c.nfilt + 2; ++i)
        if (bins[i] < 0 || bins[i] > c.nfft / 2) { err = "mel bin edges fall outside [0, nfft/2]"; return {}; }
    std::vector<std::vector<double>> dense(c.nfilt, std::vector<double>(nbins, 0.0));
    for (int j = 0; j < c.nfilt; ++j) {
        const int lo = (int)bins[j], ce = (int)bins[j + 1], hi = (int)bins[j + 2];
        for (int k = lo; k < ce; ++k) dense[j][k] = (k - bins[j]) / (bins[j + 1] - bins[j]);
        for (int k = ce; k < hi; ++k) dense[j][k] = (bins[j + 2] - k) / (bins[j + 2] - bins[j + 1]);
    }
    struct Run { int k0; std::vector<float> w, w2; };
    struct Piece { int filt, off, len, slot, lane, k0; };
    const int zero_part = ncls * kMelSlots * kGroupLanes;      // index of the always-zero partial sum
    std::vector<int> comb((size_t)kMaxNfilt * kMelMaxPieces * ncls, zero_part);
    int npc[kMaxNfilt] = {0};
    for (int cls = 0; cls < ncls; ++cls) {
    const bool joint = cls == 1;                               // K1T: residues 1 and 2 together
    const int nq = joint ? 256 : (nbins + nres - 1) / nres;    // entries of the class array
    const int units = kScratchUnits;                           // 8-byte units a piece may read (zero padded past nq)
    std::vector<Run> runs(c.nfilt);
    for (int j = 0; j < c.nfilt; ++j) {
        auto wgt = [&](int q, int res) { return dense[j][nres * q + res]; };
        auto nz = [&](int q) { return joint ? (wgt(q, 1) != 0.0 || wgt(q, 2) != 0.0) : wgt(q, 0) != 0.0; };
        int a = 0, b = nq;
        while (a < nq && !nz(a)) ++a;
        while (b > a && !nz(b - 1)) --b;
        runs[j].k0 = a < nq ? a : 0;
        for (int k = a; k < b; ++k) { runs[j].w.push_back((float)wgt(k, joint ? 1 : 0)); if (joint) runs[j].w2.push_back((float)wgt(k, 2)); }
    }
    std::vector<Piece> pieces;
    int mel_T[kMelSlots] = {0, 0, 0};
    {
        int nnz = 0, lmax = 0;
        for (auto& r : runs) { nnz += (int)r.w.size(); lmax = std::max(lmax, (int)r.w.size()); }
        // Lanes of a slot: lane l may take a piece whose first bin is k0 if it starts d = (k0 - l) mod 16 bins early
        // (zero weights in front) -- then the 16 lanes of a group read 16 different 8-byte bank pairs in every iteration
        // (no shared-memory bank conflicts).  Bipartite matching, augmenting paths.
        bool allow_conflicts = false;
        auto assign_lanes = [&](std::vector<Piece>& out, int s, int T) {
            std::vector<int> idx;
            for (int i = 0; i < (int)out.size(); ++i) if (out[i].slot == s) idx.push_back(i);
            int owner[kGroupLanes]; for (int l = 0; l < kGroupLanes; ++l) owner[l] = -1;
            auto ok = [&](int i, int l) {
                const int k0 = runs[out[i].filt].k0 + out[i].off, d = ((k0 - l) % kGroupLanes + kGroupLanes) % kGroupLanes;
                return k0 - d >= 0 && d + out[i].len <= T && k0 - d + T <= units;
            };
            bool seen[kGroupLanes];
            std::function<bool(int)> aug = [&](int i) {
                for (int l = 0; l < kGroupLanes; ++l) {
                    if (seen[l] || !ok(i, l)) continue;
                    seen[l] = true;
                    if (owner[l] < 0 || aug(owner[l])) { owner[l] = i; return true; }
                }
                return false;
            };
            bool matched = true;
            for (int i : idx) { for (bool& b : seen) b = false; if (!aug(i)) { matched = false; break; } }
            if (!matched) {
                if (!allow_conflicts) return false;
                int l = 0;   // last resort (degenerate filterbanks): natural starts, bank conflicts accepted
                for (int i : idx) {
                    const int k0 = runs[out[i].filt].k0 + out[i].off;
                    out[i].lane = l++; out[i].k0 = std::max(0, std::min(k0, units - T));
                    if (k0 - out[i].k0 + out[i].len > T) return false;
                }
                return true;
            }
            for (int l = 0; l < kGroupLanes; ++l) if (owner[l] >= 0) {
                Piece& pc = out[owner[l]];
                const int k0 = runs[pc.filt].k0 + pc.off;
                pc.lane = l; pc.k0 = k0 - ((k0 - l) % kGroupLanes + kGroupLanes) % kGroupLanes;
            }
            return true;
        };
        auto try_pack = [&](const int* T, int slack, std::vector<Piece>& out) {
            int freec[kMelSlots];
            for (int s = 0; s < kMelSlots; ++s) freec[s] = T[s] - slack > 0 ? kGroupLanes : 0;
            std::vector<int> order(c.nfilt);
            for (int j = 0; j < c.nfilt; ++j) order[j] = j;
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return runs[x].w.size() > runs[y].w.size(); });
            out.clear();
            for (int j : order) {
                int rem = (int)runs[j].w.size(), off = 0, np = 0;
                while (rem > 0) {
                    if (np == kMelMaxPieces) return false;
                    int pick = -1;   // tightest free slot that holds the rest, else the largest free slot (T is sorted descending)
                    for (int s = 0; s < kMelSlots; ++s) if (freec[s] > 0 && T[s] - slack >= rem) pick = s;
                    if (pick < 0) for (int s = 0; s < kMelSlots; ++s) if (freec[s] > 0) { pick = s; break; }
                    if (pick < 0) return false;
                    const int take = std::min(rem, T[pick] - slack);
                    out.push_back({j, off, take, pick, 0, 0});
                    --freec[pick]; off += take; rem -= take; ++np;
                }
            }
            for (int s = 0; s < kMelSlots; ++s) if (!assign_lanes(out, s, T[s])) return false;
            return true;
        };
        // iteration counts are even (the kernel consumes weight pairs); smallest total first
        bool ok = nnz == 0;
        const int bmax = kMelSlots * (lmax + kGroupLanes + 1);
        for (int B = ((nnz + kGroupLanes - 1) / kGroupLanes + 1) & ~1; !ok && B <= bmax; B += 2)
            for (int slack = 0; !ok && slack <= 3; ++slack)
                for (int t0 = B & ~1; !ok && 3 * t0 >= B; t0 -= 2)
                    for (int t1 = std::min(t0, B - t0); !ok && t1 >= 0 && 2 * t1 >= B - t0; t1 -= 2) {
                        const int T[kMelSlots] = {t0, t1, B - t0 - t1};
                        if (try_pack(T, slack, pieces)) { ok = true; for (int s = 0; s < kMelSlots; ++s) mel_T[s] = T[s]; }
                    }
        allow_conflicts = true;
        for (int B = ((nnz + kGroupLanes - 1) / kGroupLanes + 1) & ~1; !ok && B <= bmax; B += 2)
            for (int t0 = B & ~1; !ok && 3 * t0 >= B; t0 -= 2)
                for (int t1 = std::min(t0, B - t0); !ok && t1 >= 0 && 2 * t1 >= B - t0; t1 -= 2) {
                    const int T[kMelSlots] = {t0, t1, B - t0 - t1};
                    if (try_pack(T, 0, pieces)) { ok = true; for (int s = 0; s < kMelSlots; ++s) mel_T[s] = T[s]; }
                }
        if (!ok) { err = "mel filterbank does not fit the piece table"; return {}; }
    }
    int mel_iters = 0;
    for (int s = 0; s < kMelSlots; ++s) { p.mel_Tc[cls][s] = mel_T[s]; if (cls == 0) p.mel_T[s] = mel_T[s]; mel_iters += mel_T[s]; }
    {
        // weights as pairs: [iteration / 2][lane][iteration & 1]; the joint class: [iteration][lane] (class-1 weight, class-2 weight)
        std::vector<float> wt((size_t)std::max(mel_iters, 2) * kGroupLanes * (joint ? 2 : 1), 0.f);
        std::vector<int> base(kMelSlots * kGroupLanes, 0);
        for (int i = 0; i < kMelSlots * kGroupLanes; ++i) base[i] = i % kGroupLanes;   // idle lanes keep to their own banks
        int tb[kMelSlots];
        tb[0] = 0; for (int s = 1; s < kMelSlots; ++s) tb[s] = tb[s - 1] + mel_T[s - 1];
        for (const Piece& pc : pieces) {
            const int shift = runs[pc.filt].k0 + pc.off - pc.k0;
            base[pc.slot * kGroupLanes + pc.lane] = pc.k0;
            for (int t = 0; t < pc.len; ++t) {
                const int it = tb[pc.slot] + shift + t;
                if (joint) {
                    wt[((size_t)it * kGroupLanes + pc.lane) * 2] = runs[pc.filt].w[pc.off + t];
                    wt[((size_t)it * kGroupLanes + pc.lane) * 2 + 1] = runs[pc.filt].w2[pc.off + t];
                } else {
                    wt[((size_t)(it >> 1) * kGroupLanes + pc.lane) * 2 + (it & 1)] = runs[pc.filt].w[pc.off + t];
                }
            }
            comb[((size_t)pc.filt * ncls + cls) * kMelMaxPieces + npc[pc.filt]++] = cls * kMelSlots * kGroupLanes + pc.slot * kGroupLanes + pc.lane;
        }
        for (int j = 0; j < c.nfilt; ++j) npc[j] = 0;
        align4();
        p.o_melw_c[cls] = (int)blob.size();
        blob.insert(blob.end(), wt.begin(), wt.end());
        p.o_melb_c[cls] = (int)blob.size();
        for (int v : base) { float f; std::memcpy(&f, &v, 4); blob.push_back(f); }
    }
    }   // classes
    p.o_melw = p.o_melw_c[0]; p.o_melb = p.o_melb_c[0];
    align4();
    p.o_melc = (int)blob.size();      // [filter][class] int4: the partial sums that make the filter up
    for (int v : comb) { float f; std::memcpy(&f, &v, 4); blob.push_back(f); }
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
        const int plane = c.frame_len > 512 ? 768 : 256;          // (K1T LONG: all three thirds)
        for (int q = 0; q < 2; ++q)
            for (int n = 0; n < plane; ++n) { const int i = 2 * n + q; blob.push_back(i < c.frame_len ? (float)c.window[i] : 0.f); }
    }
    align4();
    p.tbl_floats = (int)blob.size();

    // shared-memory carve-up
    auto up16 = [](int v) { return (v + 15) & ~15; };
    int off = up16(p.tbl_floats * 4);
    p.sm_mbar = off; off += 16;
    // FFT scratch; the epilogue reuses it for the delta-delta rows of the tile
    p.sm_scratch = off; off += std::max(kMfccGroups * p.grp_units * 8, up16(c.seg_frames * c.numcep * 4));
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
