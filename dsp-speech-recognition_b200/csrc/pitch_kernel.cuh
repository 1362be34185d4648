// K4/K5/K6: cepstrum pitch, autocorrelation pitch and the pitch-feature tail over a ragged batch (sm_100a).
//
// Replaces reference features/pitch.py: pitch_detect (:83), pitch_detect_sr (:96), pitch_detect_frame (:135),
// pitch_detect_frame_sr (:112), center_clip (:145), smooth (:157), peak_score (:227), max_pitch (:166),
// robust_max_pitch (:191), pitch_feature (:26) with sub_endpoint_detect (:64), find_smooth_subsequence (:245),
// slope/quad_params/peakshift (:49-62); plus preprocess.downsampling (:21) and sigproc.window (:22), acr (:48).
//
//   K4a/K5a  one warp per 10 kHz frame: gather the frame through the sample-picking decimator (no filter, as
//            the reference), centre-clip at the median of its non-negative samples (exact bitwise selection),
//            apply the one-sided (complex) FIR band-pass as a 1024-point FFT product, then
//              cepstrum:        FFT512 -> log|.| -> IFFT512 -> |.|            -> row of row_len (<= 512) columns
//              autocorrelation: |y| -> FFT1024 -> |.|^2 -> IFFT1024 -> /(L-n) -> row of 180 lags (20..199)
//   K4b/K5b  one CTA per utterance walks its frames in order, 16 at a time: the reference's in-place running
//            mean (a recurrence over already-smoothed rows, one column per thread), then peak scoring / argmax
//            for the 16 rows in parallel, finally the octave-repair sweeps.
//   K6       one thread per utterance: valley split, longest smooth runs, least-squares slope / curvature and
//            median shift -> the five SVM inputs of pitch_model.py.
#pragma once
#include <math.h>
#include <stdint.h>

#include "simt.h"

#ifndef DSP_HD
#ifdef __CUDACC__
#define DSP_HD __host__ __device__ inline
#else
#define DSP_HD inline
#endif
#endif

namespace dspfe {

constexpr int kPitchFft = 1024;
constexpr int kCepLen = 512;        // cepstrum length (frame length of pitch_detect)
constexpr int kCepCols = 200;       // columns peak_score can reach: lags 20..99 look at most 99 samples either side
constexpr int kAcrLags = 180;       // lags 20..199 (pitch.py:125-129)
constexpr int kMinLag = 20;
constexpr int kPeakLags = 80;       // peak_score evaluates lags 20..99 (pitch.py:232)
constexpr int kPitchWarps = 4;      // frames per CTA in K4a/K5a
constexpr int kTrackThreads = 512;
constexpr int kTrackChunk = 16;     // frames smoothed per pass of K4b/K5b
constexpr int kMaxDsOut = 256;      // decimator pattern length limit

struct PitchParams {
    const void* pcm; int in_f32;         // packed samples: int16 or float32
    const int64_t* offsets;              // [U+1]
    const int32_t* trim;                 // optional [U,2] (left,right): pitch runs on sig[left:right]
    int n_utt;
    double preemph;                      // y[n] = x[n] - c x[n-1] over the WHOLE utterance before trimming (pitch_model.py:39-41); 0 = off
    int ds_in, ds_out;                   // decimator: output k >= 1 reads input ((k-1)/ds_out)*ds_in + ds_idx[(k-1)%ds_out]
    int32_t ds_idx[kMaxDsOut];
    int frame_len, frame_step;           // at the decimated rate: 512/100 (or 300/100 for the autocorrelation variant)
    int do_clip;                         // center_clip(frame, False) before the band-pass (pitch.py:88,103)
    int no_smooth;                       // taps only: K4b/K5b score the rows as given (peak_score on its own)
    const float2* tw;                    // W1024^k, k < 1024
    const float2* H;                     // FFT1024 of the FIR taps
    int mode;                            // 0 cepstrum, 1 autocorrelation
    int row_len;                         // columns per row: cepstrum 200 (fused) or 512 (tap); autocorrelation 180
    int64_t* frame_off;                  // [U+1] prefix sums of pitch frames
    int64_t* seg_start; int32_t* seg_len; int32_t* ds_len;   // per utterance: trimmed range and decimated length
    float* rows;                         // [F_total, row_len] raw rows (K4a/K5a output)
    float* rows_out;                     // optional: smoothed rows (tap of smooth, pitch.py:157)
    int32_t* score;                      // optional [F_total, 80]: peak_score of every smoothed row (tap)
    double* frame_amp;                   // [F_total] sum |x| of each raw frame (sub_endpoint_detect, pitch.py:65)
    double* pitch;                       // [F_total] Hz after robust_max_pitch
    int32_t* lag;                        // [F_total] 20 + argmax (before octave repair)
    double* feat;                        // [U,5] pitch_feature, or null
    double* scratch;                     // [3*F_total] K6 work area
    int64_t max_frames;
};

// ---------------------------------------------------------------------------------------------------------
// host/device scalar pieces
// ---------------------------------------------------------------------------------------------------------
// number of decimated samples of an S-sample signal (preprocess.py:21-28: sample 0 is always kept, sample i > 0
// is the k-th kept one when it is the first with i*dst/src > k-1)
DSP_HD int64_t ds_length(int64_t S, const int32_t* ds_idx, int ds_in, int ds_out) {
    if (S <= 0) return 0;
    const int64_t T = S - 1, q = T / ds_in;
    int64_t n = 1 + q * ds_out;
    for (int i = 0; i < ds_out; ++i) n += ((int64_t)ds_idx[i] <= T - q * ds_in) ? 1 : 0;
    return n;
}
DSP_HD int64_t ds_index(int64_t k, const int32_t* ds_idx, int ds_in, int ds_out) {
    if (k <= 0) return 0;
    const int64_t j = k - 1;
    return (j / ds_out) * ds_in + ds_idx[j % ds_out];
}

// robust_max_pitch (pitch.py:191-206) on lags: p = 1/(0.0001*lag), forward then backward octave repair
DSP_HD void robust_pitch(const int32_t* lag, int F, double* pitch) {
    for (int i = 0; i < F; ++i) pitch[i] = 1.0 / (0.0001 * (double)lag[i]);
    for (int i = 1; i < F; ++i)
        if (fabs(2 * pitch[i] - pitch[i - 1]) < 50 && pitch[i] < 170) pitch[i] = 2 * pitch[i];
    for (int i = F - 2; i > 0; --i)
        if (fabs(2 * pitch[i] - pitch[i + 1]) < 50 && pitch[i] < 170) pitch[i] = 2 * pitch[i];
}

// find_smooth_subsequence (pitch.py:245-279): longest run (first on ties) tolerating `tor` jumps > thres.
// Writes the run's accepted values to seg and returns its length; *i0/*j0 = the run's (start, stop) indices.
DSP_HD int smooth_run(const double* pitch, int n, int tor, double thres, double* seg, double* tmp, int* i0, int* j0) {
    int best = 0, i = 0;
    *i0 = 0; *j0 = 0;
    while (i < n) {
        int j = i + 1, k = tor, len = 1;
        double prev = pitch[i];
        tmp[0] = pitch[i];
        bool closed = false;
        while (j < n) {
            if (fabs(pitch[j] - prev) > thres) --k;
            else { tmp[len++] = pitch[j]; prev = pitch[j]; }
            if (!k) { closed = true; break; }
            ++j;
        }
        if (len > best) { best = len; *i0 = i; *j0 = j; for (int q = 0; q < len; ++q) seg[q] = tmp[q]; }
        if (!closed) break;      // reached the end of the sequence
        i = j - tor + 1;
    }
    return best;
}

// least-squares polynomial leading coefficients over x = 0..n-1 (np.polyfit(x, seq, deg)[0]); closed-form
// normal equations about the centred abscissa (odd moments vanish)
DSP_HD double ls_slope(const double* y, int n) {
    const double xm = 0.5 * (n - 1);
    double s2 = 0, t1 = 0;
    for (int i = 0; i < n; ++i) { const double x = i - xm; s2 += x * x; t1 += x * y[i]; }
    return t1 / s2;
}
DSP_HD double ls_quad(const double* y, int n) {
    const double xm = 0.5 * (n - 1);
    double s2 = 0, s4 = 0, t0 = 0, t2 = 0;
    for (int i = 0; i < n; ++i) { const double x = i - xm, x2 = x * x; s2 += x2; s4 += x2 * x2; t0 += y[i]; t2 += x2 * y[i]; }
    return (n * t2 - s2 * t0) / (n * s4 - s2 * s2);   // [s4 s2; s2 n] [a; c] = [t2; t0]
}
DSP_HD double median_inplace(double* a, int n) {   // np.median
    for (int i = 1; i < n; ++i) { double v = a[i]; int j = i - 1; while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; } a[j + 1] = v; }
    return (n & 1) ? a[n / 2] : 0.5 * (a[n / 2 - 1] + a[n / 2]);
}

// sub_endpoint_detect (pitch.py:64-81) on the per-frame sum |x|
DSP_HD int sub_endpoint(const double* amp, int F) {
    int p = 0; double max_score = -1000;
    for (int i = 10; i < F - 10; ++i) {
        bool lower = false;
        for (int j = i - 2; j <= i + 2; ++j) lower = lower || (amp[j] < amp[i]);
        if (lower) continue;
        double s = 0;
        for (int j = i - 10; j <= i + 10; ++j) s += amp[j] - amp[i];
        if (s > max_score) { max_score = s; p = i; }
    }
    return p == 0 ? F / 2 : p;
}

// pitch_feature (pitch.py:26-47) after pitch_detect: amp = sum|x| per raw frame, pitch = Hz per frame.
// work: 3*F doubles.  A half with fewer than three accepted values has no defined curvature: the five
// outputs are NaN and false is returned (the reference raises or returns a minimum-norm fit there).
DSP_HD bool pitch_feature_tail(const double* pitch, const double* amp, int F, double* work, double* out5) {
    const int p = sub_endpoint(amp, F);
    const int p_bias = p > 15 ? 5 : 0;
    double* s1 = work; double* s2 = work + F; double* tmp = work + 2 * F;
    int a, b;
    const int n1 = smooth_run(pitch + p_bias, p - p_bias, 3, 30.0, s1, tmp, &a, &b);
    const int n2 = smooth_run(pitch + p, F - p, 3, 30.0, s2, tmp, &a, &b);
    if (n1 < 3 || n2 < 3) { for (int i = 0; i < 5; ++i) out5[i] = NAN; return false; }
    out5[0] = ls_slope(s1, n1); out5[1] = ls_slope(s2, n2);
    out5[2] = ls_quad(s1, n1); out5[3] = ls_quad(s2, n2);
    const double m1 = median_inplace(s1, n1), m2 = median_inplace(s2, n2);
    out5[4] = m2 - m1;   // peakshift(seq1, seq2) = median(seq2) - median(seq1)
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// warp-cooperative pieces (device + emulator)
// ---------------------------------------------------------------------------------------------------------
DEVFN float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// Radix-4 Stockham FFT over shared memory, one warp per transform, natural order in and out.
// a: input (destroyed), b: scratch; returns the buffer holding the result.  tw = W1024^k.  No normalisation.
template <int N, bool INV>
DEVFN float2* warp_fft(float2* a, float2* b, const float2* tw, int lane) {
    constexpr int TS = kPitchFft / N;   // table stride
    int Ns = 1;
#pragma unroll 1
    for (; Ns * 4 <= N; Ns *= 4) {
        const int tstep = TS * (N / (Ns * 4));
        for (int j = lane; j < N / 4; j += 32) {
            const int k = j & (Ns - 1);
            float2 v0 = a[j], v1 = a[j + N / 4], v2 = a[j + N / 2], v3 = a[j + 3 * N / 4];
            if (Ns > 1) {
                float2 w1 = tw[k * tstep], w2 = tw[2 * k * tstep], w3 = tw[3 * k * tstep];
                if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
                v1 = cmulf(v1, w1); v2 = cmulf(v2, w2); v3 = cmulf(v3, w3);
            }
            const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y), t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
            float2 t3 = make_float2(v1.y - v3.y, v3.x - v1.x);   // -i (v1 - v3)
            if (INV) { t3.x = -t3.x; t3.y = -t3.y; }             // +i (v1 - v3)
            const int j0 = ((j - k) << 2) + k;
            b[j0] = make_float2(t0.x + t2.x, t0.y + t2.y);
            b[j0 + Ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
            b[j0 + 2 * Ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
            b[j0 + 3 * Ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
        }
        simt::warp_sync();
        float2* t = a; a = b; b = t;
    }
    if (Ns < N) {   // one radix-2 pass left (N = 512)
        const int tstep = TS * (N / (Ns * 2));
        for (int j = lane; j < N / 2; j += 32) {
            const int k = j & (Ns - 1);
            float2 v0 = a[j], v1 = a[j + N / 2];
            float2 w = tw[k * tstep];
            if (INV) w.y = -w.y;
            v1 = cmulf(v1, w);
            const int j0 = ((j - k) << 1) + k;
            b[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
            b[j0 + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
        }
        simt::warp_sync();
        float2* t = a; a = b; b = t;
    }
    return a;
}

DEVFN int warp_sum_i(int v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += simt::shfl32_i(v, (simt::tid() & 31) ^ m);
    return v;
}
DEVFN unsigned warp_min_u(unsigned v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { const unsigned o = (unsigned)simt::shfl32_i((int)v, (simt::tid() & 31) ^ m); v = o < v ? o : v; }
    return v;
}

// k-th smallest (0-based) of the warp's 16x32 unsigned keys, exact, by bitwise selection (31 value bits)
DEVFN unsigned warp_select(const unsigned (&key)[16], int k) {
    unsigned K = 0;
    for (int b = 30; b >= 0; --b) {
        const unsigned trial = K | ((1u << b) - 1u);
        int c = 0;
#pragma unroll
        for (int t = 0; t < 16; ++t) c += key[t] <= trial ? 1 : 0;
        c = warp_sum_i(c);
        if (c < k + 1) K |= 1u << b;
    }
    return K;
}

// median of the non-negative entries among the warp's 16x32 values (np.median(frame[frame >= 0]), pitch.py:146):
// NaN when there is none.  `valid` masks the entries that belong to the frame.
DEVFN float warp_median_nonneg(const float (&x)[16], int L, int lane) {
    unsigned key[16];
    int m_cnt = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const bool nn = (lane + 32 * t) < L && x[t] >= 0.f;
        key[t] = nn ? ((unsigned)__float_as_int_compat(x[t]) & 0x7fffffffu) : 0xffffffffu;   // -0.0 counts as 0
        m_cnt += nn ? 1 : 0;
    }
    m_cnt = warp_sum_i(m_cnt);
    if (m_cnt == 0) return NAN;
    const unsigned k1 = warp_select(key, (m_cnt - 1) >> 1);
    unsigned k2 = k1;
    if ((m_cnt & 1) == 0) {   // even count: average with the next order statistic
        int c = 0; unsigned nxt = 0xffffffffu;
#pragma unroll
        for (int t = 0; t < 16; ++t) { c += key[t] <= k1 ? 1 : 0; if (key[t] > k1 && key[t] < nxt) nxt = key[t]; }
        c = warp_sum_i(c); nxt = warp_min_u(nxt);
        k2 = (c >= (m_cnt >> 1) + 1) ? k1 : nxt;
    }
    // np.median of an even count is mean([a, b]) = (a + b) / 2
    return (__int_as_float_compat((int)k1) + __int_as_float_compat((int)k2)) * 0.5f;
}

// center_clip(frame, False) (pitch.py:145-155) on one value
DEVFN float clip_value(float v, float med) {
    if (v > med) return v - med;
    if (v < -med) return v + med;
    return 0.f;
}

// frame g -> utterance index: last u with frame_off[u] <= g
DEVFN int find_utt(const int64_t* frame_off, int n_utt, int64_t g) {
    int lo = 0, hi = n_utt;   // invariant: frame_off[lo] <= g < frame_off[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (frame_off[mid] <= g) lo = mid; else hi = mid; }
    return lo;
}

// One frame: gather + clip + FIR + (cepstrum | autocorrelation) -> p.rows[g, :], p.frame_amp[g].
// smem: two float2[1024] buffers owned by the warp.
DEVFN void pitch_frame_warp(const PitchParams& p, int64_t g, float2* bufa, float2* bufb) {
    const int lane = simt::tid() & 31;
    const int u = find_utt(p.frame_off, p.n_utt, g);
    const int64_t f = g - p.frame_off[u];
    const int64_t start = p.seg_start[u];
    const int64_t ubase = p.offsets[u];
    const int Ld = p.ds_len[u];
    const int L = p.frame_len;
    // ---- gather the frame (zero padded past the decimated length), 16 samples per lane: n = lane + 32 t
    float x[16];
    double asum = 0.0;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int n = lane + 32 * t;
        const int64_t k = f * p.frame_step + n;
        double v = 0.0;
        if (n < L && k < Ld) {
            const int64_t s = start + ds_index(k, p.ds_idx, p.ds_in, p.ds_out);   // packed-buffer sample index
            double cur, prev = 0.0;
            if (p.in_f32) { cur = reinterpret_cast<const float*>(p.pcm)[s]; if (s > ubase) prev = reinterpret_cast<const float*>(p.pcm)[s - 1]; }
            else { cur = reinterpret_cast<const int16_t*>(p.pcm)[s]; if (s > ubase) prev = reinterpret_cast<const int16_t*>(p.pcm)[s - 1]; }
            // preprocess.preemphasis (:11-19) in float64 like the reference: x[n] - c*x[n-1], x[0] kept
            v = (p.preemph != 0.0 && s > ubase) ? cur - p.preemph * prev : cur;
        }
        x[t] = (float)v;
        asum += fabs(v);
    }
    // sum |x| of the raw frame in float64 (pitch.py:65)
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) asum += shfl32_xor_f64(asum, m);
    if (lane == 0 && p.frame_amp) p.frame_amp[g] = asum;

    // ---- centre clip at the median of the non-negative samples (pitch.py:145-155); padding zeros are samples too
    float med = 0.f;
    if (p.do_clip) med = warp_median_nonneg(x, L, lane);
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int n = lane + 32 * t;
        const float c = p.do_clip ? clip_value(x[t], med) : x[t];
        bufa[n] = make_float2(n < L ? c : 0.f, 0.f);
        bufa[n + 512] = make_float2(0.f, 0.f);
    }
    simt::warp_sync();
    // ---- FIR band-pass: y = conv(x, h)[:L] through the 1024-point spectrum (sigproc.py:22-46)
    float2* A = warp_fft<1024, false>(bufa, bufb, p.tw, lane);
    float2* B = (A == bufa) ? bufb : bufa;
    for (int j = lane; j < 1024; j += 32) A[j] = cmulf(A[j], p.H[j]);
    simt::warp_sync();
    float2* Y = warp_fft<1024, true>(A, B, p.tw, lane);
    float2* Z = (Y == bufa) ? bufb : bufa;
    const float inv1024 = 1.0f / 1024.0f;
    if (p.mode == 0) {
        // ---- cepstrum: |ifft(log|fft(y)|)| (pitch.py:135-143); L == 512 here
        for (int j = lane; j < 512; j += 32) { float2 v = Y[j]; Y[j] = make_float2(v.x * inv1024, v.y * inv1024); }
        simt::warp_sync();
        float2* X = warp_fft<512, false>(Y, Z, p.tw, lane);
        float2* X2 = (X == bufa) ? bufb : bufa;
        for (int j = lane; j < 512; j += 32) { const float2 v = X[j]; X[j] = make_float2(dsp_logf(sqrtf(v.x * v.x + v.y * v.y)), 0.f); }
        simt::warp_sync();
        float2* C = warp_fft<512, true>(X, X2, p.tw, lane);
        const float inv512 = 1.0f / 512.0f;
        float* row = p.rows + g * p.row_len;
        for (int j = lane; j < p.row_len; j += 32) { const float2 v = C[j]; row[j] = sqrtf(v.x * v.x + v.y * v.y) * inv512; }
    } else {
        // ---- autocorrelation of |y| (pitch.py:112-132, sigproc.py:48-53) through |FFT|^2
        for (int j = lane; j < 1024; j += 32) {
            const float2 v = Y[j];
            Y[j] = make_float2(j < L ? sqrtf(v.x * v.x + v.y * v.y) * inv1024 : 0.f, 0.f);
        }
        simt::warp_sync();
        float2* V = warp_fft<1024, false>(Y, Z, p.tw, lane);
        float2* V2 = (V == bufa) ? bufb : bufa;
        for (int j = lane; j < 1024; j += 32) { const float2 v = V[j]; V[j] = make_float2(v.x * v.x + v.y * v.y, 0.f); }
        simt::warp_sync();
        float2* R = warp_fft<1024, true>(V, V2, p.tw, lane);
        float* row = p.rows + g * p.row_len;
        for (int j = lane; j < p.row_len; j += 32) {
            const int n = kMinLag + j;
            row[j] = (n < L) ? R[n].x * inv1024 / (float)(L - n) : NAN;
        }
    }
    simt::warp_sync();
}

// K4b / K5b: one CTA (512 threads) per utterance, frames in order, kTrackChunk at a time.
// smem: float chunk[kTrackChunk][row_len] + int sc[kTrackChunk][80].
DEVFN void pitch_track_cta(const PitchParams& p, float* chunk, int* sc) {
    const int u = simt::bid();
    const int tid = simt::tid();
    const int64_t f0 = p.frame_off[u];
    const int F = (int)(p.frame_off[u + 1] - f0);
    const int RL = p.row_len;
    const bool col = tid < RL;
    const float* base = p.rows + f0 * RL + tid;
    float s1 = 0.f, s2 = 0.f;                 // smoothed rows i-1 and i-2 of this thread's column
    float r0 = (col && F > 0) ? base[0] : 0.f;                 // raw rows i and i+1
    float r1 = (col && F > 1) ? base[RL] : 0.f;
    for (int c0 = 0; c0 < F; c0 += kTrackChunk) {
        const int nrows = F - c0 < kTrackChunk ? F - c0 : kTrackChunk;
        if (col) {
            for (int k = 0; k < nrows; ++k) {
                const int i = c0 + k;
                const float r2 = (i + 2 < F) ? base[(int64_t)(i + 2) * RL] : 0.f;   // prefetch
                // smooth (pitch.py:157-164): g[i] = mean(g[left:right]) in place => rows < i are already smoothed.
                // np.mean over axis 0 adds the rows in order, then divides by the count.
                const int right = (i + 2 < F) ? i + 2 : F - 1;
                const int left = i - 2 > 0 ? i - 2 : 0;
                float acc = 0.f; bool any = false;
                if (i - 2 >= 0 && i - 2 < right) { acc = s2; any = true; }
                if (i - 1 >= 0 && i - 1 < right) { acc = any ? acc + s1 : s1; any = true; }
                if (i < right) { acc = any ? acc + r0 : r0; any = true; }
                if (i + 1 < right) { acc = any ? acc + r1 : r1; any = true; }
                float gsm = any ? acc / (float)(right - left) : NAN;   // empty window: np.mean([]) = NaN
                if (p.no_smooth) gsm = r0;
                chunk[k * RL + tid] = gsm;
                if (p.rows_out) p.rows_out[(f0 + i) * RL + tid] = gsm;
                s2 = s1; s1 = gsm; r0 = r1; r1 = r2;
            }
        }
        simt::cta_sync();
        if (p.mode == 0) {
            // peak_score (pitch.py:227-242) for lags 20..99 of every row of the chunk
            for (int t = tid; t < nrows * kPeakLags; t += kTrackThreads) {
                const int k = t / kPeakLags, c = kMinLag + t % kPeakLags;
                const float* row = chunk + k * RL;
                const float v = row[c];
                int pp = c; while (pp > 0 && row[pp] <= v) --pp;
                // the right-hand scan matters only while it is shorter than the left-hand distance
                int qmax = c + (c - pp); if (qmax > RL) qmax = RL;
                int q = c; while (q < qmax && row[q] <= v) ++q;
                const int s = (c - pp) < (q - c) ? (c - pp) : (q - c);
                sc[t] = s;
                if (p.score) p.score[(f0 + c0 + k) * kPeakLags + (c - kMinLag)] = s;
            }
            simt::cta_sync();
            if (tid < nrows) {   // first argmax (pitch.py:169)
                const int* s = sc + tid * kPeakLags;
                int best = 0;
                for (int t = 1; t < kPeakLags; ++t) if (s[t] > s[best]) best = t;
                p.lag[f0 + c0 + tid] = kMinLag + best;
            }
        } else {
            // np.argmax over the smoothed scores, one warp per row: first maximum, a NaN counts as the maximum
            const int k = tid >> 5, lane = tid & 31;
            if (k < nrows) {
                const float* row = chunk + k * RL;
                float bv = 0.f; int bi = -1;
                for (int j = lane; j < RL; j += 32) {
                    const float v = row[j];
                    const bool better = bi < 0 || (v != v && bv == bv) || (bv == bv && v > bv);
                    if (better) { bv = v; bi = j; }
                }
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) {
                    const float ov = simt::shfl32_xor(bv, m);
                    const int oi = simt::shfl32_i(bi, lane ^ m);
                    const bool an = bv != bv, on = ov != ov;
                    bool take;
                    if (oi < 0) take = false;
                    else if (bi < 0) take = true;
                    else if (an || on) take = on && (!an || oi < bi);
                    else take = ov > bv || (ov == bv && oi < bi);
                    if (take) { bv = ov; bi = oi; }
                }
                if (lane == 0) p.lag[f0 + c0 + k] = kMinLag + (bi < 0 ? 0 : bi);
            }
        }
        simt::cta_sync();
    }
    if (tid == 0 && F > 0 && p.pitch) robust_pitch(p.lag + f0, F, p.pitch + f0);
}

// K6: pitch_feature tail, one thread per utterance
DEVFN void pitch_feature_thread(const PitchParams& p, int u) {
    const int64_t f0 = p.frame_off[u];
    const int F = (int)(p.frame_off[u + 1] - f0);
    pitch_feature_tail(p.pitch + f0, p.frame_amp + f0, F, p.scratch + 3 * f0, p.feat + 5 * (int64_t)u);
}

}  // namespace dspfe
