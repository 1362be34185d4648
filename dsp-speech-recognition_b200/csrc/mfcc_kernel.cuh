// K1: fused MFCC + delta + delta-delta over a ragged batch (sm_100a).
//
// Replaces, in one pass over packed int16 PCM, the reference chain
//   preemphasis (sigproc.py:178) -> framesig (sigproc.py:66) -> powspec (sigproc.py:151) ->
//   get_filterbanks/dot (base.py:28-29) -> log -> dct(ortho) (base.py:12-13) -> lifter (base.py:60)
//   -> energy substitution (base.py:15) -> delta, delta (base.py:70; model.py:76-77).
//
// Work decomposition
//   * one CTA (128 threads) per tile = up to seg_frames consecutive output frames of one utterance,
//     plus 2N halo frames either side (only when an utterance is split into several tiles);
//   * the CTA streams the tile's PCM in chunks of 16 frames: a 1-D bulk async copy (TMA engine,
//     cp.async.bulk + mbarrier) lands raw int16 in shared memory while the previous chunk computes;
//   * a vectorised conversion pass (8 samples per thread, int16->fp32 by mantissa splicing instead of
//     the slow I2F) turns the chunk into pre-emphasised fp32 once (frames overlap 2.5x);
//   * each 16-lane group owns a PAIR of frames carried in the two halves of packed-FP32 registers
//     (FADD2/FMUL2/FFMA2): 512-point real FFT as a 256-point complex FFT (radix-16 x radix-16, one
//     shared-memory transpose) + split post-pass done pairwise (k, 256-k) with 16-lane shuffles,
//     power spectrum -> shared, mel as range sums (each bin read once), log, DCT*lifter;
//   * MFCC rows stay in shared memory for the whole tile; the epilogue computes delta (edge clamp on
//     the utterance), delta-delta (edge clamp on the delta array, Appendix A-5) and writes [F,3*numcep]
//     rows with coalesced stores.
#pragma once
#include "dspfe_types.h"
#include "simt.h"
#include "fft_regs.h"

namespace dspfe {

struct MfccSmem {
    float* tables;
    simt::mbar_t* mbar;
    float2* scratch;   // per group: kScratchUnits float2
    float* mfcc;       // [(seg_frames + 4N)][numcep]
    float* fbuf;       // pre-emphasised fp32 samples of the current chunk
    unsigned char* raw;  // raw PCM chunk (bulk-copy destination): int16 or float32 samples
};

// Chunk geometry in 32-bit sample units relative to the utterance's aligned-down start (start & ~AL): the only
// 64-bit quantity is the base pointer.  Utterances are limited to 2^31 - 2^16 samples.
struct ChunkGeom {
    int a0;             // staging index 0 <-> relative sample a0 (16-byte aligned in the packed buffer)
    int bulk_bytes;     // multiple of 16, may be 0
    int tail_lo, tail_hi;  // relative sample range loaded with plain loads
    int s0;             // utterance sample index of the chunk's first frame
};

template <bool F32>
DEVFN ChunkGeom chunk_geom(const MfccParams& p, int sh, int lim, int S, int frame0) {
    const int AL = F32 ? 3 : 7;   // samples per 16 bytes, minus one
    ChunkGeom g;
    g.s0 = frame0 * p.frame_step;
    const int s_first = g.s0 > 0 ? g.s0 - 1 : 0;
    int s_last = g.s0 + p.fbuf_floats;
    if (s_last > S) s_last = S;
    if (s_last < s_first) s_last = s_first;
    const int q_first = sh + s_first, q_last = sh + s_last;
    g.a0 = q_first & ~AL;
    int a1 = (q_last + AL) & ~AL;
    if (a1 > lim) a1 = lim;
    if (a1 < g.a0) a1 = g.a0;
    g.bulk_bytes = (a1 - g.a0) * (F32 ? 4 : 2);
    g.tail_lo = a1 > q_first ? a1 : q_first;
    g.tail_hi = q_last;
    return g;
}

// NFULL = frame_len / 32 (FFT input rows that lie fully inside the frame) when known at compile time, else -1.
// F32IN: the packed batch holds float32 samples instead of int16 (e.g. a signal the caller already scaled).
// MODE: 0 = MFCC + delta + delta-delta rows [F, 3*numcep]; 1 = filterbank energies + frame energy [F, nfilt+1]
// (reference fbank, base.py:18); 2 = spectrum [F, 257]: power (sigproc.py:151), magnitude (:136) or 10*log10 power (:161).
// TRI: K1T -- nfft = 1536 (model.py:74), same tile / chunk / epilogue structure.
// The 1536-point real transform is a 768-point complex one on z[m] = x[2m] + i x[2m+1]; for frames of at most 512 samples only
// z[0..255] is non-zero, so the decimation in frequency by three needs no butterflies: Z[3q + r] = FFT256(z[m] W768^{m r})[q],
// K1's register transform.  The real split pairs Z[k] with Z[768 - k]: class 0 (k = 3q) with itself -- K1's split verbatim,
// W1536^{3q} = W512^q, both frames of the pair in the packed halves -- and class 2 (k = 3q + 2) with class 1
// (768 - k = 3 (255 - q) + 1).  For those the packed halves carry the two CLASSES of one frame instead of two frames: the
// transform of (z W768^m, z W768^2m) leaves Z_1 in the .x and Z_2 in the .y halves, the partner Z_1[255 - q] of Z_2[q] sits in
// the mirrored lane and register (as in the class-0 split), and nothing is parked in shared memory (the first version parked
// Z_1 of the pair: 4 KB per group, 2 CTAs/SM).  Three packed transforms per frame pair either way.
// Power bins stay in arrays indexed by q: class 0 as (frame A, frame B), classes 1 / 2 of one frame as (class 1, class 2);
// the mel pieces are built per class (mfcc_tables.h).
// LONG (K1T only): frames of 513 .. 1536 samples (30 ms at 44.1 / 48 kHz: 1323 / 1440) -- all three thirds of z are populated and
// the transform of class r starts from u_r[m] = (z[m] + w^r z[m+256] + w^2r z[m+512]) W768^{m r}, w = exp(-2 pi i / 3).
template <bool HAS_WIN, int NFULL, bool F32IN, int MODE, bool TRI = false, bool LONG = false>
DEVFN void mfcc_cta(const MfccParams& p, unsigned char* smem_raw) {
    const int tid = simt::tid();
    const int lane = tid & 15;
    const int grp = tid >> 4;
    const int tile_id = simt::bid();
    if (tile_id >= *p.ntiles) return;

    MfccSmem sm;
    sm.tables = reinterpret_cast<float*>(smem_raw);
    sm.mbar = reinterpret_cast<simt::mbar_t*>(smem_raw + p.sm_mbar);
    sm.scratch = reinterpret_cast<float2*>(smem_raw + p.sm_scratch);
    sm.mfcc = reinterpret_cast<float*>(smem_raw + p.sm_mfcc);
    sm.fbuf = reinterpret_cast<float*>(smem_raw + p.sm_fbuf);
    sm.raw = smem_raw + p.sm_raw;

    const Tile tile = p.tiles[tile_id];
    const int u = tile.utt;
    const int64_t start = p.seg_start[u];
    const int S = p.seg_len[u];
    const int64_t row0 = p.frame_off[u];
    const int F = (int)(p.frame_off[u + 1] - row0);
    const int N = MODE == 0 ? p.delta_n : 0;   // the tap modes need no delta halo
    const int numcep = p.numcep;
    const int v_lo = tile.f0 - 2 * N > 0 ? tile.f0 - 2 * N : 0;
    const int v_hi = tile.f0 + tile.nf - 1 + 2 * N < F - 1 ? tile.f0 + tile.nf - 1 + 2 * N : F - 1;
    const int nchunks = (v_hi - v_lo + kFramesPerPass) / kFramesPerPass;

    // ---- constant tables -> shared, mbarrier init, first chunk in flight
    for (int i = tid; i < p.tbl_floats; i += kMfccThreads) sm.tables[i] = p.tables[i];
    if (tid == 0) { simt::mbar_init(sm.mbar, 1); simt::fence_mbar_init(); }
    simt::cta_sync();

    const float2* twa = reinterpret_cast<const float2*>(sm.tables + p.o_twa);
    const float2* twp = reinterpret_cast<const float2*>(sm.tables + p.o_twp);
    const int4* melc = reinterpret_cast<const int4*>(sm.tables + p.o_melc);  // [filter] -> the partial sums that make it up
    const float* dct = sm.tables + p.o_dct;
    const float* win = sm.tables + p.o_win;
    float2* scr = sm.scratch + grp * kScratchUnits;

    // everything below addresses the packed buffer relative to the utterance's aligned-down start
    const int esz = F32IN ? 4 : 2;
    const int sh = (int)(start & (F32IN ? 3 : 7));
    const unsigned char* src8 = reinterpret_cast<const unsigned char*>(p.pcm) + (start - sh) * esz;
    int lim;   // relative end of the last whole 16-byte unit of the packed buffer
    {
        const int64_t l64 = (p.total_samples & ~(int64_t)(F32IN ? 3 : 7)) - (start - sh);
        lim = l64 > 0x7fff0000 ? 0x7fff0000 : (int)l64;
    }
    auto issue_chunk = [&](const ChunkGeom& g) {
        if (tid == 0 && g.bulk_bytes > 0) {
            simt::fence_proxy_async();
            simt::mbar_expect_tx(sm.mbar, (uint32_t)g.bulk_bytes);
            simt::bulk_g2s(sm.raw, src8 + (int64_t)g.a0 * esz, (uint32_t)g.bulk_bytes, sm.mbar);
        }
        for (int i = g.tail_lo + tid; i < g.tail_hi; i += kMfccThreads) {
            if (F32IN) reinterpret_cast<float*>(sm.raw)[i - g.a0] = reinterpret_cast<const float*>(src8)[i];
            else reinterpret_cast<int16_t*>(sm.raw)[i - g.a0] = reinterpret_cast<const int16_t*>(src8)[i];
        }
    };
    ChunkGeom g = chunk_geom<F32IN>(p, sh, lim, S, v_lo);
    issue_chunk(g);

    const int nfull = NFULL >= 0 ? NFULL : (p.frame_len >> 5);
    uint32_t parity = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int frame0 = v_lo + c * kFramesPerPass;
        if (g.bulk_bytes > 0) { simt::mbar_wait(sm.mbar, parity); parity ^= 1; }
        if (c == 0) simt::cta_sync();  // chunk-0 tail stores -> visible (later chunks: the loop-end barrier)

        // ---- raw int16 -> pre-emphasised fp32, 8 samples per thread.  Sample q of the staging area (packed sample
        // a0s + q) lands in plane q&1 at index q>>1.  Reference: y[0] = x[0], y[n] = x[n] - c*x[n-1] (sigproc.py:185), zeros past the end (:84-87).
        {
            const int rel = g.a0 - sh;  // utterance sample index of raw[0] (may be negative)
            const float cpre = p.preemph;
            for (int j0 = 0; j0 < p.fbuf_vecs; j0 += kMfccThreads) {
                const int j = j0 + tid;
                const bool act = j < p.fbuf_vecs;
                float e[9];
                if (F32IN) {
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (act) { v0 = reinterpret_cast<const float4*>(sm.raw)[2 * j]; v1 = reinterpret_cast<const float4*>(sm.raw)[2 * j + 1]; }
                    float pv = __int_as_float_compat(simt::shfl32_i(__float_as_int_compat(v1.w), (tid & 31) - 1));
                    if ((tid & 31) == 0) pv = (act && j > 0) ? reinterpret_cast<const float*>(sm.raw)[8 * j - 1] : 0.f;
                    e[0] = pv; e[1] = v0.x; e[2] = v0.y; e[3] = v0.z; e[4] = v0.w; e[5] = v1.x; e[6] = v1.y; e[7] = v1.z; e[8] = v1.w;
                } else {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (act) v = reinterpret_cast<const uint4*>(sm.raw)[j];
                    uint32_t pw = (uint32_t)simt::shfl32_i((int)v.w, (tid & 31) - 1);   // previous thread's last pair
                    if ((tid & 31) == 0) pw = (act && j > 0) ? ((uint32_t)reinterpret_cast<const uint16_t*>(sm.raw)[8 * j - 1]) << 16 : 0u;
                    e[0] = cvt_hi16(pw);
                    e[1] = cvt_lo16(v.x); e[2] = cvt_hi16(v.x); e[3] = cvt_lo16(v.y); e[4] = cvt_hi16(v.y);
                    e[5] = cvt_lo16(v.z); e[6] = cvt_hi16(v.z); e[7] = cvt_lo16(v.w); e[8] = cvt_hi16(v.w);
                }
                if (act) {
                    float y[8];
#pragma unroll
                    for (int m = 0; m < 8; ++m) y[m] = dsp_fmaf(-cpre, e[m], e[m + 1]);
                    const int s_first = rel + 8 * j;
                    if (s_first < 1 || s_first + 7 >= S) {   // utterance head / tail: patch element-wise
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
                            const int s = s_first + m;
                            if (s < 0 || s >= S) y[m] = 0.f;
                            else if (s == 0) y[m] = e[m + 1];
                        }
                    }
                    // even-index samples and odd-index samples go to separate planes: the FFT loads (re from one
                    // plane, im from the other, consecutive lanes -> consecutive words) are then conflict-free
                    reinterpret_cast<float4*>(sm.fbuf)[j] = make_float4(y[0], y[2], y[4], y[6]);
                    reinterpret_cast<float4*>(sm.fbuf + 4 * p.fbuf_vecs)[j] = make_float4(y[1], y[3], y[5], y[7]);
                }
            }
        }
        simt::cta_sync();
        const int fb_base = sh - g.a0 + g.s0;     // staging index of sample 0 of the chunk's first frame
        if (c + 1 < nchunks) {   // raw is free again: overlap the next load with the FFTs
            g = chunk_geom<F32IN>(p, sh, lim, S, frame0 + kFramesPerPass);
            issue_chunk(g);
        }

        // ---- one frame pair per 16-lane group; a warp (two groups) is active or idle as a whole.  Within a warp
        // group 0 packs frames (f, f+2) and group 1 packs (f+1, f+3): the two groups' loads then fall into
        // disjoint halves of the 32 banks (frame step/2 = 80 words = 16 mod 32 for the 10 ms step).
        const int fl = 4 * (tid >> 5) + (grp & 1);   // chunk-local index of the frame in the .x halves; .y = fl + 2
        const int vA = frame0 + fl;
        if (frame0 + 4 * (tid >> 5) <= v_hi) {
            cpx2 x[16];
            // windowed samples of the pair as z[m] = x[2m] + i x[2m+1], m = 16 n1 + lane
            auto load_pair = [&](cpx2 (&x)[16]) {
                const int fb0 = fb_base + fl * p.frame_step;   // staging index of the frame's sample 0
                const float* plane_e = sm.fbuf;
                const float* plane_o = sm.fbuf + 4 * p.fbuf_vecs;
                const bool odd = (fb0 & 1) != 0;                                   // uniform over the CTA (even frame step)
                const float* fre = (odd ? plane_o : plane_e) + (fb0 >> 1) + lane;      // sample 2n   of the frame
                const float* fim = (odd ? plane_e + 1 : plane_o) + (fb0 >> 1) + lane;  // sample 2n+1 of the frame
                const int dB = p.frame_step;                                       // two frames ahead = step words in a plane
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    // four scalar loads land directly in the (frame A, frame B) register pairs
                    float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;
                    if (n1 < nfull) {
                        ar = fre[16 * n1]; ai = fim[16 * n1]; br = fre[16 * n1 + dB]; bi = fim[16 * n1 + dB];
                    } else if (n1 == nfull) {
                        const int i0 = 32 * n1 + 2 * lane;
                        if (i0 < p.frame_len) { ar = fre[16 * n1]; br = fre[16 * n1 + dB]; }
                        if (i0 + 1 < p.frame_len) { ai = fim[16 * n1]; bi = fim[16 * n1 + dB]; }
                    }
                    if (HAS_WIN) {   // window planes are zero-padded to 256 entries each
                        const float w0 = win[16 * n1 + lane], w1 = win[256 + 16 * n1 + lane];
                        ar *= w0; br *= w0; ai *= w1; bi *= w1;
                    }
                    x[n1].re = make_float2(ar, br);
                    x[n1].im = make_float2(ai, bi);
                }
            };
            // K1T LONG, class 0: z[m] + z[m+256] + z[m+512] of the pair (window planes of 768 entries)
            auto load_pair_long = [&](cpx2 (&x)[16]) {
                const int fb0 = fb_base + fl * p.frame_step;
                const float* plane_e = sm.fbuf;
                const float* plane_o = sm.fbuf + 4 * p.fbuf_vecs;
                const bool odd = (fb0 & 1) != 0;
                const float* fre = (odd ? plane_o : plane_e) + (fb0 >> 1) + lane;
                const float* fim = (odd ? plane_e + 1 : plane_o) + (fb0 >> 1) + lane;
                const int dB = p.frame_step;
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    cpx2 acc; acc.re = make_float2(0.f, 0.f); acc.im = acc.re;
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int m0 = 256 * t + 16 * n1;                  // z index of lane 0
                        float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;
                        if (2 * m0 + 31 < p.frame_len) {
                            ar = fre[m0]; ai = fim[m0]; br = fre[m0 + dB]; bi = fim[m0 + dB];
                        } else {
                            const int i0 = 2 * (m0 + lane);
                            if (i0 < p.frame_len) { ar = fre[m0]; br = fre[m0 + dB]; }
                            if (i0 + 1 < p.frame_len) { ai = fim[m0]; bi = fim[m0 + dB]; }
                        }
                        if (HAS_WIN) {
                            const float w0 = win[m0 + lane], w1 = win[768 + m0 + lane];
                            ar *= w0; br *= w0; ai *= w1; bi *= w1;
                        }
                        acc.re = f2add(acc.re, make_float2(ar, br)); acc.im = f2add(acc.im, make_float2(ai, bi));
                    }
                    x[n1] = acc;
                }
            };
            // K1T, classes 1 and 2 of ONE frame (f = 0: frame A, 1: frame B) in the packed halves: x[n1] = (u_1[m], u_2[m]),
            // u_r[m] = (z[m] + w^r z[m+256] + w^2r z[m+512]) W768^{m r}, w = exp(-2 pi i / 3); a short frame has z[m] only
            auto load_classes = [&](cpx2 (&x)[16], int f) {
                const int fb0 = fb_base + (fl + 2 * f) * p.frame_step;
                const float* plane_e = sm.fbuf;
                const float* plane_o = sm.fbuf + 4 * p.fbuf_vecs;
                const bool odd = (fb0 & 1) != 0;
                const float* fre = (odd ? plane_o : plane_e) + (fb0 >> 1) + lane;
                const float* fim = (odd ? plane_e + 1 : plane_o) + (fb0 >> 1) + lane;
                const float4* tw12 = reinterpret_cast<const float4*>(sm.tables + p.o_tw3) + lane;
                constexpr int WP = LONG ? 768 : 256;
                const float hs = 0.8660254037844386f;
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    float zr[3] = {0.f, 0.f, 0.f}, zi[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int t = 0; t < (LONG ? 3 : 1); ++t) {
                        const int m0 = 256 * t + 16 * n1;
                        float ar = 0.f, ai = 0.f;
                        if (LONG ? (2 * m0 + 31 < p.frame_len) : (n1 < nfull)) {
                            ar = fre[m0]; ai = fim[m0];
                        } else if (LONG || n1 == nfull) {
                            const int i0 = 2 * (m0 + lane);
                            if (i0 < p.frame_len) ar = fre[m0];
                            if (i0 + 1 < p.frame_len) ai = fim[m0];
                        }
                        if (HAS_WIN) { ar *= win[m0 + lane]; ai *= win[WP + m0 + lane]; }
                        zr[t] = ar; zi[t] = ai;
                    }
                    float2 ure, uim;     // (class 1, class 2) before the twiddle
                    if (LONG) {
                        // z0 + w z1 + w^2 z2 = z0 - s/2 - i (sqrt 3 / 2) d and z0 + w^2 z1 + w z2 = z0 - s/2 + i (sqrt 3 / 2) d, s = z1 + z2, d = z1 - z2
                        const float br = zr[0] - 0.5f * (zr[1] + zr[2]), bi = zi[0] - 0.5f * (zi[1] + zi[2]);
                        const float er = hs * (zi[1] - zi[2]), ei = -hs * (zr[1] - zr[2]);       // -i (sqrt 3 / 2) d
                        ure = make_float2(br + er, br - er); uim = make_float2(bi + ei, bi - ei);
                    } else {
                        ure = make_float2(zr[0], zr[0]); uim = make_float2(zi[0], zi[0]);
                    }
                    const float4 tw = tw12[16 * n1];
                    const float2 C = make_float2(tw.x, tw.y), S = make_float2(tw.z, tw.w);
                    x[n1].re = f2fma(uim, f2neg(S), f2mul(ure, C));
                    x[n1].im = f2fma(uim, C, f2mul(ure, S));
                }
            };
            // 256-point complex transform of the pair: lane = n2 in, lane = k1 out, register k2 holds Z[k1 + 16*k2]
            auto fft256 = [&](cpx2 (&x)[16]) {
                // stage 1: DFT over n1 (registers); lane = n2
                dft16(x);
                // twiddle W256^{n2*k1}
#pragma unroll
                for (int k1 = 1; k1 < 16; ++k1) { const float2 w = twa[k1 * 16 + lane]; x[k1] = cmuls(x[k1], w.x, w.y); }
                // transpose through shared: (k1, n2) -> lane k1, register n2; real parts then imaginary parts
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1) scr[k1 * 17 + lane] = x[k1].re;
                simt::group_sync();
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) x[n2].re = scr[lane * 17 + n2];
                simt::group_sync();
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1) scr[k1 * 17 + lane] = x[k1].im;
                simt::group_sync();
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) x[n2].im = scr[lane * 17 + n2];
                simt::group_sync();
                // stage 2: DFT over n2; lane = k1, register k2 holds Z[k1 + 16*k2]
                dft16(x);
            };
            if constexpr (LONG) load_pair_long(x); else load_pair(x);
            fft256(x);

            // ---- real-FFT split, pairwise: bins k = lane + 16 r (r < 8) and 256 - k share one butterfly.
            // partner Z[(256-k) mod 256] lives in lane (16-lane)&15, register 15-r (lane 0: register (16-r)&15).
            // (K1T: the same butterflies give the class-0 bins X[3k] and X[3 (256 - k)] of the 1536-point transform.)
            const int src = (16 - lane) & 15;
            const float sc = p.pow_scale;  // 1 / (4 * NFFT)
            float2 esum = make_float2(0.f, 0.f);   // this lane's share of sum_k P[k] (frame energy, reference base.py:25)
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                if (r == 8 && lane != 0) break;  // lane 0 also owns the self-paired bin 128
                cpx2 a = x[r], b;
                if (r < 8) {
                    b.re.x = simt::shfl16(x[15 - r].re.x, src); b.re.y = simt::shfl16(x[15 - r].re.y, src);
                    b.im.x = simt::shfl16(x[15 - r].im.x, src); b.im.y = simt::shfl16(x[15 - r].im.y, src);
                    if (lane == 0) b = x[(16 - r) & 15];
                } else {
                    b = x[8];
                }
                // s = a + conj(b), d = a - conj(b); X[k] = (s + W^k * (-i d)) / 2, X[256-k] = conj(s - W^k * (-i d)) / 2
                const float2 sre = f2add(a.re, b.re), sim = f2sub(a.im, b.im);
                const float2 dre = f2sub(a.re, b.re), dim = f2add(a.im, b.im);
                const float2 w = twp[r * 16 + lane];  // W512^k
                // t = W^k * (d.im, -d.re)
                const float2 tre = f2fmas(dim, w.x, f2muls(dre, w.y));
                const float2 tim = f2fmas(dim, w.y, f2muls(dre, -w.x));
                // square X[k] = s + t and X[256-k] = conj(s - t) separately: forming |s|^2+|t|^2 +- 2Re(s conj t)
                // would cancel catastrophically for the weaker bin of the pair
                const float2 are = f2add(sre, tre), aim = f2add(sim, tim);
                const float2 bre = f2sub(sre, tre), bim = f2sub(sim, tim);
                const float2 pa = f2muls(f2fma(aim, aim, f2mul(are, are)), sc);
                const float2 pb = f2muls(f2fma(bim, bim, f2mul(bre, bre)), sc);
                const int k = lane + 16 * r;
                // the stage-2 reads of scr are complete (group_sync above); P aliases the transpose tile
                if (r < 8) {
                    scr[k] = pa;                     // lane 0, r 0: Z[0] pairs with itself -> X[0] and X[256]
                    scr[256 - k] = pb;
                    esum = f2add(esum, f2add(pa, pb));
                } else {
                    scr[128] = pa;
                    esum = f2add(esum, pa);
                }
                if (r == 0 && lane != 0) scr[256 + lane] = make_float2(0.f, 0.f);   // the mel pieces may read (with zero weights) past bin 256
            }
            simt::group_sync();
            if (MODE == 2) {   // spectrum tap: rows straight to global memory
                const float nfft = (float)kNfft;
                for (int k = lane; k < kBins; k += 16) {
                    float2 v = scr[k];
                    if (p.spec_kind == 1) { v.x = sqrtf(v.x * nfft); v.y = sqrtf(v.y * nfft); }
                    else if (p.spec_kind == 2) { v.x = 10.f * log10f(fmaxf(v.x, 1e-30f)); v.y = 10.f * log10f(fmaxf(v.y, 1e-30f)); }
                    if (vA <= v_hi) p.out[(row0 + vA) * kBins + k] = v.x;
                    if (vA + 2 <= v_hi) p.out[(row0 + vA + 2) * kBins + k] = v.y;
                }
            } else {

            // ---- mel filterbank (reference base.py:28-29): every lane accumulates one piece of a filter per slot; the
            // trip counts are uniform over the lanes and the weights come from a [iteration][lane] table (mfcc_tables.h)
            constexpr int NCLS = TRI ? 2 : 1;
            float2 macc[NCLS][kMelSlots];
            auto mel_round = [&](int cls, const float2* P, float2 (&acc)[kMelSlots]) {
                const float2* melw = reinterpret_cast<const float2*>(sm.tables + p.o_melw_c[cls]) + lane;   // [iteration pair][lane] filter weights
                const int* melb = reinterpret_cast<const int*>(sm.tables + p.o_melb_c[cls]) + lane;         // [slot][lane] first bin of the lane's piece
                int tb = 0;
#pragma unroll
                for (int s = 0; s < kMelSlots; ++s) {
                    const int T = p.mel_Tc[cls][s];
                    const float2* pp = P + melb[s * kGroupLanes];
                    const float2* wp = melw + (tb >> 1) * kGroupLanes;
                    float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll 4
                    for (int t = 0; t < T; t += 2) {   // T is even
                        const float2 w = wp[(t >> 1) * kGroupLanes];
                        a0 = f2fmas(pp[t], w.x, a0);
                        a1 = f2fmas(pp[t + 1], w.y, a1);
                    }
                    acc[s] = f2add(a0, a1);
                    tb += T;
                }
            };
            mel_round(0, scr, macc[0]);
            if constexpr (TRI) {
                const float2* tws2 = reinterpret_cast<const float2*>(sm.tables + p.o_tws2) + lane;
                const float2* melw12 = reinterpret_cast<const float2*>(sm.tables + p.o_melw_c[1]) + lane;   // [iteration][lane] (class-1, class-2) weights
                const int* melb12 = reinterpret_cast<const int*>(sm.tables + p.o_melb_c[1]) + lane;
                float* qf = reinterpret_cast<float*>(scr);       // Q[q] = (class-1 bin 3q + 1, class-2 bin 3q + 2) of the frame
                const int mir = 15 - lane;
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    simt::group_sync();          // the bins in the tile have been consumed: it is free again
                    load_classes(x, f);
                    fft256(x);                   // .x halves: Z_1[q], .y halves: Z_2[q], q = lane + 16 k2
                    float es = 0.f;
                    const bool swp = (grp & 1) != 0;
                    const int st1 = swp ? -32 : 32;                                   // words per k2 step of the first store
                    float* d1 = qf + (swp ? 2 * (255 - lane) : 2 * lane + 1);
                    float* d2 = qf + (swp ? 2 * lane + 1 : 2 * (255 - lane));
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        // Z[3q + 2] = Z_2[q] pairs with Z[768 - (3q + 2)] = Z_1[255 - q]: lane 15 - lane, register 15 - k2
                        const int q = lane + 16 * k2;
                        const float are_ = x[k2].re.y, aim_ = x[k2].im.y;
                        const float bre_ = simt::shfl16(x[15 - k2].re.x, mir), bim_ = simt::shfl16(x[15 - k2].im.x, mir);
                        const float sre = are_ + bre_, sim = aim_ - bim_, dre = are_ - bre_, dim = aim_ + bim_;
                        const float2 w = tws2[16 * k2];  // W1536^{3q + 2}
                        const float tre = dsp_fmaf(dim, w.x, dre * w.y), tim = dsp_fmaf(dim, w.y, -dre * w.x);
                        const float2 xr = make_float2(sre + tre, sre - tre), xi = make_float2(sim + tim, sim - tim);
                        const float2 pw = f2muls(f2fma(xi, xi, f2mul(xr, xr)), sc);   // (bin 3q + 2, bin 3 (255 - q) + 1)
                        // (the two groups of a warp issue these stores together and their tiles start on the same bank: the odd group
                        // stores the pair in the opposite order -- d1 / d2 below -- so one instruction writes odd words in one group and
                        // even words in the other)
                        d1[st1 * k2] = swp ? pw.y : pw.x;
                        d2[-st1 * k2] = swp ? pw.x : pw.y;
                        es += pw.x + pw.y;
                    }
                    if (f == 0) esum.x += es; else esum.y += es;
                    scr[256 + lane] = make_float2(0.f, 0.f);     // 256 bins per class; the pieces may read (with zero weights) past them
                    simt::group_sync();
                    int tb = 0;
                    float2 tot[kMelSlots];
#pragma unroll
                    for (int s2 = 0; s2 < kMelSlots; ++s2) {
                        const int T = p.mel_Tc[1][s2];
                        const float2* pp = scr + melb12[s2 * kGroupLanes];
                        const float2* wp = melw12 + tb * kGroupLanes;
                        float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll 4
                        for (int t = 0; t < T; t += 2) {   // T is even
                            a0 = f2fma(pp[t], wp[t * kGroupLanes], a0);
                            a1 = f2fma(pp[t + 1], wp[(t + 1) * kGroupLanes], a1);
                        }
                        tot[s2] = f2add(a0, a1);
                        tb += T;
                    }
#pragma unroll
                    for (int s2 = 0; s2 < kMelSlots; ++s2) {
                        const float v = tot[s2].x + tot[s2].y;       // the lane's piece over both classes
                        if (f == 0) macc[1][s2].x = v; else macc[1][s2].y = v;
                    }
                }
            }
            // total frame energy (reference base.py:25): reduce over the group
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) {
                esum.x += simt::shfl16(esum.x, lane ^ m);
                esum.y += simt::shfl16(esum.y, lane ^ m);
            }
            simt::group_sync();  // all lanes are done reading P before the staging area overwrites it
            float2* part = scr;                                   // [NCLS * kMelSlots * 16 + 1]; the last entry stays zero
            float2* lmel = scr + NCLS * kMelSlots * kGroupLanes + 2;     // [nfilt + 1]; the last entry is log(energy); 16-byte aligned
#pragma unroll
            for (int c = 0; c < NCLS; ++c)
#pragma unroll
                for (int s = 0; s < kMelSlots; ++s) part[(c * kMelSlots + s) * kGroupLanes + lane] = macc[c][s];
            if (lane == 0) part[NCLS * kMelSlots * kGroupLanes] = make_float2(0.f, 0.f);
            simt::group_sync();
            const float eps64 = 2.220446049250313e-16f;  // numpy.finfo(float64).eps, reference base.py:26,30
            for (int j = lane; j <= p.nfilt; j += 16) {
                float2 f = esum;
                if (j < p.nfilt) {
                    const int4 ci = melc[j * NCLS];
                    f = f2add(f2add(part[ci.x], part[ci.y]), f2add(part[ci.z], part[ci.w]));
#pragma unroll
                    for (int c = 1; c < NCLS; ++c) {
                        const int4 cj = melc[j * NCLS + c];
                        f = f2add(f, f2add(f2add(part[cj.x], part[cj.y]), f2add(part[cj.z], part[cj.w])));
                    }
                }
                if (f.x == 0.f) f.x = eps64;
                if (f.y == 0.f) f.y = eps64;
                if (MODE == 1) {   // filterbank tap (column nfilt = frame energy)
                    if (vA <= v_hi) p.out[(row0 + vA) * (p.nfilt + 1) + j] = f.x;
                    if (vA + 2 <= v_hi) p.out[(row0 + vA + 2) * (p.nfilt + 1) + j] = f.y;
                } else {
                    lmel[j] = make_float2(dsp_fast_logf(f.x), dsp_fast_logf(f.y));
                }
            }
            simt::group_sync();
            // ---- DCT-II (ortho) * lifter, c0 := log(energy)
            if (MODE == 0 && lane < numcep) {
                float2 acc = make_float2(0.f, 0.f);
                const float2* drow = reinterpret_cast<const float2*>(dct + lane * p.dct_stride);
                const float4* lm4 = reinterpret_cast<const float4*>(lmel);
                const int half = (p.nfilt + 1) >> 1;   // odd nfilt: the pad coefficient is 0 and multiplies log(energy)
#pragma unroll 4
                for (int m = 0; m < half; ++m) {
                    const float4 l = lm4[m];
                    const float2 w = drow[m];
                    acc = f2fmas(make_float2(l.x, l.y), w.x, acc);
                    acc = f2fmas(make_float2(l.z, l.w), w.y, acc);
                }
                if (lane == 0 && p.append_energy) acc = lmel[p.nfilt];
                if (vA <= v_hi) sm.mfcc[(vA - v_lo) * numcep + lane] = acc.x;
                if (vA + 2 <= v_hi) sm.mfcc[(vA + 2 - v_lo) * numcep + lane] = acc.y;
            }
            }   // MODE != 2
        }
        simt::cta_sync();  // fbuf may be overwritten by the next conversion pass
    }

    if (MODE != 0) return;
    // ---- epilogue: delta (clamped on the utterance, reference base.py:70-79), delta-delta (= delta of the already
    // clamped delta array, model.py:76-77 / Appendix A-5) and the [F, 3*numcep] row stores.  Thread = (row r of a
    // block of 8 rows, column c): a thread keeps its column, rows advance by 8, so there is no index arithmetic in
    // the loops; rows away from the utterance ends take the clamp-free path.
    float* dbuf = sm.fbuf;                                   // aliases fbuf+raw: no copy is in flight any more
    const int u_lo = tile.f0 - N > 0 ? tile.f0 - N : 0;
    const int u_hi = tile.f0 + tile.nf - 1 + N < F - 1 ? tile.f0 + tile.nf - 1 + N : F - 1;
    const int width = 3 * numcep;
    const float dscale = p.delta_scale;
    const int nc2 = 2 * numcep;
    auto taps = [&](const float* q, int row) {   // sum_n n * (q[row + n] - q[row - n]), rows clamped to [0, F-1]
        float acc = 0.f;
        if (row >= N && row + N <= F - 1) {
            if (N == 2) {   // the reference call sites' N (model.py:76-77 uses 3; python_speech_features' default is 2)
                acc = dsp_fmaf(2.f, q[nc2] - q[-nc2], q[numcep] - q[-numcep]);
            } else {
                for (int n = 1; n <= N; ++n) acc = dsp_fmaf((float)n, q[n * numcep] - q[-n * numcep], acc);
            }
        } else {
            for (int n = 1; n <= N; ++n) {
                const int hi = row + n > F - 1 ? F - 1 - row : n, lo = row - n < 0 ? row : n;
                acc = dsp_fmaf((float)n, q[hi * numcep] - q[-lo * numcep], acc);
            }
        }
        return acc * dscale;
    };
    // (Round 2 experiment, reverted: the delta-delta rows to shared memory as well and the tile's contiguous [nf, 3*numcep] block
    // written with 16-byte vector stores -- every vector element located by a divide-by-width -- measured 4 % slower on the
    // whole kernel (0.603 against 0.582 ms per 4096 x 2 s): the kernel is bound by the shared-memory pipe, not by its stores,
    // and the extra pass adds wavefronts.  The 13-lane row stores below fill whole 32-byte sectors in L2 all the same.)
    const int rstep = kMfccGroups * numcep, ostep = kMfccGroups * width;
    if (lane < numcep) {
        int uu = u_lo + grp;
        const float* m0 = sm.mfcc + (uu - v_lo) * numcep + lane;
        float* db = dbuf + grp * numcep + lane;
        float* o = p.out + (row0 + uu) * width + lane;
        const int t_lo = tile.f0, t_hi = tile.f0 + tile.nf;
        for (; uu <= u_hi; uu += kMfccGroups, m0 += rstep, db += rstep, o += ostep) {
            const float d = taps(m0, uu);
            *db = d;
            if (uu >= t_lo && uu < t_hi) { o[0] = *m0; o[numcep] = d; }
        }
    }
    simt::cta_sync();
    if (lane < numcep) {
        int tt = tile.f0 + grp;
        const float* d0 = dbuf + (tt - u_lo) * numcep + lane;
        float* o = p.out + (row0 + tt) * width + nc2 + lane;
        for (; tt < tile.f0 + tile.nf; tt += kMfccGroups, d0 += rstep, o += ostep) *o = taps(d0, tt);
    }
}

}  // namespace dspfe
