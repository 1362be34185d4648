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

struct ChunkGeom {
    int64_t a0s;        // packed-buffer sample index that lands at raw[0] (16-byte aligned)
    int64_t bulk_src;   // == a0s
    int bulk_bytes;     // multiple of 16, may be 0
    int64_t tail_lo, tail_hi;  // packed sample range loaded with plain loads
    int64_t s0;         // utterance sample index of fbuf[0]
};

template <bool F32>
DEVFN ChunkGeom chunk_geom(const MfccParams& p, int64_t start, int S, int frame0) {
    const int64_t AL = F32 ? 3 : 7;   // samples per 16 bytes, minus one
    ChunkGeom g;
    g.s0 = (int64_t)frame0 * p.frame_step;
    int64_t s_first = g.s0 > 0 ? g.s0 - 1 : 0;
    int64_t s_last = g.s0 + p.fbuf_floats;
    if (s_last > S) s_last = S;
    if (s_last < s_first) s_last = s_first;
    int64_t p_first = start + s_first, p_last = start + s_last;
    g.a0s = p_first & ~AL;
    int64_t a1s = (p_last + AL) & ~AL;
    int64_t lim = p.total_samples & ~AL;
    if (a1s > lim) a1s = lim;
    if (a1s < g.a0s) a1s = g.a0s;
    g.bulk_src = g.a0s;
    g.bulk_bytes = (int)(a1s - g.a0s) * (F32 ? 4 : 2);
    g.tail_lo = a1s > p_first ? a1s : p_first;
    g.tail_hi = p_last;
    return g;
}

// NFULL = frame_len / 32 (FFT input rows that lie fully inside the frame) when known at compile time, else -1.
// F32IN: the packed batch holds float32 samples instead of int16 (e.g. a signal the caller already scaled).
// MODE: 0 = MFCC + delta + delta-delta rows [F, 3*numcep]; 1 = filterbank energies + frame energy [F, nfilt+1]
// (reference fbank, base.py:18); 2 = spectrum [F, 257]: power (sigproc.py:151), magnitude (:136) or 10*log10 power (:161).
template <bool HAS_WIN, int NFULL, bool F32IN, int MODE>
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
    const int* subi = reinterpret_cast<const int*>(sm.tables + p.o_sub);   // per sub-range: lo, len | fi0, gi0, inv
    const float* subf = sm.tables + p.o_sub;
    const int* rsub = reinterpret_cast<const int*>(sm.tables + p.o_rsub);  // per range: first sub-range, count
    const int* task = reinterpret_cast<const int*>(sm.tables + p.o_task);
    const float* dct = sm.tables + p.o_dct;
    const float* win = sm.tables + p.o_win;
    float2* scr = sm.scratch + grp * kScratchUnits;

    auto issue_chunk = [&](int c) {
        ChunkGeom g = chunk_geom<F32IN>(p, start, S, v_lo + c * kFramesPerPass);
        const int esz = F32IN ? 4 : 2;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.pcm);
        if (tid == 0 && g.bulk_bytes > 0) {
            simt::fence_proxy_async();
            simt::mbar_expect_tx(sm.mbar, (uint32_t)g.bulk_bytes);
            simt::bulk_g2s(sm.raw, src + g.bulk_src * esz, (uint32_t)g.bulk_bytes, sm.mbar);
        }
        for (int64_t i = g.tail_lo + tid; i < g.tail_hi; i += kMfccThreads) {
            if (F32IN) reinterpret_cast<float*>(sm.raw)[i - g.a0s] = reinterpret_cast<const float*>(src)[i];
            else reinterpret_cast<int16_t*>(sm.raw)[i - g.a0s] = reinterpret_cast<const int16_t*>(src)[i];
        }
    };
    issue_chunk(0);

    const int nfull = NFULL >= 0 ? NFULL : (p.frame_len >> 5);
    uint32_t parity = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int frame0 = v_lo + c * kFramesPerPass;
        const ChunkGeom g = chunk_geom<F32IN>(p, start, S, frame0);
        if (g.bulk_bytes > 0) { simt::mbar_wait(sm.mbar, parity); parity ^= 1; }
        if (c == 0) simt::cta_sync();  // chunk-0 tail stores -> visible (later chunks: the loop-end barrier)

        // ---- raw int16 -> pre-emphasised fp32, 8 samples per thread.  Sample q of the staging area (packed sample
        // a0s + q) lands in plane q&1 at index q>>1.  Reference: y[0] = x[0], y[n] = x[n] - c*x[n-1] (sigproc.py:185), zeros past the end (:84-87).
        {
            const int rel = (int)(g.a0s - start);  // utterance sample index of raw[0] (may be negative)
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
        if (c + 1 < nchunks) issue_chunk(c + 1);  // raw is free again: overlap the next load with the FFTs

        // ---- one frame pair per 16-lane group; a warp (two groups) is active or idle as a whole.  Within a warp
        // group 0 packs frames (f, f+2) and group 1 packs (f+1, f+3): the two groups' loads then fall into
        // disjoint halves of the 32 banks (frame step/2 = 80 words = 16 mod 32 for the 10 ms step).
        const int fl = 4 * (tid >> 5) + (grp & 1);   // chunk-local index of the frame in the .x halves; .y = fl + 2
        const int vA = frame0 + fl;
        if (frame0 + 4 * (tid >> 5) <= v_hi) {
            cpx2 x[16];
            {
                const int fb0 = (int)(start - g.a0s + g.s0) + fl * p.frame_step;   // staging index of the frame's sample 0
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
            }
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

            // ---- real-FFT split, pairwise: bins k = lane + 16 r (r < 8) and 256 - k share one butterfly.
            // partner Z[(256-k) mod 256] lives in lane (16-lane)&15, register 15-r (lane 0: register (16-r)&15).
            const int src = (16 - lane) & 15;
            const float sc = p.pow_scale;  // 1 / (4 * NFFT)
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
                } else {
                    scr[128] = pa;
                }
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

            // ---- mel filterbank as sums over sub-ranges of the triangle edges (see mfcc_tables.h): the weights
            // are generated arithmetically, each power bin is read exactly once.
            float2 upv[kMaxTasks], dnv[kMaxTasks];
            float2 esum = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < kMaxTasks; ++t) {
                const int si = task[lane * kMaxTasks + t];
                float2 up = make_float2(0.f, 0.f), dn = up;
                if (si >= 0) {
                    const int len = subi[5 * si + 1];
                    const float2* pp = scr + subi[5 * si];
                    float fi = subf[5 * si + 2], gi = subf[5 * si + 3];
                    const float inv = subf[5 * si + 4];
#pragma unroll 4
                    for (int k = 0; k < len; ++k) {
                        const float2 pw = pp[k];
                        up = f2fmas(pw, fi, up);
                        dn = f2fmas(pw, gi, dn);
                        fi += 1.f; gi -= 1.f;
                    }
                    up = f2muls(up, inv); dn = f2muls(dn, inv);
                    esum = f2add(esum, f2add(up, dn));   // rising + falling weights sum to 1: this is sum(P) of the sub-range
                }
                upv[t] = up; dnv[t] = dn;
            }
            // total frame energy (reference base.py:25): reduce over the group
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) {
                esum.x += simt::shfl16(esum.x, lane ^ m);
                esum.y += simt::shfl16(esum.y, lane ^ m);
            }
            simt::group_sync();  // all lanes are done reading P before the staging area overwrites it
            float2* segu = scr;                   // [kMaxSubs]
            float2* segd = scr + kMaxSubs;        // [kMaxSubs]
            float2* lmel = scr + 2 * kMaxSubs;    // [nfilt + 1]; the last entry is log(energy)
#pragma unroll
            for (int t = 0; t < kMaxTasks; ++t) {
                const int si = task[lane * kMaxTasks + t];
                if (si >= 0) { segu[si] = upv[t]; segd[si] = dnv[t]; }
            }
            simt::group_sync();
            const float eps64 = 2.220446049250313e-16f;  // numpy.finfo(float64).eps, reference base.py:26,30
            for (int j = lane; j <= p.nfilt; j += 16) {
                float2 f = esum;
                if (j < p.nfilt) {   // filter j = rising part over range j+1 + falling part over range j+2
                    f = make_float2(0.f, 0.f);
                    const int a0 = rsub[2 * (j + 1)], an = rsub[2 * (j + 1) + 1];
                    for (int q = 0; q < an; ++q) f = f2add(f, segu[a0 + q]);
                    const int b0 = rsub[2 * (j + 2)], bn = rsub[2 * (j + 2) + 1];
                    for (int q = 0; q < bn; ++q) f = f2add(f, segd[b0 + q]);
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
    // ---- epilogue: delta (clamped on the utterance), delta-delta (clamped on the delta array), stores.
    // Three uniform passes (no divergent per-column work); flat indices advance by the CTA size without any
    // division: (row, col) += (128 / w, 128 % w).
    float* dbuf = sm.fbuf;                                   // aliases fbuf+raw: no copy is in flight any more
    float* ddbuf = reinterpret_cast<float*>(sm.scratch);     // aliases the FFT scratch
    const int u_lo = tile.f0 - N > 0 ? tile.f0 - N : 0;
    const int u_hi = tile.f0 + tile.nf - 1 + N < F - 1 ? tile.f0 + tile.nf - 1 + N : F - 1;
    const int qstep = kMfccThreads / numcep, rstep = kMfccThreads % numcep;
    {
        const int nd = (u_hi - u_lo + 1) * numcep;
        int uu = u_lo + tid / numcep, cc = tid % numcep;
        for (int i = tid; i < nd; i += kMfccThreads) {
            float acc = 0.f;
            for (int n = 1; n <= N; ++n) {
                int hi = uu + n; if (hi > F - 1) hi = F - 1;
                int lo = uu - n; if (lo < 0) lo = 0;
                acc = dsp_fmaf((float)n, sm.mfcc[(hi - v_lo) * numcep + cc] - sm.mfcc[(lo - v_lo) * numcep + cc], acc);
            }
            dbuf[i] = acc * p.delta_scale;
            uu += qstep; cc += rstep;
            if (cc >= numcep) { cc -= numcep; ++uu; }
        }
    }
    simt::cta_sync();
    {
        const int ndd = tile.nf * numcep;
        int tt = tile.f0 + tid / numcep, cc = tid % numcep;
        for (int i = tid; i < ndd; i += kMfccThreads) {
            float acc = 0.f;
            for (int n = 1; n <= N; ++n) {
                int hi = tt + n; if (hi > F - 1) hi = F - 1;
                int lo = tt - n; if (lo < 0) lo = 0;
                acc = dsp_fmaf((float)n, dbuf[(hi - u_lo) * numcep + cc] - dbuf[(lo - u_lo) * numcep + cc], acc);
            }
            ddbuf[i] = acc * p.delta_scale;
            tt += qstep; cc += rstep;
            if (cc >= numcep) { cc -= numcep; ++tt; }
        }
    }
    simt::cta_sync();
    {
        const int width = 3 * numcep;
        const int nout = tile.nf * width;
        float* outp = p.out + (row0 + tile.f0) * width;
        const float* src0 = sm.mfcc + (tile.f0 - v_lo) * numcep;
        const float* src1 = dbuf + (tile.f0 - u_lo) * numcep;
        const int qs = kMfccThreads / width, rs = kMfccThreads % width;
        int row = tid / width, col = tid % width;
        for (int i = tid; i < nout; i += kMfccThreads) {
            const float* s = src0; int cc = col;
            if (col >= numcep) { s = src1; cc = col - numcep; }
            if (col >= 2 * numcep) { s = ddbuf; cc = col - 2 * numcep; }
            outp[i] = s[row * numcep + cc];
            row += qs; col += rs;
            if (col >= width) { col -= width; ++row; }
        }
    }
}

}  // namespace dspfe
