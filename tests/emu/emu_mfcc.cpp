// Emulator entry point for K1 (fused MFCC + delta + delta-delta).  TEST INFRASTRUCTURE ONLY.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../dsp-speech-recognition_b200/csrc/mfcc_kernel.cuh"
#include "../../dsp-speech-recognition_b200/csrc/mfcc_tables.h"
#include "../../dsp-speech-recognition_b200/csrc/mfcc_long_tables.h"
#include "../../include/dspfe.h"

namespace emu { bool run_cta(int bid, int nthreads, void (*body)(void*), void* arg); }
using namespace dspfe;

namespace {
struct Args { MfccParams p; bool has_win; bool f32; std::vector<unsigned char>* smem; };
void body(void* a) {
    Args* A = (Args*)a;
    const int nfull = A->p.frame_len >> 5;
    if (A->p.tri && A->p.frame_len > 512) {   // K1T LONG
        if (A->f32) { if (A->has_win) mfcc_cta<true, -1, true, 0, true, true>(A->p, A->smem->data()); else mfcc_cta<false, -1, true, 0, true, true>(A->p, A->smem->data()); return; }
        if (A->has_win) mfcc_cta<true, -1, false, 0, true, true>(A->p, A->smem->data()); else mfcc_cta<false, -1, false, 0, true, true>(A->p, A->smem->data());
        return;
    }
    if (A->p.tri) {   // K1T (nfft = 1536, frames of at most 512 samples)
        if (A->f32) { if (A->has_win) mfcc_cta<true, -1, true, 0, true>(A->p, A->smem->data()); else mfcc_cta<false, -1, true, 0, true>(A->p, A->smem->data()); return; }
        if (A->has_win) { if (nfull == 15) mfcc_cta<true, 15, false, 0, true>(A->p, A->smem->data()); else mfcc_cta<true, -1, false, 0, true>(A->p, A->smem->data()); }
        else mfcc_cta<false, -1, false, 0, true>(A->p, A->smem->data());
        return;
    }
    if (A->f32) { if (A->has_win) mfcc_cta<true, -1, true, 0>(A->p, A->smem->data()); else mfcc_cta<false, -1, true, 0>(A->p, A->smem->data()); return; }
    if (A->has_win) { if (nfull == 12) mfcc_cta<true, 12, false, 0>(A->p, A->smem->data()); else mfcc_cta<true, -1, false, 0>(A->p, A->smem->data()); }
    else { if (nfull == 12) mfcc_cta<false, 12, false, 0>(A->p, A->smem->data()); else mfcc_cta<false, -1, false, 0>(A->p, A->smem->data()); }
}
}  // namespace

// Same contract as dspfe_mfcc_delta but host pointers and synchronous.  Returns rows written, <0 on error.
extern "C" long long emu_mfcc_delta(const dspfe_mfcc_params* q, const void* pcm, int in_f32, long long total_samples,
                                    const long long* offsets, const int* trim, int n_utt, float* out, long long max_rows,
                                    long long* frame_off_out, char* errbuf, int errcap) {
    MfccConfig c;
    c.samplerate = q->samplerate; c.frame_len = q->frame_len; c.frame_step = q->frame_step; c.nfft = q->nfft;
    c.nfilt = q->nfilt; c.numcep = q->numcep; c.ceplifter = q->ceplifter; c.append_energy = q->append_energy;
    c.delta_n = q->delta_n; c.seg_frames = q->seg_frames > 0 ? q->seg_frames : 256;
    c.preemph = q->preemph; c.lowfreq = q->lowfreq; c.highfreq = q->highfreq;
    if (q->window) c.window.assign(q->window, q->window + q->frame_len);
    MfccParams p; std::memset(&p, 0, sizeof(p));
    std::string err;
    std::vector<float> blob = build_mfcc_tables(c, p, err, in_f32 != 0);
    if (!err.empty()) { std::snprintf(errbuf, errcap, "%s", err.c_str()); return -2; }
    // host mirror of prep_kernel
    std::vector<int64_t> seg_start(n_utt + 1), frame_off(n_utt + 1);
    std::vector<int32_t> seg_len(n_utt + 1);
    std::vector<Tile> tiles;
    int64_t fo = 0;
    for (int u = 0; u < n_utt; ++u) {
        int64_t a = offsets[u], len = offsets[u + 1] - offsets[u];
        if (trim) {
            int64_t l = trim[2 * u], r = trim[2 * u + 1];
            if (l < 0) l = 0; if (r < 0) r = 0;
            if (l > len) l = len; if (r > len) r = len;
            a += l; len = r > l ? r - l : 0;
        }
        seg_start[u] = a; seg_len[u] = (int32_t)len;
        const int F = (int)num_frames(len, c.frame_len, c.frame_step);
        const int T = (F + c.seg_frames - 1) / c.seg_frames, per = (F + T - 1) / T;
        frame_off[u] = fo;
        for (int t = 0; t < T; ++t) { Tile tl; tl.utt = u; tl.f0 = t * per; tl.nf = std::min(per, F - tl.f0); tl.pad = 0; tiles.push_back(tl); }
        fo += F;
    }
    frame_off[n_utt] = fo;
    if (fo > max_rows) { std::snprintf(errbuf, errcap, "out too small"); return -1; }
    if (frame_off_out) std::memcpy(frame_off_out, frame_off.data(), (n_utt + 1) * sizeof(int64_t));
    int32_t ntiles = (int32_t)tiles.size();
    p.pcm = pcm; p.total_samples = total_samples; p.seg_start = seg_start.data(); p.seg_len = seg_len.data();
    p.frame_off = frame_off.data(); p.tiles = tiles.data(); p.ntiles = &ntiles; p.tables = blob.data(); p.out = out;
    std::vector<unsigned char> smem(p.sm_total + 64);
    Args A{p, !c.window.empty(), in_f32 != 0, &smem};
    for (int b = 0; b < ntiles + 1; ++b) {   // +1: exercises the early-exit path of surplus CTAs
        std::memset(smem.data(), 0xCD, smem.size());   // poison: uninitialised shared memory shows up as garbage
        if (!emu::run_cta(b, kMfccThreads, body, &A)) { std::snprintf(errbuf, errcap, "deadlock in CTA %d", b); return -3; }
    }
    return fo;
}


// ---- K1L (nfft = 1536) on the emulator
namespace {
struct LArgs { MfccLongParams p; std::vector<unsigned char>* smem; int64_t total; };
void long_body(void* a) {
    LArgs* A = (LArgs*)a;
    const int w = simt::tid() >> 5;
    const int64_t g0 = 2 * ((int64_t)simt::bid() * kLongWarps + w);
    if (g0 >= A->total) return;
    const float2* tws = reinterpret_cast<const float2*>(A->p.tab);
    mfcc_long_pair(A->p, g0, A->total, A->smem->data() + w * A->p.warp_smem, tws, tws + kLtW32 / 2);
}
}  // namespace

extern "C" long long emu_mfcc_long(const dspfe_mfcc_params* q, const void* pcm, int in_f32, const long long* offsets, int n_utt,
                                   float* out, long long max_rows, long long* frame_off_out, char* errbuf, int errcap) {
    MfccConfig c;
    c.samplerate = q->samplerate; c.frame_len = q->frame_len; c.frame_step = q->frame_step; c.nfft = q->nfft;
    c.nfilt = q->nfilt; c.numcep = q->numcep; c.ceplifter = q->ceplifter; c.append_energy = q->append_energy;
    c.delta_n = q->delta_n; c.preemph = q->preemph; c.lowfreq = q->lowfreq; c.highfreq = q->highfreq;
    if (q->window) c.window.assign(q->window, q->window + q->frame_len);
    const std::string err = mfcc_long_config_check(c);
    if (!err.empty()) { std::snprintf(errbuf, errcap, "%s", err.c_str()); return -2; }
    const std::vector<float> tab = build_long_tables(c);
    std::vector<int64_t> seg_start(n_utt + 1), frame_off(n_utt + 1);
    std::vector<int32_t> seg_len(n_utt + 1);
    int64_t fo = 0;
    for (int u = 0; u < n_utt; ++u) {
        seg_start[u] = offsets[u]; seg_len[u] = (int32_t)(offsets[u + 1] - offsets[u]);
        frame_off[u] = fo; fo += num_frames(seg_len[u], c.frame_len, c.frame_step);
    }
    frame_off[n_utt] = fo;
    if (fo > max_rows) { std::snprintf(errbuf, errcap, "out too small"); return -1; }
    if (frame_off_out) std::memcpy(frame_off_out, frame_off.data(), (n_utt + 1) * sizeof(int64_t));
    std::vector<float> cep((size_t)fo * c.numcep);
    MfccLongParams p;
    p.pcm = pcm; p.in_f32 = in_f32; p.seg_start = seg_start.data(); p.seg_len = seg_len.data(); p.frame_off = frame_off.data(); p.n_utt = n_utt;
    p.frame_len = c.frame_len; p.frame_step = c.frame_step; p.nfilt = c.nfilt; p.numcep = c.numcep; p.append_energy = c.append_energy;
    p.preemph = (float)c.preemph; p.tab = tab.data(); p.mfcc = cep.data(); p.max_frames = fo;
    long_fill_size_params(p, c.nfft, 0, 0);
    long_fill_mel_params(p, tab.data());
    std::vector<unsigned char> smem(kLongWarps * p.warp_smem + 64);
    LArgs A{p, &smem, fo};
    for (int64_t b = 0; b * 2 * kLongWarps < fo + 2 * kLongWarps; ++b) {
        std::memset(smem.data(), 0xCD, smem.size());
        if (!emu::run_cta((int)b, 32 * kLongWarps, long_body, &A)) { std::snprintf(errbuf, errcap, "deadlock in CTA %lld", (long long)b); return -3; }
    }
    int den = 0; for (int i = 1; i <= c.delta_n; ++i) den += i * i;
    for (int64_t i = 0; i < fo * c.numcep; ++i) delta_batch_thread(cep.data(), frame_off.data(), n_utt, c.numcep, c.delta_n, (float)(1.0 / (2.0 * den)), i, out);
    return fo;
}
