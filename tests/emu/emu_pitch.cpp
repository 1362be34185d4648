// Emulator entry point for K4/K5/K6 (pitch).  TEST INFRASTRUCTURE ONLY.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../dsp-speech-recognition_b200/csrc/dspfe_types.h"
#include "../../dsp-speech-recognition_b200/csrc/pitch_kernel.cuh"
#include "../../dsp-speech-recognition_b200/csrc/pitch_tables.h"
#include "../../include/dspfe.h"

namespace emu { bool run_cta(int bid, int nthreads, void (*body)(void*), void* arg); }
using namespace dspfe;

namespace {
struct Args { PitchParams p; std::vector<unsigned char>* smem; int64_t total_frames; };
void frame_body(void* a) {
    Args* A = (Args*)a;
    const int w = simt::tid() >> 5;
    const bool quad = A->p.mode != 0 && acr_short_frames(A->p.frame_len, A->p.row_len);
    const int per = quad ? 4 : 2;
    const int64_t h0 = per * ((int64_t)simt::bid() * kPitchWarps + w);
    if (h0 >= A->p.slot_off[A->p.n_utt]) return;
    int nvalid; int64_t g0;
    slot_unit_from_desc(A->p, h0, per, g0, nvalid);
    if (nvalid == 0) return;
    // the device kernel stages the twiddles and the decimator pattern in shared memory; the emulator reads them in place
    unsigned char* wsm = A->smem->data() + w * kWarpSmemBytes;
    if (A->p.mode == 0) pitch_fft_pair<0>(A->p, h0, g0, nvalid > 1, wsm, A->p.tab, A->p.tab + kTabW32);
    else if (quad) pitch_acr_quad(A->p, h0, g0, nvalid, wsm, A->p.tab, A->p.tab + kTabW32);
    else pitch_fft_pair<1>(A->p, h0, g0, nvalid > 1, wsm, A->p.tab, A->p.tab + kTabW32);
}
void clip_body(void* a) {
    Args* A = (Args*)a;
    const int w = simt::tid() >> 5;
    const int64_t h0 = kClipRun * ((int64_t)simt::bid() * kPitchWarps + w);
    if (h0 >= A->p.slot_off[A->p.n_utt]) return;
    unsigned char* wsm = A->smem->data() + w * kClipWarpSmemBytes;
    const bool i16 = clip_i16_keys(A->p);
    if (A->p.frame_len <= 320) {   // e.g. the 300-sample frames of model.py:92
        if (i16) pitch_clip_run<10, true>(A->p, h0, wsm, A->p.ds_idx); else pitch_clip_run<10, false>(A->p, h0, wsm, A->p.ds_idx);
    } else {
        if (i16) pitch_clip_run<16, true>(A->p, h0, wsm, A->p.ds_idx); else pitch_clip_run<16, false>(A->p, h0, wsm, A->p.ds_idx);
    }
}
void track_body(void* a) {
    Args* A = (Args*)a;
    float* chunk = reinterpret_cast<float*>(A->smem->data());
    int* sc = reinterpret_cast<int*>(chunk + (track_chunk(A->p.row_len) + 1) * A->p.row_len);
    pitch_track_cta(A->p, chunk, sc, reinterpret_cast<double*>(sc + track_chunk(A->p.row_len) * kPeakLags));
}
void feature_body(void* a) {
    Args* A = (Args*)a;
    pitch_feature_warp(A->p, simt::bid(), reinterpret_cast<double*>(A->smem->data()));
}
}  // namespace

// Same contract as dspfe_pitch but host pointers and synchronous.  Returns total frames, <0 on error.
extern "C" long long emu_pitch(const dspfe_pitch_params* q, const void* pcm, int in_f32, const long long* offsets, const int* trim,
                               int n_utt, double* pitch, int* lag, double* feat, float* rows, float* rows_smoothed, int* score,
                               long long* frame_off_out, long long max_frames, char* errbuf, int errcap) {
    PitchParams p;
    std::vector<float2> tab;
    std::string err;
    if (build_pitch_tables(*q, p, tab, err)) { std::snprintf(errbuf, errcap, "%s", err.c_str()); return -2; }
    std::vector<int64_t> seg_start(n_utt + 1), frame_off(n_utt + 1), slot_off(n_utt + 1);
    std::vector<int32_t> seg_len(n_utt + 1), ds_len(n_utt + 1);
    int64_t fo = 0, so = 0;
    for (int u = 0; u < n_utt; ++u) {
        int64_t a = offsets[u], len = offsets[u + 1] - offsets[u];
        if (trim) {
            int64_t l = trim[2 * u], r = trim[2 * u + 1];
            if (l < 0) l = 0; if (r < 0) r = 0;
            if (l > len) l = len; if (r > len) r = len;
            a += l; len = r > l ? r - l : 0;
        }
        seg_start[u] = a; seg_len[u] = (int32_t)len;
        ds_len[u] = (int32_t)ds_length(len, p.ds_idx, p.ds_in, p.ds_out);
        frame_off[u] = fo; slot_off[u] = so;
        const int64_t nf = num_frames(ds_len[u], p.frame_len, p.frame_step);
        fo += nf; so += (nf + kClipRun - 1) / kClipRun * kClipRun;
    }
    frame_off[n_utt] = fo; slot_off[n_utt] = so;
    if (fo > max_frames) { std::snprintf(errbuf, errcap, "outputs too small"); return -1; }
    if (frame_off_out) std::memcpy(frame_off_out, frame_off.data(), (n_utt + 1) * sizeof(int64_t));
    std::vector<float> rows_own; std::vector<double> amp(fo), pitch_own(fo), scratch(3 * fo); std::vector<int32_t> lag_own(fo);
    if (!rows) { rows_own.resize((size_t)fo * p.row_len); rows = rows_own.data(); }
    p.pcm = pcm; p.in_f32 = in_f32; p.total_samples = offsets[n_utt]; p.offsets = (const int64_t*)offsets; p.trim = trim; p.n_utt = n_utt;
    p.tab = tab.data(); p.frame_off = frame_off.data(); p.slot_off = slot_off.data(); p.seg_start = seg_start.data(); p.seg_len = seg_len.data();
    p.ds_len = ds_len.data(); p.rows = rows; p.rows_out = rows_smoothed; p.score = score; p.frame_amp = amp.data();
    p.pitch = pitch ? pitch : pitch_own.data(); p.lag = lag ? lag : lag_own.data(); p.feat = feat; p.scratch = scratch.data();
    p.max_frames = fo;
    std::vector<float2> clip((size_t)((so + 1) / 2) * 512);
    p.clip = clip.data();
    std::vector<int4> run_desc((size_t)(so / kClipRun + 1));
    p.run_desc = run_desc.data();
    std::vector<unsigned char> smem0(kPitchWarps * kClipWarpSmemBytes + 64);
    Args A0{p, &smem0, fo};
    for (int64_t b = 0; b * kClipRun * kPitchWarps < so + kClipRun * kPitchWarps; ++b) {
        std::memset(smem0.data(), 0xCD, smem0.size());
        if (!emu::run_cta((int)b, 32 * kPitchWarps, clip_body, &A0)) { std::snprintf(errbuf, errcap, "deadlock in clip CTA %lld", (long long)b); return -3; }
    }
    std::vector<unsigned char> smem(kPitchWarps * kWarpSmemBytes + 64);
    Args A{p, &smem, fo};
    for (int64_t b = 0; b * 2 * kPitchWarps < so + 2 * kPitchWarps; ++b) {
        std::memset(smem.data(), 0xCD, smem.size());
        if (!emu::run_cta((int)b, 32 * kPitchWarps, frame_body, &A)) { std::snprintf(errbuf, errcap, "deadlock in frame CTA %lld", (long long)b); return -3; }
    }
    std::vector<unsigned char> smem2((track_chunk(p.row_len) + 1) * p.row_len * sizeof(float) + kTrackChunk * kPeakLags * sizeof(int) + kTrackMaxFrames * 12 + (kTrackThreads / 32) * kTrackListPerWarp * 2 + 8 + 64);
    Args B{p, &smem2, fo};
    for (int u = 0; u < n_utt; ++u) {
        std::memset(smem2.data(), 0xCD, smem2.size());
        if (!emu::run_cta(u, kTrackThreads, track_body, &B)) { std::snprintf(errbuf, errcap, "deadlock in track CTA %d", u); return -3; }
    }
    if (feat) {
        std::vector<unsigned char> smem3(5 * kFeatMaxFrames * sizeof(double) + 64);
        Args C{p, &smem3, fo};
        for (int u = 0; u < n_utt; ++u) {
            std::memset(smem3.data(), 0xCD, smem3.size());
            if (!emu::run_cta(u, 32, feature_body, &C)) { std::snprintf(errbuf, errcap, "deadlock in feature CTA %d", u); return -3; }
        }
    }
    return fo;
}

// the dual exact median of K4a-1 on two caller-supplied frames (len <= 512): out[0], out[1] = np.median(frame[frame >= 0])
namespace {
struct MedArgs { const float* a; const float* b; int len; int i16; float out[2]; };
void median_body(void* q) {
    MedArgs* M = (MedArgs*)q;
    const int lane = simt::tid() & 31;
    float xa[16], xb[16];
    for (int t = 0; t < 16; ++t) { const int n = 32 * t + lane; xa[t] = n < M->len ? M->a[n] : 0.f; xb[t] = n < M->len ? M->b[n] : 0.f; }
    const float2 m = M->i16 ? warp_median_nonneg2_i16<16>(xa, xb, M->len, lane) : warp_median_nonneg2<16>(xa, xb, M->len, lane);
    if (lane == 0) { M->out[0] = m.x; M->out[1] = m.y; }
}
}  // namespace
extern "C" int emu_median2(const float* a, const float* b, int len, int i16, float* out2) {
    MedArgs M{a, b, len, i16, {0.f, 0.f}};
    if (!emu::run_cta(0, 32, median_body, &M)) return -1;
    out2[0] = M.out[0]; out2[1] = M.out[1];
    return 0;
}
