// C ABI of the endpoint-detection path (include/dspfe.h, "endpoint" section).  Kernels: endpoint_kernel.cuh.
#include <cstring>
#include <vector>

#include "abi_common.h"
#include "dspfe_types.h"
#include "endpoint_kernel.cuh"

using namespace dspfe;

struct dspfe_endpoint_plan {
    dspfe_endpoint_params prm;
    EpRule rule;
    int frame_len, frame_step, q, rem;
    // workspaces
    int64_t cap_utt = 0, cap_blocks = 0, cap_frames = 0;
    int64_t *frame_off = nullptr, *block_off = nullptr; int32_t* order = nullptr;
    int32_t *blk = nullptr, *asum = nullptr, *zcr = nullptr;
    // host-path staging
    cudaStream_t stream = nullptr;
    int16_t* d_pcm = nullptr; int64_t cap_samples = 0;
    int64_t* d_off = nullptr; int32_t* d_lr = nullptr; int64_t cap_hutt = 0;
};

namespace {

int fill_rule(const dspfe_endpoint_params& q, EpRule& r, int& frame_len, int& frame_step) {
    if (q.samplerate <= 0 || !(q.cfg_frame > 0) || !(q.cfg_step > 0)) return fail(DSPFE_ERR_INVALID_ARG, "bad endpoint parameters");
    r.cfg_frame = q.cfg_frame; r.cfg_step = q.cfg_step; r.mh1 = q.mh1; r.mh2 = q.mh2; r.th = q.th;
    r.l_sil = q.l_sil; r.r_sil = q.r_sil; r.sigma = q.sigma; r.zcr_max_shift = q.zcr_max_shift; r.zcr_r_sil = q.zcr_r_sil;
    r.min_span = q.min_span; r.rate = q.samplerate;
    frame_len = (int)((double)q.samplerate * q.cfg_frame);   // to_frames: int(rate * t)   (sigproc.py:19)
    frame_step = (int)(q.cfg_step * (double)q.samplerate);   //            int(step * rate)
    if (frame_len < 1 || frame_step < 1) return fail(DSPFE_ERR_UNSUPPORTED, "endpoint frame length/step below one sample");
    if ((int)(q.l_sil / q.cfg_step) + (int)(q.r_sil / q.cfg_step) > 64 || (int)(q.zcr_r_sil / q.cfg_step) > 64)
        return fail(DSPFE_ERR_UNSUPPORTED, "silence windows longer than 64 frames are not supported");
    return DSPFE_OK;
}

int ensure(dspfe_endpoint_plan* pl, int64_t n_utt, int64_t blocks, int64_t frames) {
    if (n_utt + 1 > pl->cap_utt) {
        cudaFree(pl->frame_off); cudaFree(pl->block_off); cudaFree(pl->order); pl->frame_off = pl->block_off = nullptr; pl->order = nullptr; pl->cap_utt = 0;
        CUDA_TRY(cudaMalloc(&pl->frame_off, (n_utt + 1) * sizeof(int64_t)));
        CUDA_TRY(cudaMalloc(&pl->order, (n_utt + 1) * sizeof(int32_t)));
        CUDA_TRY(cudaMalloc(&pl->block_off, (n_utt + 1) * sizeof(int64_t)));
        pl->cap_utt = n_utt + 1;
    }
    if (blocks > pl->cap_blocks) {
        cudaFree(pl->blk); pl->blk = nullptr; pl->cap_blocks = 0;
        CUDA_TRY(cudaMalloc(&pl->blk, blocks * 6 * sizeof(int32_t)));
        pl->cap_blocks = blocks;
    }
    if (frames > pl->cap_frames) {
        cudaFree(pl->asum); cudaFree(pl->zcr); pl->asum = pl->zcr = nullptr; pl->cap_frames = 0;
        CUDA_TRY(cudaMalloc(&pl->asum, frames * sizeof(int32_t)));
        CUDA_TRY(cudaMalloc(&pl->zcr, frames * sizeof(int32_t)));
        pl->cap_frames = frames;
    }
    return DSPFE_OK;
}

}  // namespace

extern "C" {

void dspfe_endpoint_params_default(dspfe_endpoint_params* p, int32_t samplerate) {
    if (!p) return;
    p->samplerate = samplerate; p->min_span = 50;
    p->cfg_frame = 0.03; p->cfg_step = 0.01; p->mh1 = 0.25; p->mh2 = 0.125; p->th = 0.100; p->l_sil = 0.100; p->r_sil = 0.100;
    p->sigma = 3.0; p->zcr_max_shift = 0.400; p->zcr_r_sil = 0.100;
}

int dspfe_endpoint_decide_host(const dspfe_endpoint_params* p, const int32_t* asum, const int32_t* zcr, int32_t n_frames, int32_t* lr) {
    if (!p || !asum || !zcr || !lr || n_frames < 1) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    EpRule r; int fl, fs;
    int rc = fill_rule(*p, r, fl, fs);
    if (rc) return rc;
    endpoint_decide(asum, zcr, n_frames, fl, r, lr);
    return DSPFE_OK;
}

int dspfe_amplitude_rule_host(const dspfe_endpoint_params* p, const double* amp, int32_t n_frames, double mh,
                              int32_t* segs, int32_t seg_cap, int32_t* n_segs) {
    if (!p || !amp || !n_segs || n_frames < 1 || seg_cap < 1 || !segs) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    EpRule r; int fl, fs;
    int rc = fill_rule(*p, r, fl, fs);
    if (rc) return rc;
    int left, right;
    const int n = amplitude_rule(AmpFromF64{amp}, n_frames, r, mh, &left, &right, segs, seg_cap);
    if (n == 0) { segs[0] = 0; segs[1] = n_frames; *n_segs = 1; } else { *n_segs = n; }   // reference: [(0, len(amp))]
    return DSPFE_OK;
}

int dspfe_amplitude_rule_gated_host(const dspfe_endpoint_params* p, const double* amp, const int32_t* gate, int32_t n_frames, double mh,
                                    int32_t* segs, int32_t seg_cap, int32_t* n_segs) {
    if (!p || !amp || !gate || !n_segs || n_frames < 1 || seg_cap < 1 || !segs) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    EpRule r; int fl, fs;
    int rc = fill_rule(*p, r, fl, fs);
    if (rc) return rc;
    int left, right;
    const int n = amplitude_rule(AmpFromF64{amp}, n_frames, r, mh, &left, &right, segs, seg_cap, GateFromFlags{gate});
    if (n == 0) { segs[0] = 0; segs[1] = n_frames; *n_segs = 1; } else { *n_segs = n; }
    return DSPFE_OK;
}

int dspfe_acr_gate_rows_f64(const double* d_frames, int64_t n_rows, int32_t len, int32_t samplerate, int32_t* d_gate, void* stream) {
    if (n_rows < 0 || len < 1 || samplerate < 1 || (n_rows > 0 && (!d_frames || !d_gate))) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    if (n_rows == 0) return DSPFE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    acr_gate_rows_kernel<<<(unsigned)((n_rows + 3) / 4), 128, 0, st>>>(d_frames, n_rows, len, samplerate / 500, samplerate / 50, d_gate);
    LAUNCH_CHECK("acr_gate_rows_kernel", st);
    return DSPFE_OK;
}

int dspfe_zcr_rule_host(const dspfe_endpoint_params* p, const double* zcr, int32_t n_frames, double l_sil, int32_t left,
                        int32_t right, int32_t* out_jk) {
    if (!p || !zcr || !out_jk || n_frames < 1 || left < 0 || left >= n_frames || right < 0 || right > n_frames)
        return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    EpRule r; int fl, fs;
    int rc = fill_rule(*p, r, fl, fs);
    if (rc) return rc;
    int j, k;
    zcr_rule(AmpFromF64{zcr}, n_frames, r, l_sil, left, right, &j, &k);
    out_jk[0] = j; out_jk[1] = k;
    return DSPFE_OK;
}

int dspfe_endpoint_create(const dspfe_endpoint_params* p, dspfe_endpoint_plan** plan) {
    if (!p || !plan) return fail(DSPFE_ERR_INVALID_ARG, "null argument");
    *plan = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(DSPFE_ERR_CUDA, "no CUDA device: libdspfe has no CPU fallback");
    dspfe_endpoint_plan* pl = new (std::nothrow) dspfe_endpoint_plan();
    if (!pl) return fail(DSPFE_ERR_NOMEM, "out of host memory");
    pl->prm = *p;
    int rc = fill_rule(*p, pl->rule, pl->frame_len, pl->frame_step);
    if (rc) { delete pl; return rc; }
    pl->q = pl->frame_len / pl->frame_step; pl->rem = pl->frame_len % pl->frame_step;
    *plan = pl;
    return DSPFE_OK;
}

void dspfe_endpoint_destroy(dspfe_endpoint_plan* pl) {
    if (!pl) return;
    cudaFree(pl->frame_off); cudaFree(pl->block_off); cudaFree(pl->order); cudaFree(pl->blk); cudaFree(pl->asum); cudaFree(pl->zcr);
    cudaFree(pl->d_pcm); cudaFree(pl->d_off); cudaFree(pl->d_lr);
    if (pl->stream) cudaStreamDestroy(pl->stream);
    delete pl;
}

int dspfe_endpoint_reserve(dspfe_endpoint_plan* pl, int64_t max_utt, int64_t max_total_samples) {
    if (!pl || max_utt < 0 || max_total_samples < 0) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    const int64_t frames = dspfe_endpoint_frames_bound(pl, max_total_samples, max_utt);
    return ensure(pl, max_utt, frames + max_utt * (pl->q + 1), frames);
}

int32_t dspfe_endpoint_frame_len(const dspfe_endpoint_plan* pl) { return pl ? pl->frame_len : -1; }
int32_t dspfe_endpoint_frame_step(const dspfe_endpoint_plan* pl) { return pl ? pl->frame_step : -1; }

int64_t dspfe_endpoint_frames_bound(const dspfe_endpoint_plan* pl, int64_t total_samples, int64_t n_utt) {
    if (!pl) return -1;
    // framesig's count is 1 + ceil((len - frame_len) / step) <= len / step + 1 per utterance when frame_len >= step, and
    // <= len / step + 2 when cfg.step exceeds cfg.frame (gaps between frames)
    return total_samples / pl->frame_step + (pl->frame_len >= pl->frame_step ? 1 : 2) * n_utt;
}

int dspfe_endpoint(dspfe_endpoint_plan* pl, const int16_t* d_pcm, int64_t total_samples, const int64_t* d_offsets, int32_t n_utt,
                   int32_t* d_lr, int32_t* d_asum, int32_t* d_zcr, int64_t* d_frame_off, int64_t max_frames, void* stream) {
    if (!pl || !d_offsets || !d_lr || n_utt < 0 || total_samples < 0) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    if (n_utt == 0) return DSPFE_OK;
    if (!d_pcm && total_samples > 0) return fail(DSPFE_ERR_INVALID_ARG, "d_pcm is null");
    if (((uintptr_t)d_pcm & 15) != 0) return fail(DSPFE_ERR_INVALID_ARG, "d_pcm must be 16-byte aligned");
    const int64_t frames_bound = dspfe_endpoint_frames_bound(pl, total_samples, n_utt);
    if ((d_asum || d_zcr) && max_frames < frames_bound) return fail(DSPFE_ERR_INVALID_ARG, "max_frames is below dspfe_endpoint_frames_bound()");
    const int64_t blocks_bound = frames_bound + (int64_t)n_utt * (pl->q + 1);
    int rc = ensure(pl, n_utt, blocks_bound, frames_bound);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    EpParams p;
    p.pcm = d_pcm; p.offsets = d_offsets; p.n_utt = n_utt; p.frame_len = pl->frame_len; p.frame_step = pl->frame_step;
    p.q = pl->q; p.rem = pl->rem; p.frame_off = d_frame_off ? d_frame_off : pl->frame_off; p.block_off = pl->block_off;
    p.blk = pl->blk; p.asum = d_asum ? d_asum : pl->asum; p.zcr = d_zcr ? d_zcr : pl->zcr; p.lr = d_lr;
    p.max_blocks = blocks_bound; p.max_frames = frames_bound; p.rule = pl->rule; p.order = pl->order;
    ep_prep_kernel<<<1, kEpPrepThreads, 0, st>>>(p);
    LAUNCH_CHECK("ep_prep_kernel", st);
    ep_block_kernel<<<(unsigned)((blocks_bound + 127) / 128), 128, 0, st>>>(p);
    LAUNCH_CHECK("ep_block_kernel", st);
    ep_frame_kernel<<<(unsigned)((frames_bound + 255) / 256), 256, 0, st>>>(p);
    LAUNCH_CHECK("ep_frame_kernel", st);
    ep_decide_kernel<<<(unsigned)((n_utt + kEpDecideWarps - 1) / kEpDecideWarps), 32 * kEpDecideWarps, 0, st>>>(p);
    LAUNCH_CHECK("ep_decide_kernel", st);
    return DSPFE_OK;
}

int dspfe_endpoint_robust(dspfe_endpoint_plan* pl, const int16_t* d_pcm, int64_t total_samples, const int64_t* d_offsets, int32_t n_utt,
                          int32_t* d_lr, void* stream) {
    if (!pl || !d_offsets || !d_lr || n_utt < 0 || total_samples < 0) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    if (n_utt == 0) return DSPFE_OK;
    if (!d_pcm && total_samples > 0) return fail(DSPFE_ERR_INVALID_ARG, "d_pcm is null");
    if (pl->frame_len > kEpGateMaxLen) return fail(DSPFE_ERR_UNSUPPORTED, "robust endpoint frames longer than 1536 samples are not built");
    // frame statistics exactly as the basic path (its own decision is discarded), then the gated rule
    int rc = dspfe_endpoint(pl, d_pcm, total_samples, d_offsets, n_utt, d_lr, nullptr, nullptr, nullptr, 0, stream);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t frames_bound = dspfe_endpoint_frames_bound(pl, total_samples, n_utt);
    EpParams p;
    p.pcm = d_pcm; p.offsets = d_offsets; p.n_utt = n_utt; p.frame_len = pl->frame_len; p.frame_step = pl->frame_step;
    p.q = pl->q; p.rem = pl->rem; p.frame_off = pl->frame_off; p.block_off = pl->block_off; p.blk = pl->blk; p.asum = pl->asum; p.zcr = pl->zcr;
    p.lr = d_lr; p.max_blocks = 0; p.max_frames = frames_bound; p.rule = pl->rule; p.order = pl->order;
    CUDA_TRY(cudaFuncSetAttribute(ep_decide_robust_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEpRobustSmem));
    CUDA_TRY(cudaFuncSetAttribute(ep_decide_robust_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ep_decide_robust_kernel<<<(unsigned)n_utt, 32 * kEpRobustWarps, kEpRobustSmem, st>>>(p);
    LAUNCH_CHECK("ep_decide_robust_kernel", st);
    return DSPFE_OK;
}

int dspfe_endpoint_robust_host(dspfe_endpoint_plan* pl, const int16_t* h_pcm, const int64_t* h_offsets, int32_t n_utt, int32_t* h_lr) {
    if (!pl || !h_offsets || !h_lr || n_utt < 0) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    if (n_utt == 0) return DSPFE_OK;
    // stage through the basic host path's buffers (it leaves pcm / offsets on the device), then re-decide with the gate
    int rc = dspfe_endpoint_host(pl, h_pcm, h_offsets, n_utt, h_lr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    const int64_t total = h_offsets[n_utt] - h_offsets[0];
    rc = dspfe_endpoint_robust(pl, pl->d_pcm, total, pl->d_off, n_utt, pl->d_lr, pl->stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h_lr, pl->d_lr, (int64_t)n_utt * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, pl->stream));
    CUDA_TRY(cudaStreamSynchronize(pl->stream));
    return DSPFE_OK;
}

int dspfe_endpoint_host(dspfe_endpoint_plan* pl, const int16_t* h_pcm, const int64_t* h_offsets, int32_t n_utt, int32_t* h_lr,
                        int32_t* h_asum, int32_t* h_zcr, int64_t* h_frame_off) {
    if (!pl || !h_offsets || !h_lr || n_utt < 0) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    if (n_utt == 0) return DSPFE_OK;
    const int64_t base = h_offsets[0], total = h_offsets[n_utt] - base;
    if (total < 0) return fail(DSPFE_ERR_INVALID_ARG, "offsets must be non-decreasing");
    if (!pl->stream) CUDA_TRY(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    if (total + 8 > pl->cap_samples) {
        cudaFree(pl->d_pcm); pl->d_pcm = nullptr; pl->cap_samples = 0;
        CUDA_TRY(cudaMalloc(&pl->d_pcm, (total + 8) * sizeof(int16_t)));
        pl->cap_samples = total + 8;
    }
    if (n_utt + 1 > pl->cap_hutt) {
        cudaFree(pl->d_off); cudaFree(pl->d_lr); pl->d_off = nullptr; pl->d_lr = nullptr; pl->cap_hutt = 0;
        CUDA_TRY(cudaMalloc(&pl->d_off, (n_utt + 1) * sizeof(int64_t)));
        CUDA_TRY(cudaMalloc(&pl->d_lr, (int64_t)n_utt * 2 * sizeof(int32_t)));
        pl->cap_hutt = n_utt + 1;
    }
    std::vector<int64_t> rel(n_utt + 1);
    int64_t frames = 0;
    for (int32_t u = 0; u <= n_utt; ++u) {
        rel[u] = h_offsets[u] - base;
        if (u < n_utt) {
            if (h_frame_off) h_frame_off[u] = frames;
            frames += num_frames(h_offsets[u + 1] - h_offsets[u], pl->frame_len, pl->frame_step);
        }
    }
    if (h_frame_off) h_frame_off[n_utt] = frames;
    cudaStream_t st = pl->stream;
    if (total > 0) CUDA_TRY(cudaMemcpyAsync(pl->d_pcm, h_pcm + base, total * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(pl->d_off, rel.data(), (n_utt + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    int rc = dspfe_endpoint(pl, pl->d_pcm, total, pl->d_off, n_utt, pl->d_lr, nullptr, nullptr, nullptr, 0, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h_lr, pl->d_lr, (int64_t)n_utt * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (h_asum) CUDA_TRY(cudaMemcpyAsync(h_asum, pl->asum, frames * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (h_zcr) CUDA_TRY(cudaMemcpyAsync(h_zcr, pl->zcr, frames * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return DSPFE_OK;
}

}  // extern "C"
