// K0: batch preparation on the device (no host round trip).
// From the packed utterance boundaries (and optionally the endpoint kernel's (left,right) pairs) it
// derives each utterance's sample range, its frame count (reference sigproc.py:79-82), the exclusive
// prefix sums that place the utterance in the output, and the tile table the MFCC kernel walks.
#pragma once
#include <cuda_runtime.h>

#include "dspfe_types.h"

namespace dspfe {

constexpr int kPrepThreads = 1024;

// tiles of one utterance: T = ceil(F/seg) tiles of ceil(F/T) frames (the last may be shorter)
__host__ __device__ inline int tiles_of(int F, int seg) { return (F + seg - 1) / seg; }

__global__ void __launch_bounds__(kPrepThreads) prep_kernel(PrepParams p) {
    __shared__ long long s_fr[kPrepThreads];
    __shared__ int s_ti[kPrepThreads];
    const int tid = threadIdx.x;
    const int per = (p.n_utt + kPrepThreads - 1) / kPrepThreads;
    const int u0 = min(tid * per, p.n_utt), u1 = min(u0 + per, p.n_utt);

    long long fr = 0; int ti = 0;
    for (int u = u0; u < u1; ++u) {
        long long a = p.offsets[u], b = p.offsets[u + 1];
        long long len = b - a;
        if (p.trim) {  // Python slice sig[l:r] with l, r >= 0
            long long l = p.trim[2 * u], r = p.trim[2 * u + 1];
            if (l < 0) l = 0; if (r < 0) r = 0;
            if (l > len) l = len; if (r > len) r = len;
            a += l; len = r > l ? r - l : 0;
        }
        p.seg_start[u] = a;
        p.seg_len[u] = (int)len;
        const int F = (int)num_frames(len, p.frame_len, p.frame_step);
        fr += F; ti += tiles_of(F, p.seg_frames);
    }
    s_fr[tid] = fr; s_ti[tid] = ti;
    __syncthreads();
    // inclusive scan of the per-thread totals
    for (int d = 1; d < kPrepThreads; d <<= 1) {
        long long f = tid >= d ? s_fr[tid - d] : 0; int t = tid >= d ? s_ti[tid - d] : 0;
        __syncthreads();
        s_fr[tid] += f; s_ti[tid] += t;
        __syncthreads();
    }
    long long fo = s_fr[tid] - fr; int to = s_ti[tid] - ti;
    for (int u = u0; u < u1; ++u) {
        const int F = (int)num_frames(p.seg_len[u], p.frame_len, p.frame_step);
        const int T = tiles_of(F, p.seg_frames);
        const int per_tile = (F + T - 1) / T;
        p.frame_off[u] = fo; p.tile_off[u] = to;
        for (int t = 0; t < T; ++t) {
            if (to + t < p.max_tiles) {
                Tile tl; tl.utt = u; tl.f0 = t * per_tile; tl.nf = min(per_tile, F - tl.f0); tl.pad = 0;
                p.tiles[to + t] = tl;
            }
        }
        fo += F; to += T;
    }
    if (tid == kPrepThreads - 1) {
        p.frame_off[p.n_utt] = s_fr[tid];
        p.tile_off[p.n_utt] = s_ti[tid];
        *p.ntiles = min(s_ti[tid], p.max_tiles);
    }
}

}  // namespace dspfe
