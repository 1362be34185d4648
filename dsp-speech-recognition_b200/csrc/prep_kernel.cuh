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

// frame count of a (trimmed) utterance: 32-bit arithmetic (seg_len is an int32), same rule as num_frames
__device__ __forceinline__ int num_frames32(int len, int frame_len, int frame_step) {
    if (len <= frame_len) return 1;
    return 1 + (int)(((unsigned)(len - frame_len) + (unsigned)frame_step - 1u) / (unsigned)frame_step);
}

// One CTA: every thread takes `per` consecutive utterances, the per-thread totals are scanned with warp shuffles and one
// shared-memory step (two barriers; this kernel sits in front of every MFCC launch, so its latency is fully exposed).
__global__ void __launch_bounds__(kPrepThreads) prep_kernel(PrepParams p) {
    __shared__ long long s_fr[32];
    __shared__ int s_ti[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (p.n_utt + kPrepThreads - 1) / kPrepThreads;
    const int u0 = min(tid * per, p.n_utt), u1 = min(u0 + per, p.n_utt);
    constexpr int kKeep = 4;             // frame counts kept in registers between the two passes
    int Fk[kKeep];

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
        const int F = num_frames32((int)len, p.frame_len, p.frame_step);
#pragma unroll
        for (int k = 0; k < kKeep; ++k) if (u - u0 == k) Fk[k] = F;
        fr += F; ti += tiles_of(F, p.seg_frames);
    }
    // inclusive scan of the per-thread totals: inside the warp, then over the 32 warp totals
    long long sf = fr; int st = ti;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long f = __shfl_up_sync(0xffffffffu, sf, d); const int t = __shfl_up_sync(0xffffffffu, st, d);
        if (lane >= d) { sf += f; st += t; }
    }
    if (lane == 31) { s_fr[warp] = sf; s_ti[warp] = st; }
    __syncthreads();
    if (warp == 0) {
        long long wf = s_fr[lane]; int wt = s_ti[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long f = __shfl_up_sync(0xffffffffu, wf, d); const int t = __shfl_up_sync(0xffffffffu, wt, d);
            if (lane >= d) { wf += f; wt += t; }
        }
        s_fr[lane] = wf; s_ti[lane] = wt;
    }
    __syncthreads();
    if (warp > 0) { sf += s_fr[warp - 1]; st += s_ti[warp - 1]; }
    long long fo = sf - fr; int to = st - ti;
    for (int u = u0; u < u1; ++u) {
        int F = 0;
        if (u - u0 < kKeep) {
#pragma unroll
            for (int k = 0; k < kKeep; ++k) if (u - u0 == k) F = Fk[k];
        } else {
            F = num_frames32(p.seg_len[u], p.frame_len, p.frame_step);
        }
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
        p.frame_off[p.n_utt] = sf;
        p.tile_off[p.n_utt] = st;
        *p.ntiles = min(st, p.max_tiles);
    }
}

}  // namespace dspfe
