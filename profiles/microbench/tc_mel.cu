// tcgen05 experiment for the mel filterbank contraction of K1 (north_star: "the mel and DCT contractions stay on FP32 FFMA unless
// ncu shows a TF32/bf16 tcgen05 tile is both faster and within tolerance").
//
// The contraction: mel[f][j] = sum_k P[f][k] * W[j][k], 26 filters (N = 32), 257 bins (K = 264), a tile of M = 128 frames.
//   kernel tc:   3xTF32 split on the 5th-generation tensor cores: P = P_hi + P_lo, W = W_hi + W_lo (each rounded to TF32),
//                D = P_hi W_hi + P_hi W_lo + P_lo W_hi accumulated in TMEM by tcgen05.mma.kind::tf32 (M128 N32 K8 per instruction,
//                99 instructions per tile), operands in shared memory in the canonical K-major no-swizzle layout (8 x 16-byte core
//                matrices), the power spectrum written there by the threads that produce it (one row = one frame per thread), D read
//                back with tcgen05.ld, log, DCT-II x lifter on FFMA (a thread owns a frame's 26 log-mel values: 338 FFMA per frame).
//   kernel ffma: the same 128-frame tile on FP32 FFMA: the power spectrum goes to shared memory once ([k][frame], conflict-free),
//                a thread owns a frame and walks the SPARSE triangular filters (459 non-zero weights, 2 FFMA-equivalents per bin: every bin
//                belongs to at most two filters), weights from shared memory (broadcast reads), then the same log + DCT.
// Both produce the 13 cepstra per frame from procedurally generated power spectra (same values), so they can be compared for time and
// against a float64 host evaluation for error.  Reported: ms per 1e6 frames, shared-memory bytes written + read per frame (counted
// from the layout), max |err| / (1 + |ref|) of the cepstra.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_mel tc_mel.cu      Run: ./tc_mel [tiles_per_cta]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int M = 128, NF = 26, N = 32, KB = 257, K = 264, KC = K / 8, NCEP = 13;
constexpr int A_CHUNK = 4096, B_CHUNK = 1024;                 // bytes per 8-wide K chunk: 2 core-matrix columns x (rows / 8) x 128 B
constexpr int STAGE_CHUNKS = 11;                              // K chunks staged at a time (3 stages per tile)
constexpr int SM_B = 2 * KC * B_CHUNK;                        // W_hi | W_lo
constexpr int SM_A = 2 * STAGE_CHUNKS * A_CHUNK;              // P_hi | P_lo of one stage
constexpr int SM_TC = SM_B + SM_A + 64;

// procedural power spectrum, exactly reproducible on the host: a 16-bit pseudo-random mantissa times a power-of-two envelope that
// falls by 2^-16 over the 257 bins (a 96 dB range with the mantissa's)
__host__ __device__ inline float gen_power(int frame, int k) {
    uint32_t h = (uint32_t)frame * 1315423911u ^ (uint32_t)k * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float scale = 1.0f / (float)(1u << (k >> 4));
    return (float)((h & 0xffffu) + 1u) * scale;
}
static double gen_power_host(int frame, int k) { return (double)gen_power(frame, k); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }
// K-major, no swizzle: element (row, k) of a chunk lives at (k / 4) * lbo + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                                    // descriptor version of sm_100
    return d;                                                  // base offset 0, layout type 0 = no swizzle
}
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // F32 accum, TF32 x TF32, K-major

struct Tables { float dct[NCEP][NF]; int edge[NF + 2]; };
__constant__ Tables c_tab;

// ------------------------------------------------------------------------------------------------ tensor-core kernel
__global__ void __launch_bounds__(128, 1) tc_kernel(const float* __restrict__ w_dense /*[N][K]*/, int tiles, float* __restrict__ out, int* __restrict__ err_flag) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;                     // [2][KC][B_CHUNK]
    unsigned char* sA = smem + SM_B;              // [2][STAGE_CHUNKS][A_CHUNK]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + SM_B + SM_A);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_B + SM_A + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    // W_hi / W_lo in the B-operand layout (N rows x K)
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i - n * K;
        const float w = w_dense[i], hi = to_tf32(w), lo = to_tf32(w - hi);
        const int off = (k >> 3) * B_CHUNK + ((k >> 2) & 1) * (N / 8 * 128) + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4;
        *reinterpret_cast<float*>(sB + off) = hi;
        *reinterpret_cast<float*>(sB + KC * B_CHUNK + off) = lo;
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    uint32_t parity = 0;
    bool dead = false;
    for (int tile = 0; tile < tiles && !dead; ++tile) {
        const int frame = (blockIdx.x * tiles + tile) * M + tid;       // this thread's row
        for (int st = 0; st < KC / STAGE_CHUNKS; ++st) {
            // ---- the thread writes its frame's power values of this stage as P_hi / P_lo (two 16-byte vectors per 8 bins and half)
#pragma unroll 1
            for (int c = 0; c < STAGE_CHUNKS; ++c) {
                const int k0 = (st * STAGE_CHUNKS + c) * 8;
                float hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float p = k0 + e < KB ? gen_power(frame, k0 + e) : 0.f;
                    hi[e] = to_tf32(p); lo[e] = to_tf32(p - hi[e]);
                }
                unsigned char* a = sA + c * A_CHUNK + (tid >> 3) * 128 + (tid & 7) * 16;
                *reinterpret_cast<float4*>(a) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(a + 2048) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                *reinterpret_cast<float4*>(a + STAGE_CHUNKS * A_CHUNK) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(a + STAGE_CHUNKS * A_CHUNK + 2048) = make_float4(lo[4], lo[5], lo[6], lo[7]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core's async proxy
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c = 0; c < STAGE_CHUNKS; ++c) {
                    const int kc = st * STAGE_CHUNKS + c;
                    const uint64_t a_hi = make_desc(smem_u32(sA + c * A_CHUNK), 2048, 128);
                    const uint64_t a_lo = make_desc(smem_u32(sA + (STAGE_CHUNKS + c) * A_CHUNK), 2048, 128);
                    const uint64_t b_hi = make_desc(smem_u32(sB + kc * B_CHUNK), N / 8 * 128, 128);
                    const uint64_t b_lo = make_desc(smem_u32(sB + (KC + kc) * B_CHUNK), N / 8 * 128, 128);
                    const uint64_t da[3] = {a_hi, a_hi, a_lo}, db[3] = {b_hi, b_lo, b_hi};
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const uint32_t acc = (kc > 0 || t > 0) ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                                     ::"r"(tmem), "l"(da[t]), "l"(db[t]), "r"(kIdesc), "r"(acc), "r"(0u) : "memory");
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
            }
            // ---- everybody waits for the stage's MMAs (they read sA): bounded spin, a stuck barrier ends the kernel instead of hanging it
            uint32_t done = 0;
            for (int spin = 0; spin < (1 << 22) && !done; ++spin)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
            if (!done) { if (tid == 0) atomicExch(err_flag, 1); dead = true; }
            dead = __syncthreads_or(dead ? 1 : 0) != 0;
            if (dead) break;
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (dead) break;
        // ---- D (128 lanes x 32 columns fp32) -> registers: warp w reads lanes 32 w .. 32 w + 31, one row per thread
        uint32_t d[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
                     "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]),
                       "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]),
                       "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        float lm[NF];
#pragma unroll
        for (int j = 0; j < NF; ++j) lm[j] = __logf(fmaxf(__uint_as_float(d[j]), 2.220446049250313e-16f));
#pragma unroll
        for (int c = 0; c < NCEP; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < NF; ++j) acc = fmaf(c_tab.dct[c][j], lm[j], acc);
            out[(size_t)frame * NCEP + c] = acc;
        }
        __syncthreads();                                               // TMEM is read out before the next tile's first MMA overwrites it
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------ FFMA kernel (sparse filters)
constexpr int SM_FF = K * M * 4 + 2 * K * 4;
__global__ void __launch_bounds__(128, 1) ffma_kernel(const float* __restrict__ w_up /*[K]*/, const float* __restrict__ w_dn /*[K]*/, int tiles, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* sP = reinterpret_cast<float*>(smem);                 // [K][M]: bin-major, a thread's frame in its own column (conflict-free)
    float* sUp = sP + K * M;                                    // weight of bin k in the filter that rises through it, and in the one that falls
    float* sDn = sUp + K;
    const int tid = threadIdx.x;
    for (int i = tid; i < K; i += 128) { sUp[i] = w_up[i]; sDn[i] = w_dn[i]; }
    __syncthreads();
    for (int tile = 0; tile < tiles; ++tile) {
        const int frame = (blockIdx.x * tiles + tile) * M + tid;
#pragma unroll 4
        for (int k = 0; k < KB; ++k) sP[k * M + tid] = gen_power(frame, k);      // the spectrum goes through shared memory once, as in K1
        float lm[NF];
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            // filter j rises over [e_j, e_j+1) and falls over [e_j+1, e_j+2): bin k's rising weight is sUp[k], its falling weight sDn[k]
            float acc = 0.f;
            for (int k = c_tab.edge[j]; k < c_tab.edge[j + 1]; ++k) acc = fmaf(sUp[k], sP[k * M + tid], acc);
            for (int k = c_tab.edge[j + 1]; k < c_tab.edge[j + 2]; ++k) acc = fmaf(sDn[k], sP[k * M + tid], acc);
            lm[j] = __logf(fmaxf(acc, 2.220446049250313e-16f));
        }
#pragma unroll
        for (int c = 0; c < NCEP; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < NF; ++j) acc = fmaf(c_tab.dct[c][j], lm[j], acc);
            out[(size_t)frame * NCEP + c] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------ common part only: power values, log, DCT, stores
// (the 26 "mel" sums are plain register sums of every 26th bin: no shared memory, no contraction) -- subtract it from both kernels
__global__ void __launch_bounds__(128, 1) base_kernel(int tiles, float* __restrict__ out) {
    const int tid = threadIdx.x;
    for (int tile = 0; tile < tiles; ++tile) {
        const int frame = (blockIdx.x * tiles + tile) * M + tid;
        float lm[NF];
#pragma unroll
        for (int j = 0; j < NF; ++j) lm[j] = 0.f;
#pragma unroll 1
        for (int k0 = 0; k0 < 260; k0 += NF) {
#pragma unroll
            for (int j = 0; j < NF; ++j) lm[j] += k0 + j < KB ? gen_power(frame, k0 + j) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NF; ++j) lm[j] = __logf(fmaxf(lm[j], 2.220446049250313e-16f));
#pragma unroll
        for (int c = 0; c < NCEP; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < NF; ++j) acc = fmaf(c_tab.dct[c][j], lm[j], acc);
            out[(size_t)frame * NCEP + c] = acc;
        }
    }
}

int main(int argc, char** argv) {
    const int tiles = argc > 1 ? atoi(argv[1]) : 64;
    int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int ctas = prop.multiProcessorCount;
    const long long frames = (long long)ctas * tiles * M;
    // mel filterbank of base.py:40-58 (26 filters, nfft 512, 16 kHz) and the ortho DCT-II x lifter of base.py:12-14
    Tables tab; std::vector<float> wd(N * K, 0.f), up(K, 0.f), dn(K, 0.f);
    {
        auto hz2mel = [](double hz) { return 2595.0 * log10(1 + hz / 700.0); };
        auto mel2hz = [](double mel) { return 700.0 * (pow(10.0, mel / 2595.0) - 1); };
        double bin[NF + 2];
        for (int i = 0; i < NF + 2; ++i) { const double mel = hz2mel(0) + (hz2mel(8000) - hz2mel(0)) * i / (NF + 1); bin[i] = floor(513.0 * mel2hz(mel) / 16000.0); tab.edge[i] = (int)bin[i]; }
        for (int j = 0; j < NF; ++j) {
            for (int i = (int)bin[j]; i < (int)bin[j + 1]; ++i) { wd[j * K + i] = (float)((i - bin[j]) / (bin[j + 1] - bin[j])); up[i] = wd[j * K + i]; }
            for (int i = (int)bin[j + 1]; i < (int)bin[j + 2]; ++i) { wd[j * K + i] = (float)((bin[j + 2] - i) / (bin[j + 2] - bin[j + 1])); dn[i] = wd[j * K + i]; }
        }
        for (int c = 0; c < NCEP; ++c)
            for (int j = 0; j < NF; ++j)
                tab.dct[c][j] = (float)((c == 0 ? sqrt(1.0 / NF) : sqrt(2.0 / NF)) * cos(M_PI * c * (2 * j + 1) / (2.0 * NF)) * (1 + 11.0 * sin(M_PI * c / 22.0)));
    }
    CK(cudaMemcpyToSymbol(c_tab, &tab, sizeof(tab)));
    float *d_w, *d_up, *d_dn, *d_out_tc, *d_out_ff; int* d_err;
    CK(cudaMalloc(&d_w, wd.size() * 4)); CK(cudaMalloc(&d_up, K * 4)); CK(cudaMalloc(&d_dn, K * 4));
    CK(cudaMalloc(&d_out_tc, frames * NCEP * 4)); CK(cudaMalloc(&d_out_ff, frames * NCEP * 4)); CK(cudaMalloc(&d_err, 4));
    CK(cudaMemcpy(d_w, wd.data(), wd.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_up, up.data(), K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_dn, dn.data(), K * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(d_err, 0, 4));
    CK(cudaMemset(d_out_tc, 0, frames * NCEP * 4));
    CK(cudaFuncSetAttribute(tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TC));
    CK(cudaFuncSetAttribute(ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_FF));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms_tc = 0, ms_ff = 0, ms_base = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0)); base_kernel<<<ctas, 128>>>(tiles, d_out_ff); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms_base, e0, e1));
        CK(cudaEventRecord(e0)); tc_kernel<<<ctas, 128, SM_TC>>>(d_w, tiles, d_out_tc, d_err); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms_tc, e0, e1));
        CK(cudaEventRecord(e0)); ffma_kernel<<<ctas, 128, SM_FF>>>(d_up, d_dn, tiles, d_out_ff); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms_ff, e0, e1));
    }
    int h_err = 0; CK(cudaMemcpy(&h_err, d_err, 4, cudaMemcpyDeviceToHost));
    const int check = 4 * M;
    std::vector<float> tc(check * NCEP), ff(check * NCEP);
    CK(cudaMemcpy(tc.data(), d_out_tc, tc.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ff.data(), d_out_ff, ff.size() * 4, cudaMemcpyDeviceToHost));
    double err_tc = 0, err_ff = 0;
    for (int f = 0; f < check; ++f) {
        double lm[NF];
        for (int j = 0; j < NF; ++j) { double a = 0; for (int k = 0; k < KB; ++k) a += (double)wd[j * K + k] * gen_power_host(f, k); lm[j] = log(a > 2.220446049250313e-16 ? a : 2.220446049250313e-16); }
        for (int c = 0; c < NCEP; ++c) {
            double r = 0; for (int j = 0; j < NF; ++j) r += (double)tab.dct[c][j] * lm[j];
            err_tc = fmax(err_tc, fabs(tc[f * NCEP + c] - r) / (1 + fabs(r))); err_ff = fmax(err_ff, fabs(ff[f * NCEP + c] - r) / (1 + fabs(r)));
        }
    }
    // shared-memory bytes per frame, from the layouts: tc writes P_hi + P_lo (2 x 264 x 4) and the tensor core reads P_hi twice, P_lo once per K chunk
    // (3 x 264 x 4) plus the weights (3 x 33 x 1 KB per 128 frames); ffma writes 257 x 4 and reads every bin once per filter it belongs to (<= 2 x 257 x 4)
    printf("{\"common_part_ns_per_frame\": %.3f, \"tc_net_ns_per_frame\": %.3f, \"ffma_net_ns_per_frame\": %.3f}\n", ms_base * 1e6 / frames,
           (ms_tc - ms_base) * 1e6 / frames, (ms_ff - ms_base) * 1e6 / frames);
    printf("{\"frames\": %lld, \"tc_3xtf32\": {\"ms\": %.4f, \"ns_per_frame\": %.3f, \"max_err\": %.3e, \"smem_bytes_per_frame\": %d, \"barrier_timeout\": %d}, "
           "\"ffma_sparse\": {\"ms\": %.4f, \"ns_per_frame\": %.3f, \"max_err\": %.3e, \"smem_bytes_per_frame\": %d}, \"note\": \"errors against a float64 evaluation of the same float32 power values\"}\n",
           frames, ms_tc, ms_tc * 1e6 / frames, err_tc, 2 * K * 4 + 3 * K * 4 + 3 * KC * B_CHUNK / M, h_err,
           ms_ff, ms_ff * 1e6 / frames, err_ff, KB * 4 + 459 * 4);
    return 0;
}
