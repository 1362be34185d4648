// SIMT portability shim for the dspfe kernels.
//
// The kernel bodies (mfcc_kernel.cuh, endpoint_kernel.cuh, pitch_kernel.cuh) are written once
// against this small vocabulary.  Under nvcc (the product build, sm_100a only) every item maps to
// a CUDA intrinsic or inline PTX.  Under -DDSPFE_EMU (tests/emu only, plain g++) the same bodies
// run on a fibre-based SIMT emulator so that index arithmetic, halo handling and barrier placement
// can be checked in the GPU-less container.  The emulator is test infrastructure: it is never
// linked into libdspfe.so and no product path can reach it.
#pragma once
#include <stdint.h>

#ifdef DSPFE_EMU
// ---------------------------------------------------------------- emulator (tests/emu only)
#include <cmath>
#include <cstring>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#define DEVFN static inline
#define DSPFE_RESTRICT
namespace simt {
int tid();                         // threadIdx.x
int bid();                         // blockIdx.x
int nthreads();                    // blockDim.x
void cta_sync();                   // __syncthreads
void group_sync();                 // barrier over the caller's 16-lane group
float shfl16(float v, int src);    // __shfl_sync(halfmask, v, src, 16)
float shfl32_xor(float v, int m);  // full-warp xor shuffle
int shfl32_i(int v, int src);
unsigned ballot32(bool pred);      // __ballot_sync(full mask, pred)
void warp_sync();                  // __syncwarp()
struct mbar_t { uint64_t v; };
static inline void mbar_init(mbar_t*, int) {}
static inline void fence_mbar_init() {}
static inline void fence_proxy_async() {}
static inline void mbar_expect_tx(mbar_t*, uint32_t) {}
static inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, mbar_t*) { std::memcpy(dst, src, bytes); }
static inline void mbar_wait(mbar_t*, uint32_t) {}
}  // namespace simt
DEVFN float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DEVFN float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
DEVFN float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
DEVFN float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
DEVFN float dsp_logf(float x) { return std::log(x); }
static inline float sqrtf_(float x) { return std::sqrt(x); }
DEVFN float dsp_fmaf(float a, float b, float c) { return std::fmaf(a, b, c); }
DEVFN float dsp_fast_logf(float x) { return std::log(x); }
DEVFN float dsp_fast_sqrtf(float x) { return std::sqrt(x); }
DEVFN float cvt_i16(int v) { return (float)v; }
template <class T> DEVFN T ldg(const T* p) { return *p; }
// int16 halves of a 32-bit word -> float
DEVFN float cvt_lo16(uint32_t w) { return (float)(int16_t)(w & 0xffffu); }
DEVFN float cvt_hi16(uint32_t w) { return (float)(int16_t)(w >> 16); }
struct uint4 { uint32_t x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#else
// ---------------------------------------------------------------- sm_100a device build
#include <cuda_runtime.h>
#define DEVFN __device__ __forceinline__
#define DSPFE_RESTRICT __restrict__
namespace simt {
DEVFN int tid() { return threadIdx.x; }
DEVFN int bid() { return blockIdx.x; }
DEVFN int nthreads() { return blockDim.x; }
DEVFN void cta_sync() { __syncthreads(); }
// the two 16-lane groups of a warp always run the same code path, so full-warp primitives are safe
DEVFN void group_sync() { __syncwarp(); }
DEVFN float shfl16(float v, int src) { return __shfl_sync(0xffffffffu, v, src, 16); }
DEVFN float shfl32_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
DEVFN int shfl32_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
DEVFN unsigned ballot32(bool pred) { return __ballot_sync(0xffffffffu, pred); }
DEVFN void warp_sync() { __syncwarp(); }

// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP): global -> shared::cta.
struct mbar_t { uint64_t v; };
DEVFN uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVFN void mbar_init(mbar_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
DEVFN void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DEVFN void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DEVFN void mbar_expect_tx(mbar_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
DEVFN void bulk_g2s(void* dst, const void* src, uint32_t bytes, mbar_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
DEVFN void mbar_wait(mbar_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
}  // namespace simt
// Packed FP32 (FADD2/FMUL2/FFMA2 on sm_100a): .x and .y carry two independent frames.
DEVFN float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
DEVFN float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
DEVFN float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
DEVFN float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
DEVFN float dsp_logf(float x) { return logf(x); }
DEVFN float dsp_fmaf(float a, float b, float c) { return fmaf(a, b, c); }
// MUFU-based log / sqrt (about 2 ulp): the pitch kernels take 512 logs and up to 512 square roots per frame
DEVFN float dsp_fast_logf(float x) { return __logf(x); }
DEVFN float dsp_fast_sqrtf(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// one int16 value -> float by mantissa splicing (exact; I2F runs at 16/clk/SM)
DEVFN float cvt_i16(int v) { return __uint_as_float(0x4B000000u | (((unsigned)v & 0xffffu) ^ 0x8000u)) - 8421376.0f; }
template <class T> DEVFN T ldg(const T* p) { return __ldg(p); }
// int16 halves of a 32-bit word -> float without the (slow, XU-pipe) I2F: splice the biased 16-bit
// value into the mantissa of 2^23 and subtract 2^23 + 2^15 (exact).
DEVFN float cvt_lo16(uint32_t w) { return __uint_as_float(__byte_perm(w ^ 0x80008000u, 0x4B000000u, 0x7610)) - 8421376.0f; }
DEVFN float cvt_hi16(uint32_t w) { return __uint_as_float(__byte_perm(w ^ 0x80008000u, 0x4B000000u, 0x7632)) - 8421376.0f; }
#endif

#ifdef DSPFE_EMU
DEVFN int __float_as_int_compat(float f) { int i; std::memcpy(&i, &f, 4); return i; }
DEVFN float __int_as_float_compat(int i) { float f; std::memcpy(&f, &i, 4); return f; }
#else
DEVFN int __float_as_int_compat(float f) { return __float_as_int(f); }
DEVFN float __int_as_float_compat(int i) { return __int_as_float(i); }
#endif

// population count
DEVFN int dsp_popc(unsigned v) {
#ifdef DSPFE_EMU
    return __builtin_popcount(v);
#else
    return __popc(v);
#endif
}

// warp-wide integer sum (REDUX on sm_100a)
DEVFN int warp_redux_add(int v) {
#ifdef DSPFE_EMU
    for (int m = 16; m >= 1; m >>= 1) v += simt::shfl32_i(v, (simt::tid() & 31) ^ m);
    return v;
#else
    return __reduce_add_sync(0xffffffffu, v);
#endif
}
DEVFN unsigned warp_redux_min(unsigned v) {
#ifdef DSPFE_EMU
    for (int m = 16; m >= 1; m >>= 1) { const unsigned o = (unsigned)simt::shfl32_i((int)v, (simt::tid() & 31) ^ m); v = o < v ? o : v; }
    return v;
#else
    return __reduce_min_sync(0xffffffffu, v);
#endif
}

DEVFN unsigned warp_redux_max(unsigned v) {
#ifdef DSPFE_EMU
    for (int m = 16; m >= 1; m >>= 1) { const unsigned o = (unsigned)simt::shfl32_i((int)v, (simt::tid() & 31) ^ m); v = o > v ? o : v; }
    return v;
#else
    return __reduce_max_sync(0xffffffffu, v);
#endif
}

// float64 xor-shuffle over the full warp (two 32-bit shuffles)
DEVFN double shfl32_xor_f64(double v, int m) {
    const int src = (simt::tid() & 31) ^ m;
#ifdef DSPFE_EMU
    int w[2]; std::memcpy(w, &v, 8);
    w[0] = simt::shfl32_i(w[0], src); w[1] = simt::shfl32_i(w[1], src);
    double r; std::memcpy(&r, w, 8); return r;
#else
    return __hiloint2double(simt::shfl32_i(__double2hiint(v), src), simt::shfl32_i(__double2loint(v), src));
#endif
}

// float64 from a given lane
DEVFN double shfl32_f64(double v, int src) {
#ifdef DSPFE_EMU
    int w[2]; std::memcpy(w, &v, 8);
    w[0] = simt::shfl32_i(w[0], src); w[1] = simt::shfl32_i(w[1], src);
    double r; std::memcpy(&r, w, 8); return r;
#else
    return __hiloint2double(simt::shfl32_i(__double2hiint(v), src), simt::shfl32_i(__double2loint(v), src));
#endif
}

// pre-emphasis cur - c * prev exactly as NumPy evaluates it in float64 (product rounded, then the difference rounded; no
// contraction), rounded once more to float32: the sign and the exact zeros of the result decide which samples enter the
// median of center_clip, so they must not depend on a float32 evaluation order
DEVFN float dsp_preemph_f64(double c, float cur, float prev) {
#ifdef DSPFE_EMU
    volatile double prod = c * (double)prev;
    volatile double diff = (double)cur - prod;
    return (float)diff;
#else
    return (float)__dsub_rn((double)cur, __dmul_rn(c, (double)prev));
#endif
}

// scalar-broadcast forms (the scalar folds into the packed instruction's .F32 operand)
DEVFN float2 f2muls(float2 a, float s) { return f2mul(a, make_float2(s, s)); }
DEVFN float2 f2fmas(float2 a, float s, float2 c) { return f2fma(a, make_float2(s, s), c); }
DEVFN float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
