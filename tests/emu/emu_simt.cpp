// Fibre-based SIMT emulator for the dspfe kernel bodies (TEST INFRASTRUCTURE ONLY).
//
// Compiles the very same kernel bodies as libdspfe.so (csrc/*.cuh) with -DDSPFE_EMU and runs every
// CUDA thread of a CTA as a ucontext fibre on one host thread; __syncthreads / group barriers /
// shuffles become fibre yields.  It exists so that the GPU-less container can check index
// arithmetic, halo handling and barrier placement of the kernels against the oracle.  It is built
// into tests/emu/_build/libdspfe_emu.so, which only tests/ load; the product library never links it.
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../dsp-speech-recognition_b200/csrc/simt.h"

namespace emu {
struct Fiber { ucontext_t ctx; std::vector<char> stack; bool done = false; };
static std::vector<Fiber> g_fibers;
static ucontext_t g_sched;
static int g_cur = 0, g_bid = 0, g_nthreads = 0;
static int g_cta_count = 0, g_cta_gen = 0;
static std::vector<int> g_grp_count, g_grp_gen, g_warp_count, g_warp_gen;
static std::vector<float> g_slot;
static void (*g_body)(void*) = nullptr;
static void* g_arg = nullptr;

static void yield() { swapcontext(&g_fibers[g_cur].ctx, &g_sched); }
static void barrier(int& count, int& gen, int need) {
    const int my = gen;
    if (++count == need) { count = 0; ++gen; return; }
    while (gen == my) yield();
}
static void entry() { g_body(g_arg); g_fibers[g_cur].done = true; swapcontext(&g_fibers[g_cur].ctx, &g_sched); }

// Runs one CTA; returns false if the fibres deadlock (a barrier not reached by every thread).
bool run_cta(int bid, int nthreads, void (*body)(void*), void* arg) {
    g_bid = bid; g_nthreads = nthreads; g_body = body; g_arg = arg;
    g_cta_count = 0; g_cta_gen = 0;
    g_grp_count.assign(nthreads / 16 + 1, 0); g_grp_gen.assign(nthreads / 16 + 1, 0);
    g_warp_count.assign(nthreads / 32 + 1, 0); g_warp_gen.assign(nthreads / 32 + 1, 0);
    g_slot.assign(nthreads, 0.f);
    g_fibers.clear(); g_fibers.resize(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        Fiber& f = g_fibers[t];
        f.stack.resize(256 * 1024);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data(); f.ctx.uc_stack.ss_size = f.stack.size(); f.ctx.uc_link = &g_sched;
        makecontext(&f.ctx, (void (*)())entry, 0);
    }
    int live = nthreads, idle_rounds = 0;
    // exited threads no longer count towards __syncthreads (as on the device for early-exit CTAs)
    while (live > 0) {
        int before_gen = g_cta_gen; int progressed = 0;
        for (int t = 0; t < nthreads; ++t) {
            if (g_fibers[t].done) continue;
            g_cur = t;
            swapcontext(&g_sched, &g_fibers[t].ctx);
            if (g_fibers[t].done) { --live; ++progressed; }
        }
        (void)before_gen;
        if (++idle_rounds > 2000000) return false;
        (void)progressed;
    }
    return true;
}
}  // namespace emu

namespace simt {
int tid() { return emu::g_cur; }
int bid() { return emu::g_bid; }
int nthreads() { return emu::g_nthreads; }
void cta_sync() { emu::barrier(emu::g_cta_count, emu::g_cta_gen, emu::g_nthreads); }
void group_sync() { int g = emu::g_cur / 16; emu::barrier(emu::g_grp_count[g], emu::g_grp_gen[g], 16); }
float shfl16(float v, int src) {
    const int me = emu::g_cur;
    emu::g_slot[me] = v;
    group_sync();
    const float r = emu::g_slot[(me & ~15) + (src & 15)];
    group_sync();
    return r;
}
void warp_sync() { int w = emu::g_cur / 32; emu::barrier(emu::g_warp_count[w], emu::g_warp_gen[w], 32); }
float shfl32_xor(float v, int m) {
    const int me = emu::g_cur;
    emu::g_slot[me] = v;
    warp_sync();
    const float r = emu::g_slot[(me & ~31) + ((me ^ m) & 31)];
    warp_sync();
    return r;
}
int shfl32_i(int v, int src) {
    float f; std::memcpy(&f, &v, 4);
    const int me = emu::g_cur;
    emu::g_slot[me] = f;
    warp_sync();
    const float r = emu::g_slot[(me & ~31) + (src & 31)];
    warp_sync();
    int o; std::memcpy(&o, &r, 4);
    return o;
}
unsigned ballot32(bool pred) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)(shfl32_i(pred ? 1 : 0, l) != 0) << l;
    return m;
}
}  // namespace simt
