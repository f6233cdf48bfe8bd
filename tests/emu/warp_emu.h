// warp_emu.h -- TEST INFRASTRUCTURE: a 32-fiber lock-step emulator of one CUDA warp, so that
// crispresso2_b200/csrc/c2b_core.cuh (the real kernel logic) can be compiled with g++ and checked against the
// oracle on a box without a GPU.  Never shipped, never loaded by the product package.
//
// Each lane is a ucontext fiber; a warp collective writes the lane's value into a double-buffered slot array,
// yields round-robin, and reads the peers' values when control returns.  A tag per collective asserts that all
// lanes reached the same kind of collective (a divergence there would be undefined behaviour on the GPU).
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <execinfo.h>
#include <ucontext.h>

#define C2B_DEV static inline
#define C2B_DEVNOINL static
struct int4 { int x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t a, uint32_t b) { uint2 r = {a, b}; return r; }
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint4 r = {a, b, c, d}; return r; }

namespace emu {
struct Warp {
    ucontext_t main_ctx, ctx[32];
    char *stacks[32];
    bool done[32];
    int cur = 0;
    uint64_t slot[2][32]; int tag[2][32]; unsigned cnt[32];
    std::function<void()> body;
};
extern thread_local Warp *g_warp;
static void fiber_entry() { g_warp->body(); g_warp->done[g_warp->cur] = true; swapcontext(&g_warp->ctx[g_warp->cur], &g_warp->main_ctx); }
inline void yield() { Warp *w = g_warp; swapcontext(&w->ctx[w->cur], &w->main_ctx); }
inline void run_warp(const std::function<void()> &fn)
{
    static thread_local Warp *W = nullptr;
    if (!W) { W = new Warp(); for (int l = 0; l < 32; l++) W->stacks[l] = (char *)malloc(1 << 19); }
    g_warp = W; W->body = fn;
    for (int l = 0; l < 32; l++) {
        getcontext(&W->ctx[l]); W->ctx[l].uc_stack.ss_sp = W->stacks[l]; W->ctx[l].uc_stack.ss_size = 1 << 19;
        W->ctx[l].uc_link = &W->main_ctx; makecontext(&W->ctx[l], fiber_entry, 0);
        W->done[l] = false; W->cnt[l] = 0;
    }
    bool any = true;
    while (any) {
        any = false;
        for (int l = 0; l < 32; l++) if (!W->done[l]) { W->cur = l; swapcontext(&W->main_ctx, &W->ctx[l]); any = true; }
    }
}
inline uint64_t exchange(uint64_t v, int kind, int src_lane_fn(int, int), int arg)
{
    Warp *w = g_warp; const int l = w->cur; const int b = w->cnt[l] & 1;
    w->slot[b][l] = v; w->tag[b][l] = kind; w->cnt[l]++;
    yield();
    for (int q = 0; q < 32; q++) {
        if (w->cnt[q] < w->cnt[l]) { fprintf(stderr, "warp_emu: lane %d did not reach a collective (kind %d) that lane %d executes\n", q, kind, l); abort(); }
        if (w->tag[b][q] != kind) {
            fprintf(stderr, "warp_emu: divergent collective (lane %d kind %d vs lane %d kind %d)\n", q, w->tag[b][q], l, kind);
            void *bt[24]; const int nbt = backtrace(bt, 24); backtrace_symbols_fd(bt, nbt, 2);     // call path of lane l (build with -O0 -g -rdynamic)
            abort();
        }
    }
    const int s = src_lane_fn(l, arg);
    return w->slot[b][s];
}
}  // namespace emu

namespace wp {
C2B_DEV int lane() { return emu::g_warp->cur; }
C2B_DEV int shfl_up(int v, int d) { return (int)(uint32_t)emu::exchange((uint32_t)v, 1, [](int l, int a) { return l >= a ? l - a : l; }, d); }
C2B_DEV int shfl(int v, int src) { return (int)(uint32_t)emu::exchange((uint32_t)v, 2, [](int, int a) { return a & 31; }, src); }
C2B_DEV uint32_t shflu(uint32_t v, int src) { return (uint32_t)emu::exchange(v, 2, [](int, int a) { return a & 31; }, src); }
C2B_DEV uint32_t shflu_up(uint32_t v, int d) { return (uint32_t)emu::exchange(v, 1, [](int l, int a) { return l >= a ? l - a : l; }, d); }
C2B_DEV int shfl_xor(int v, int m) { return (int)(uint32_t)emu::exchange((uint32_t)v, 3, [](int l, int a) { return (l ^ a) & 31; }, m); }
C2B_DEV uint32_t ballot(bool p)
{
    emu::Warp *w = emu::g_warp; const int l = w->cur; const int b = w->cnt[l] & 1;
    emu::exchange(p ? 1 : 0, 4, [](int l2, int) { return l2; }, 0);
    uint32_t m = 0; for (int q = 0; q < 32; q++) if (w->slot[b][q]) m |= 1u << q;
    return m;
}
C2B_DEV void sync() { emu::exchange(0, 5, [](int l, int) { return l; }, 0); }
static thread_local long g_grp_syncs = 0;       // phase barriers executed by lane 0 (the engine checks the count per work group)
C2B_DEV void grp_sync(int) { if (emu::g_warp->cur == 0) g_grp_syncs++; }      // the emulator runs one warp at a time
C2B_DEV int max3(int a, int b, int c) { return std::max(a, std::max(b, c)); }
C2B_DEV int addmax(int a, int b, int c) { return std::max(a + b, c); }
static inline int16_t h_lo(uint32_t v) { return (int16_t)(v & 0xffffu); }
static inline int16_t h_hi(uint32_t v) { return (int16_t)(v >> 16); }
static inline uint32_t h_pack(int lo, int hi) { return ((uint32_t)(uint16_t)(int16_t)lo) | (((uint32_t)(uint16_t)(int16_t)hi) << 16); }
C2B_DEV uint32_t max3_2(uint32_t a, uint32_t b, uint32_t c)
{ return h_pack(std::max<int>(h_lo(a), std::max<int>(h_lo(b), h_lo(c))), std::max<int>(h_hi(a), std::max<int>(h_hi(b), h_hi(c)))); }
C2B_DEV uint32_t addmax_2(uint32_t a, uint32_t b, uint32_t c)
{ return h_pack(std::max<int>((int16_t)(h_lo(a) + h_lo(b)), h_lo(c)), std::max<int>((int16_t)(h_hi(a) + h_hi(b)), h_hi(c))); }
C2B_DEV uint4 ldg4u(const uint4 *p) { return *p; }
C2B_DEV void prefetch_l2(const void *) {}
// emulator: "shared addresses" are offsets into a per-thread pointer table
static thread_local const unsigned char *g_smem_base = nullptr;
C2B_DEV uint32_t smem_addr(const void *p) { if (!g_smem_base) g_smem_base = (const unsigned char *)p - 4096; return (uint32_t)((const unsigned char *)p - g_smem_base); }
C2B_DEV uint32_t lds_u8(uint32_t a) { return g_smem_base[a]; }
C2B_DEV uint4 lds_v4(uint32_t a) { return *reinterpret_cast<const uint4 *>(g_smem_base + a); }
C2B_DEV uint2 ldcg2(const uint2 *p) { return *p; }
C2B_DEV uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
C2B_DEV int popc(uint32_t x) { return __builtin_popcount(x); }
C2B_DEV int popcll(uint64_t x) { return __builtin_popcountll(x); }
C2B_DEV int clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
C2B_DEV int ffs(uint32_t x) { return __builtin_ffs((int)x); }
C2B_DEV uint32_t ldcg(const uint32_t *p) { return *p; }
C2B_DEV int ldcgi(const int *p) { return *p; }
C2B_DEV uint64_t ldcg64(const uint64_t *p) { return *p; }
C2B_DEV int4 ldg4(const int4 *p) { return *p; }
C2B_DEV void addg(unsigned long long *p, long long v) { *p += (unsigned long long)v; }
C2B_DEV void maxg(unsigned long long *p, unsigned long long v) { if (v > *p) *p = v; }
C2B_DEV uint32_t adds(uint32_t *p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
C2B_DEV unsigned long long fetch_work(unsigned long long *p) { return (*p)++; }
C2B_DEV unsigned long long fetch_add(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
}  // namespace wp
