// c2b_engine.cu -- the sm_100a kernel entry + the C ABI declared in include/c2b200.h.
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC \
//        -Iinclude -o crispresso2_b200/libc2b200.so crispresso2_b200/csrc/c2b_engine.cu
// The same file compiles with g++ -DC2B_EMU against tests/emu/warp_emu.h (CPU-only logic tests).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "c2b_core.cuh"

using namespace c2b;

// ----------------------------------------------------------------------------------- runtime shim
#ifndef C2B_EMU
#define RT_OK cudaSuccess
typedef cudaStream_t rt_stream;
static const char *rt_errstr(cudaError_t e) { return cudaGetErrorString(e); }
typedef cudaError_t rt_err;
static rt_err rt_malloc(void **p, size_t n) { return cudaMalloc(p, n ? n : 16); }
static rt_err rt_free(void *p) { return cudaFree(p); }
static rt_err rt_h2d(void *d, const void *h, size_t n, rt_stream s) { return n ? cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s) : cudaSuccess; }
static rt_err rt_d2h(void *h, const void *d, size_t n, rt_stream s) { return n ? cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s) : cudaSuccess; }
static rt_err rt_zero(void *d, size_t n, rt_stream s) { return cudaMemsetAsync(d, 0, n, s); }
static rt_err rt_d2h_2d(void *h, const void *d, size_t pitch, size_t width, size_t rows, rt_stream s)
{ return (width && rows) ? cudaMemcpy2DAsync(h, pitch, d, pitch, width, rows, cudaMemcpyDeviceToHost, s) : cudaSuccess; }
static rt_err rt_sync(rt_stream s) { return cudaStreamSynchronize(s); }
static rt_err cudaMemcpyAsyncOrCopy(void *d, const void *s_, size_t n, rt_stream st) { return cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, st); }
typedef cudaEvent_t rt_event;
static rt_err rt_event_create(rt_event *e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); }
static rt_err rt_event_destroy(rt_event e) { return cudaEventDestroy(e); }
static rt_err rt_record(rt_event e, rt_stream s) { return cudaEventRecord(e, s); }
static rt_err rt_wait(rt_stream s, rt_event e) { return cudaStreamWaitEvent(s, e, 0); }
static rt_err rt_event_sync(rt_event e) { return cudaEventSynchronize(e); }
static rt_err rt_stream_create(rt_stream *s) { return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking); }
static rt_err rt_stream_destroy(rt_stream s) { return cudaStreamDestroy(s); }
static void *rt_host_alloc(size_t n) { void *p = nullptr; return cudaHostAlloc(&p, n ? n : 16, cudaHostAllocDefault) == cudaSuccess ? p : nullptr; }
static void rt_host_free(void *p) { if (p) cudaFreeHost(p); }
#else
typedef int rt_err;
typedef int rt_stream;
#define RT_OK 0
thread_local emu::Warp *emu::g_warp = nullptr;
static const char *rt_errstr(int) { return "emu"; }
static rt_err rt_malloc(void **p, size_t n) { *p = calloc(1, n ? n : 16); return *p ? 0 : 1; }
static rt_err rt_free(void *p) { free(p); return 0; }
static rt_err rt_h2d(void *d, const void *h, size_t n, rt_stream) { if (n) memcpy(d, h, n); return 0; }
static rt_err rt_d2h(void *h, const void *d, size_t n, rt_stream) { if (n) memcpy(h, d, n); return 0; }
static rt_err rt_zero(void *d, size_t n, rt_stream) { memset(d, 0, n); return 0; }
static rt_err rt_d2h_2d(void *h, const void *d, size_t pitch, size_t width, size_t rows, rt_stream)
{ for (size_t r = 0; r < rows; r++) memcpy((char *)h + r * pitch, (const char *)d + r * pitch, width); return 0; }
static rt_err rt_sync(rt_stream) { return 0; }
static rt_err cudaMemcpyAsyncOrCopy(void *d, const void *s_, size_t n, rt_stream) { memcpy(d, s_, n); return 0; }
typedef int rt_event;
static rt_err rt_event_create(rt_event *e) { *e = 0; return 0; }
static rt_err rt_event_destroy(rt_event) { return 0; }
static rt_err rt_record(rt_event, rt_stream) { return 0; }
static rt_err rt_wait(rt_stream, rt_event) { return 0; }
static rt_err rt_event_sync(rt_event) { return 0; }
static rt_err rt_stream_create(rt_stream *s) { *s = 0; return 0; }
static rt_err rt_stream_destroy(rt_stream) { return 0; }
static void *rt_host_alloc(size_t n) { return malloc(n ? n : 16); }
static void rt_host_free(void *p) { free(p); }
#endif

// ----------------------------------------------------------------------------------------- kernel
#ifndef C2B_EMU
#ifndef C2B_WARPS_PER_CTA
#define C2B_WARPS_PER_CTA 8
#endif
#ifndef C2B_MIN_CTAS_PER_SM
#define C2B_MIN_CTAS_PER_SM 2
#endif
constexpr int WARPS_PER_CTA = C2B_WARPS_PER_CTA;      // launch-bounds maximum; the launch may use fewer (env C2B_WARPS_PER_CTA)

// ONE: every read has one candidate reference (a single amplicon configured, or Pooled ref_id): the lean instantiation
// carries none of the several-references code.  The host picks the instantiation per launch.
// P is a __grid_constant__: the out-of-line device functions take it by reference, and without the qualifier every launch
// copied the 330-byte struct to each thread's local memory and read its fields back with LDL (r01k: 27.3 -> 25.7 ms).
// STREAM: the batch's read bytes arrive while the kernel runs (c2b_align_batch, streamed launch); the resident-batch
// instantiations keep the plain work loop, without the availability wait and the per-group completion signalling.
template <bool ONE, bool STREAM>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, C2B_MIN_CTAS_PER_SM) c2b_align_classify_kernel(const __grid_constant__ KParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpSmem *S = reinterpret_cast<WarpSmem *>(smem_raw) + (threadIdx.x >> 5);
    QuadSmem *Q = reinterpret_cast<QuadSmem *>(smem_raw + (size_t)(blockDim.x >> 5) * sizeof(WarpSmem)) + (threadIdx.x >> 5);
    const int warp_slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // Reference tile: the packed substitution profile of reference 0 is staged once per CTA into shared memory by the
    // TMA engine (cp.async.bulk, completion on an mbarrier); every DP step then reads it with two 16-byte LDS.
    const uint32_t *staged_prof = nullptr;
    if (P.stage_bytes > 0) {
        __shared__ __align__(8) unsigned long long mbar;
        unsigned char *dst = smem_raw + (((size_t)(blockDim.x >> 5) * (sizeof(WarpSmem) + sizeof(QuadSmem)) + 127) & ~(size_t)127);
        const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(&mbar), dst_a = (uint32_t)__cvta_generic_to_shared(dst);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"((uint32_t)P.stage_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst_a), "l"(P.stage_src), "r"((uint32_t)P.stage_bytes), "r"(mbar_a) : "memory");
        }
        asm volatile("{\n .reg .pred p;\n C2B_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra C2B_DONE;\n bra C2B_WAIT;\n C2B_DONE:\n}" ::"r"(mbar_a) : "memory");
        staged_prof = reinterpret_cast<const uint32_t *>(dst);
    }
    // Work groups (8 reads each) are handed out per phase set (g consecutive warps, one group per warp), one hand-out
    // ahead, so that the loop count -- and with it the number of barriers executed by process_group's phases -- is the
    // same for every warp of the set, and the next group's read bytes are on their way to L2 while this one computes.
    if constexpr (!STREAM) {
    __shared__ unsigned long long next_base[WARPS_PER_CTA];
    const int gs = P.phase_sync, g = gs > 0 ? gs : gs < 0 ? -gs : 1, wib = threadIdx.x >> 5, nsets = (int)(blockDim.x >> 5) / g;
    const int set = gs < 0 ? wib % nsets : wib / g, wis = gs < 0 ? wib / nsets : wib % g;
    const unsigned long long total = ((unsigned long long)P.n_reads + 7) / 8;
    auto hand_out = [&]() -> unsigned long long {
        if (wis == 0 && (threadIdx.x & 31) == 0) next_base[set] = atomicAdd(P.work_counter, (unsigned long long)g);
        if (g > 1) wp::grp_sync(gs); else __syncwarp();
        const unsigned long long b = next_base[set];
        if (g > 1) wp::grp_sync(gs); else __syncwarp();
        return b;
    };
    unsigned long long base = hand_out();
    while (base < total) {
        const unsigned long long nb = hand_out();
        const unsigned long long wn = nb + wis;
        if (wn < total && !P.pair_order) {
            const int64_t last = (int64_t)(8 * wn + 8) < P.n_reads ? (int64_t)(8 * wn + 8) : P.n_reads;
            const int64_t b0 = P.offsets[8 * wn], b1 = P.offsets[last];
            const int64_t a = b0 + (int64_t)(threadIdx.x & 31) * 128;
            if (a < b1) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.reads + a));
        }
        const unsigned long long w = base + wis;
        if (w < total) process_group<ONE>(P, *S, *Q, staged_prof, (int64_t)w, warp_slot);  // reads 8w .. 8w+7
        else if (P.phase_sync) for (int b = group_phases(P); b > 0; b--) wp::grp_sync(gs);
        __syncwarp();
        base = nb;
    }
    } else {
    __shared__ unsigned long long next_base[2 * WARPS_PER_CTA];   // [set]: next hand-out; [WARPS_PER_CTA + set]: wait timed out
    const int gs = P.phase_sync, g = gs > 0 ? gs : gs < 0 ? -gs : 1, wib = threadIdx.x >> 5, nsets = (int)(blockDim.x >> 5) / g;
    const int set = gs < 0 ? wib % nsets : wib / g, wis = gs < 0 ? wib / nsets : wib % g;
    const unsigned long long total = ((unsigned long long)P.n_reads + 7) / 8;
    // Streamed launch: before a set starts on work groups base..base+g-1 its leader waits until their read bytes have
    // arrived (P.avail is advanced by the copy stream after each chunk's H2D).  A wait longer than 20 s is reported in
    // stats[7] and ends the launch instead of hanging the GPU.
    auto hand_out = [&](unsigned long long cur) -> unsigned long long {
        if (wis == 0 && (threadIdx.x & 31) == 0) {
            unsigned long long nxt = atomicAdd(P.work_counter, (unsigned long long)g);
            if (P.avail && cur < total) {
                const unsigned long long need = cur + g < total ? cur + g : total;
                unsigned long long t0 = 0, have;
                for (;;) {
                    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(have) : "l"(P.avail));
                    if (have >= need) break;
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (!t0) t0 = now;
                    if (now - t0 > 20000000000ull) { atomicExch(P.stats + 7, 1ull); nxt = ~0ull; next_base[WARPS_PER_CTA + set] = 1; break; }
                    __nanosleep(1000);
                }
            }
            next_base[set] = nxt;
        }
        if (g > 1) wp::grp_sync(gs); else __syncwarp();
        const unsigned long long b = next_base[set];
        if (g > 1) wp::grp_sync(gs); else __syncwarp();
        return b;
    };
    if (threadIdx.x < 2 * WARPS_PER_CTA) next_base[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long base = hand_out(~0ull);              // nothing to wait for yet
    int chunk = 0;
    while (base < total) {
        const unsigned long long nb = hand_out(base);        // next hand-out; returns once `base` itself is resident
        if (next_base[WARPS_PER_CTA + set]) break;            // the wait timed out
        const unsigned long long wn = nb + wis;
        if (wn < total && !P.pair_order && !P.avail) {
            const int64_t last = (int64_t)(8 * wn + 8) < P.n_reads ? (int64_t)(8 * wn + 8) : P.n_reads;
            const int64_t b0 = P.offsets[8 * wn], b1 = P.offsets[last];
            const int64_t a = b0 + (int64_t)(threadIdx.x & 31) * 128;
            if (a < b1) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.reads + a));
        }
        const unsigned long long w = base + wis;
        if (w < total) {
            process_group<ONE>(P, *S, *Q, staged_prof, (int64_t)w, warp_slot);  // reads 8w .. 8w+7
            if (P.chunk_done) {                              // streamed launch: tell the host that this group's outputs are complete
                __syncwarp();
                while (w >= P.chunk_end[chunk]) chunk++;
                const int lane = threadIdx.x & 31;
                unsigned long long wmax = 0;
                if (lane < 8 && 8 * (int64_t)w + lane < P.n_reads) {
                    const int64_t rd = P.pair_order ? P.pair_order[8 * (int64_t)w + lane] : 8 * (int64_t)w + lane;   // the reads this group aligned
                    for (int r = 0; r < P.out_refs; r++) {
                        const unsigned long long v = *reinterpret_cast<const volatile uint16_t *>(&P.alns[rd * P.out_refs + r].aln_len);
                        wmax = v > wmax ? v : wmax;
                    }
                }
#pragma unroll
                for (int d = 4; d >= 1; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, wmax, d); wmax = o > wmax ? o : wmax; }
                if (lane == 0) {
                    atomicMax(P.chunk_done + P.n_chunks + chunk, wmax);
                    __threadfence();
                    atomicAdd(P.chunk_done + chunk, 1ull);
                }
            }
        }
        else if (P.phase_sync) for (int b = group_phases(P); b > 0; b--) wp::grp_sync(gs);
        __syncwarp();
        base = nb;
    }
    }
}
#endif

// ----------------------------------------------------------------------------------------- engine
struct RefHost {
    std::string seq; std::vector<int64_t> gi, inc; double min_aln; std::vector<int64_t> rows;
    std::vector<std::string> fw, rc;
};

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
};

struct c2b_engine {
    int device = 0;
    rt_stream stream = 0;
    std::string err;
    bool configured = false;
    c2b_params prm;
    int n_refs = 0, max_I = 0, max_nrb = 1, vstride = 0, hstride = 0, hist_zero = 0;
    std::vector<RefHost> refs;
    std::vector<RefDev> refdev;         // host mirror (device pointers inside)
    void *d_tables = nullptr; RefDev *d_refs = nullptr;
    unsigned long long *d_counts = nullptr; size_t counts_n = 0;
    // scratch
    DevBuf tb, tbb, tbq, bnd, ops, rgo, work, lut;
    int n_warps = 0, grid = 0, wpc = 8, stage_cap = 0;
    int scratch_TS = 0;
    // staging for the host-pointer API: two buffer sets, copy-in / compute / copy-out streams
    struct Stage { DevBuf reads, off, cnt, qw, rid, recs, alns, str, ed, maxlen, ord; int32_t *h_ord = nullptr; size_t h_ord_cap = 0; int64_t *h_off = nullptr; size_t h_off_cap = 0;
                   rt_event in_done, k_done, out_done; bool used = false; } stage[2];
    rt_stream s_in = 0, s_out = 0, stream2 = 0;     // stream2: second compute stream, kernels of odd chunks
    rt_event fork_ev = 0;                           // orders stream2 after what is already queued on `stream`
    size_t set_tb = 0, set_tbb = 0, set_tbq = 0, set_bnd = 0, set_ops = 0, set_rgo = 0;   // bytes per scratch set (two sets: kernels of
                                                    // consecutive chunks overlap their tail / head on the two streams)
    bool pipe_ready = false;
    // streamed launch (one persistent launch per host batch): whole-batch device buffers + control block
    struct Streamed { DevBuf reads, off, cnt, qw, rid, ord, recs, alns, str, ed, ctl; unsigned long long *h_ctl = nullptr; size_t h_ctl_cap = 0;
                      int64_t *h_off = nullptr; size_t h_off_cap = 0; int32_t *h_ord = nullptr; size_t h_ord_cap = 0; rt_stream s_poll = 0; bool ready = false; } sm;
    double last_ms = 0; int64_t launches = 0;
    const uint64_t *forced_ops = nullptr; const int32_t *forced_n = nullptr;
    const unsigned long long *k_avail = nullptr, *k_chunk_end = nullptr; unsigned long long *k_chunk_done = nullptr; int k_n_chunks = 0;   // streamed launch (set around launch_on)
    const int32_t *pair_order = nullptr;
    int64_t band_reruns = 0, ring_pairs = 0, ring_fallbacks = 0;
#ifndef C2B_EMU
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
#endif
};

static std::string g_create_err;

static int fail(c2b_engine *e, int code, const std::string &m) { if (e) e->err = m; else g_create_err = m; return code; }

#define RTCHK(call) do { rt_err _r = (call); if (_r != RT_OK) return fail(e, C2B_E_CUDA, std::string(#call) + ": " + rt_errstr(_r)); } while (0)

static int ensure(c2b_engine *e, DevBuf &b, size_t n)
{
    if (n <= b.cap) return C2B_OK;
    if (b.p) rt_free(b.p);
    b.p = nullptr; b.cap = 0;
    size_t want = n + n / 8 + 256;
    RTCHK(rt_malloc(&b.p, want));
    b.cap = want;
    return C2B_OK;
}

extern "C" {

int c2b_create(int device, c2b_engine **out)
{
    if (!out) return C2B_E_ARG;
    c2b_engine *e = new c2b_engine();
    e->device = device;
#ifndef C2B_EMU
    cudaError_t r = cudaSetDevice(device);
    if (r == cudaSuccess) r = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (r == cudaSuccess) r = cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking);
    if (r == cudaSuccess) r = cudaEventCreateWithFlags(&e->fork_ev, cudaEventDisableTiming);
    if (r == cudaSuccess) r = cudaEventCreate(&e->ev0);
    if (r == cudaSuccess) r = cudaEventCreate(&e->ev1);
    int nsm = 0, occ = 0, smem_sm = 0;
    if (r == cudaSuccess) r = cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    if (r == cudaSuccess) r = cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    // room for a TMA-staged reference tile next to C2B_MIN_CTAS_PER_SM CTAs of per-warp state (1 KB per CTA is reserved by the driver)
    // ... and the kernels' static shared memory (up to 1.3 KB: hand-out slots, mbarrier) -- if the sum is a byte too large the
    // occupancy query answers 1 CTA per SM and the persistent grid silently halves (r01k: 37 ms instead of 27)
    e->stage_cap = smem_sm / C2B_MIN_CTAS_PER_SM - 1024 - (int)((sizeof(WarpSmem) + sizeof(QuadSmem)) * WARPS_PER_CTA) - 2048;
    {   // ... and within the per-block opt-in limit
        int optin = 0;
        if (r == cudaSuccess) r = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        const int fixed = (int)((sizeof(WarpSmem) + sizeof(QuadSmem)) * WARPS_PER_CTA) + 128;
        if (e->stage_cap > optin - fixed) e->stage_cap = optin - fixed;
    }
    if (e->stage_cap < 0) e->stage_cap = 0;
    e->stage_cap &= ~127;
    const int dyn_smem = (int)((sizeof(WarpSmem) + sizeof(QuadSmem)) * WARPS_PER_CTA) + 128 + e->stage_cap;
    {
        const void *kernels[4] = {(const void *)c2b_align_classify_kernel<true, false>, (const void *)c2b_align_classify_kernel<false, false>,
                                  (const void *)c2b_align_classify_kernel<true, true>, (const void *)c2b_align_classify_kernel<false, true>};
        occ = 1 << 20;
        for (const void *k : kernels) {                    // all instantiations must fit the same persistent grid
            int o = 0;
            if (r == cudaSuccess) r = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
            if (r == cudaSuccess) r = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k, WARPS_PER_CTA * 32, dyn_smem);
            occ = std::min(occ, o);
        }
    }
    if (r != cudaSuccess) { g_create_err = std::string("c2b_create: ") + cudaGetErrorString(r); delete e; return C2B_E_CUDA; }
    if (occ < 1) occ = 1;
    if (occ < C2B_MIN_CTAS_PER_SM && !getenv("C2B_CTAS_PER_SM"))
        fprintf(stderr, "[c2b] warning: only %d CTA(s) of the align kernel fit an SM (built for %d)\n", occ, C2B_MIN_CTAS_PER_SM);
    e->wpc = WARPS_PER_CTA;
    if (const char *v = getenv("C2B_WARPS_PER_CTA")) { int k = atoi(v); if (k >= 1 && k <= WARPS_PER_CTA) e->wpc = k; }
    if (const char *v = getenv("C2B_CTAS_PER_SM")) { int k = atoi(v); if (k >= 1 && k <= occ) occ = k; }
    e->grid = nsm * occ;                 // persistent: one wave of CTAs, warps pull work items from a counter
    e->n_warps = e->grid * e->wpc;
#else
    e->grid = 1; e->n_warps = 1;
#endif
    *out = e;
    return C2B_OK;
}

void c2b_destroy(c2b_engine *e)
{
    if (!e) return;
    DevBuf *bufs[] = {&e->tb, &e->tbb, &e->tbq, &e->bnd, &e->ops, &e->rgo, &e->work, &e->lut};
    for (DevBuf *b : bufs) if (b->p) rt_free(b->p);
    for (auto &st : e->stage) {
        DevBuf *sb[] = {&st.reads, &st.off, &st.cnt, &st.qw, &st.rid, &st.recs, &st.alns, &st.str, &st.ed, &st.maxlen, &st.ord};
        if (st.h_ord) rt_host_free(st.h_ord);
        for (DevBuf *b : sb) if (b->p) rt_free(b->p);
        if (st.h_off) rt_host_free(st.h_off);
        if (e->pipe_ready) { rt_event_destroy(st.in_done); rt_event_destroy(st.k_done); rt_event_destroy(st.out_done); }
    }
    if (e->pipe_ready) { rt_stream_destroy(e->s_in); rt_stream_destroy(e->s_out); }
    {
        DevBuf *mb[] = {&e->sm.reads, &e->sm.off, &e->sm.cnt, &e->sm.qw, &e->sm.rid, &e->sm.ord, &e->sm.recs, &e->sm.alns, &e->sm.str, &e->sm.ed, &e->sm.ctl};
        for (DevBuf *b : mb) if (b->p) rt_free(b->p);
        if (e->sm.h_ctl) rt_host_free(e->sm.h_ctl);
        if (e->sm.h_off) rt_host_free(e->sm.h_off);
        if (e->sm.h_ord) rt_host_free(e->sm.h_ord);
        if (e->sm.ready) rt_stream_destroy(e->sm.s_poll);
    }
    if (e->d_tables) rt_free(e->d_tables);
    if (e->d_counts) rt_free(e->d_counts);
#ifndef C2B_EMU
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->stream2) cudaStreamDestroy(e->stream2);
    if (e->fork_ev) cudaEventDestroy(e->fork_ev);
#endif
    delete e;
}

const char *c2b_last_error(const c2b_engine *e) { return e ? e->err.c_str() : g_create_err.c_str(); }

static uint64_t pack_seed(const c2b_params &p, const std::string &s)
{
    uint64_t v = 0;
    for (size_t c = 0; c < s.size(); c++) {
        int code = -1;
        for (int q = 0; q < p.nq; q++) if (s[c] == p.alphabet[q]) code = q;
        if (code < 0) return ~0ull;                     // can never equal a read k-mer
        v |= (uint64_t)code << (3 * c);
    }
    return v;
}

int c2b_configure(c2b_engine *e, const c2b_params *p, int32_t n_refs, const c2b_ref *refs)
{
    if (!e || !p || !refs || n_refs < 1) return fail(e, C2B_E_ARG, "c2b_configure: bad argument");
    if (n_refs > C2B_MAX_POOLED_REFS) return fail(e, C2B_E_LIMIT, "c2b_configure: more than C2B_MAX_POOLED_REFS references");
    if (p->nq < 1 || p->nq > C2B_MAX_Q) return fail(e, C2B_E_ARG, "c2b_configure: alphabet size out of range");
    if (p->seed_count < 0 || p->edit_cap < 0) return fail(e, C2B_E_ARG, "c2b_configure: negative seed_count/edit_cap");
    e->configured = false;
    e->prm = *p;
    e->n_refs = n_refs;
    e->refs.assign(n_refs, RefHost());
    e->refdev.assign(n_refs, RefDev());
    int maxI = 0, max_nrb = 1;
    size_t bytes = 0;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    std::vector<size_t> base(n_refs);
    for (int r = 0; r < n_refs; r++) {
        const c2b_ref &rf = refs[r];
        if (!rf.seq || rf.len < 1 || !rf.gap_incentive || !rf.score_rows) return fail(e, C2B_E_ARG, "c2b_configure: incomplete reference");
        if (rf.len > C2B_MAX_REF_LEN) return fail(e, C2B_E_LIMIT, "c2b_configure: reference longer than C2B_MAX_REF_LEN");
        if (rf.n_seeds > 0 && (!rf.fw_seeds || !rf.rc_seeds)) return fail(e, C2B_E_ARG, "c2b_configure: seeds missing");
        maxI = std::max(maxI, rf.len);
        const int nrb = (rf.len + 255) / 256, Ipad = nrb * 256;
        max_nrb = std::max(max_nrb, nrb);
        base[r] = bytes;
        bytes += al((size_t)p->nq * Ipad * 4) + 2 * al((size_t)Ipad * 4) + 2 * al(Ipad) + al(Ipad + 1) + 3 * al((size_t)(Ipad + 2) * 2);
        bytes += al((size_t)p->nq * p->nq * Ipad * 4) + 2 * al((size_t)Ipad * 4);      // packed-path tables
    }
    const size_t refs_off = bytes;
    bytes += al(sizeof(RefDev) * n_refs);
    std::vector<unsigned char> blob(bytes, 0);
    if (e->d_tables) { rt_free(e->d_tables); e->d_tables = nullptr; }
    RTCHK(rt_malloc(&e->d_tables, bytes));
    e->vstride = (maxI + 31) & ~31;
    {   // histogram buckets: sizes 0..I, effective lengths 0..I+J, frame keys -I+tem..J+tem around hist_zero
        int max_tem = 0;
        for (int r = 0; r < n_refs; r++) max_tem = std::max(max_tem, std::abs((int)refs[r].tot_exon_len_mod));
        if (max_tem > 4 * C2B_MAX_REF_LEN) return fail(e, C2B_E_LIMIT, "c2b_configure: tot_exon_len_mod out of range");
        e->hist_zero = maxI + max_tem;
        const int need = std::max(e->hist_zero + C2B_MAX_READ_LEN + max_tem, std::min(maxI + C2B_MAX_READ_LEN, C2B_MAX_ALN_LEN)) + 1;
        e->hstride = (need + 31) & ~31;
    }
    const size_t per_ref = C2B_NVEC * (size_t)e->vstride + C2B_NHIST * (size_t)e->hstride + C2B_NSCAL;
    e->counts_n = (size_t)n_refs * per_ref;
    if (e->d_counts) { rt_free(e->d_counts); e->d_counts = nullptr; }
    RTCHK(rt_malloc((void **)&e->d_counts, e->counts_n * 8));
    RTCHK(rt_zero(e->d_counts, e->counts_n * 8, e->stream));

    for (int r = 0; r < n_refs; r++) {
        const c2b_ref &rf = refs[r];
        const int I = rf.len, nrb = (I + 255) / 256, Ipad = nrb * 256;
        RefDev &d = e->refdev[r];
        d.I = I; d.nrb = nrb; d.Ipad = Ipad; d.kstar = (I - 1) & 7; d.lstar = ((I - 1) >> 3) & 31;
        d.min_aln = rf.min_aln_score;
        unsigned char *hb = blob.data() + base[r];
        unsigned char *db = (unsigned char *)e->d_tables + base[r];
        size_t o = 0;
        int32_t *prof = (int32_t *)(hb + o); d.prof = (const int32_t *)(db + o); o += al((size_t)p->nq * Ipad * 4);
        int32_t *cIe = (int32_t *)(hb + o); d.cIe = (const int32_t *)(db + o); o += al((size_t)Ipad * 4);
        int32_t *g4 = (int32_t *)(hb + o); d.g4 = (const int32_t *)(db + o); o += al((size_t)Ipad * 4);
        uint8_t *asc = hb + o; d.asc = db + o; o += al(Ipad);
        uint8_t *rcode = hb + o; d.rcode = db + o; o += al(Ipad);
        uint8_t *incl = hb + o; d.incl = db + o; o += al(Ipad + 1);
        uint16_t *cum = (uint16_t *)(hb + o); d.cum = (const uint16_t *)(db + o); o += al((size_t)(Ipad + 2) * 2);
        uint16_t *cumx = (uint16_t *)(hb + o); d.cumx = (const uint16_t *)(db + o); o += al((size_t)(Ipad + 2) * 2);
        uint16_t *cums = (uint16_t *)(hb + o); d.cums = (const uint16_t *)(db + o); o += al((size_t)(Ipad + 2) * 2);
        uint32_t *prof2 = (uint32_t *)(hb + o); d.prof2 = (const uint32_t *)(db + o); o += al((size_t)p->nq * p->nq * Ipad * 4);
        uint32_t *cIe2 = (uint32_t *)(hb + o); d.cIe2 = (const uint32_t *)(db + o); o += al((size_t)Ipad * 4);
        uint32_t *g42 = (uint32_t *)(hb + o); d.g42 = (const uint32_t *)(db + o); o += al((size_t)Ipad * 4);
        const int64_t lim = (1ll << 27);
        for (int q = 0; q < p->nq; q++)
            for (int i = 0; i < I; i++) {
                const int64_t v = rf.score_rows[(size_t)q * I + i];
                if (v > lim || v < -lim) return fail(e, C2B_E_LIMIT, "c2b_configure: substitution score out of range");
                prof[(size_t)q * Ipad + i] = (int32_t)(4 * v);
            }
        for (int i = 0; i <= I; i++) if (rf.gap_incentive[i] > lim || rf.gap_incentive[i] < -lim) return fail(e, C2B_E_LIMIT, "c2b_configure: gap incentive out of range");
        for (int row = 0; row < I; row++) {
            cIe[row] = (int32_t)(4 * (p->gap_extend + rf.gap_incentive[row + 1]));
            g4[row] = (int32_t)(4 * rf.gap_incentive[row]);
            asc[row] = (uint8_t)rf.seq[row];
            int code = 255;
            for (int q = 0; q < p->nq; q++) if (rf.seq[row] == p->alphabet[q]) code = q;
            rcode[row] = (uint8_t)code;
        }
        d.gi0_4 = (int32_t)(4 * rf.gap_incentive[0]);
        {   // packed 16-bit path: biased scores (beta = -gap_extend per unit of i+j) + offset; see DESIGN.md section 6
            const int64_t go = p->gap_open, ge = p->gap_extend, beta = -ge, OFFu = 512;
            int64_t gmin = rf.gap_incentive[0], gmax = rf.gap_incentive[0], smin = rf.score_rows[0], smax = rf.score_rows[0];
            for (int i = 0; i <= I; i++) { gmin = std::min(gmin, rf.gap_incentive[i]); gmax = std::max(gmax, rf.gap_incentive[i]); }
            for (size_t k = 0; k < (size_t)p->nq * I; k++) { smin = std::min(smin, rf.score_rows[k]); smax = std::max(smax, rf.score_rows[k]); }
            bool ok = go <= ge && ge <= 0 && gmin >= 0 && gmax <= 64 && smin + 2 * beta >= 0 && smax + 2 * beta <= 1000 &&
                      (go - ge) > -1900 && 4 * (OFFu + (go - ge)) > 256 + 4 * gmax + 3 + 64 && !(p->flags & C2B_F_NO_PAIRING);
            d.pk_maxJ = 0;
            if (ok) {
                for (int J = 1; J <= C2B_MAX_READ_LEN && I + J <= PK_MAX_ALN; J++) {
                    const int64_t bound = 4 * ((smax + 2 * beta) * std::min(I, J) + gmax * (I + J + 2) + OFFu) + 3;
                    if (bound > 32000) break;
                    d.pk_maxJ = J;
                }
            }
            const uint32_t rep = 0x00010001u;
            d.pk_XB = (uint32_t)((4 * (rf.gap_incentive[0] + OFFu)) | 2) * rep;
            d.pk_YB = (uint32_t)((4 * (rf.gap_incentive[0] + OFFu)) | 1) * rep;
            d.pk_M00 = (uint32_t)(4 * OFFu) * rep;
            {   // ring-banded path: the out-of-band score bound (ring_bound) must be decreasing in the number of gap columns
                int64_t gsum = 0;
                for (int i = 0; i <= I; i++) gsum += rf.gap_incentive[i];
                d.rg_smax = (int32_t)smax; d.rg_gmax = (int32_t)gmax; d.rg_gsum = (int32_t)std::min<int64_t>(gsum, 1 << 24);
                d.rg_ok = d.pk_maxJ > 0 && nrb == 1 && smax >= 0 && 2 * ge + gmax <= smax && gsum < (1 << 24) && !(p->flags & C2B_F_NO_RING);
            }
            if (d.pk_maxJ > 0) {
                for (int qa = 0; qa < p->nq; qa++)
                    for (int qb = 0; qb < p->nq; qb++)
                        for (int i = 0; i < I; i++) {
                            const uint32_t a = (uint32_t)(4 * (rf.score_rows[(size_t)qa * I + i] + 2 * beta));
                            const uint32_t b = (uint32_t)(4 * (rf.score_rows[(size_t)qb * I + i] + 2 * beta));
                            const int rbk = i >> 8, ln = (i >> 3) & 31, hf = (i >> 2) & 1, wd = i & 3;
                            prof2[((size_t)qa * p->nq + qb) * Ipad + rbk * 256 + hf * 128 + ln * 4 + wd] = a | (b << 16);
                        }
                for (int row = 0; row < I; row++) {
                    cIe2[row] = (uint32_t)(4 * rf.gap_incentive[row + 1]) * rep;
                    g42[row] = (uint32_t)(4 * rf.gap_incentive[row]) * rep;
                }
            }
        }
        for (int k = 0; k < rf.n_include; k++) {
            const int64_t v = rf.include_idx[k];
            if (v >= 0 && v < I) incl[v] = 1;
        }
        cum[0] = 0;
        for (int q = 0; q <= Ipad; q++) cum[q + 1] = (uint16_t)(cum[q] + (q < I && incl[q] ? 1 : 0));
        d.coding = rf.coding_mask != nullptr; d.tem = rf.tot_exon_len_mod; d.hist_zero = e->hist_zero;
        if (rf.coding_mask)
            for (int q = 0; q < I; q++) incl[q] |= (uint8_t)((rf.coding_mask[q] & 3) << 1);   // after cum[]: bit 0 stays the window
        cumx[0] = cums[0] = 0;
        for (int q = 0; q <= Ipad; q++) {
            cumx[q + 1] = (uint16_t)(cumx[q] + (q < I && (incl[q] & 2) ? 1 : 0));
            cums[q + 1] = (uint16_t)(cums[q] + (q < I && (incl[q] & 4) ? 1 : 0));
        }
        const int ns = std::min({(int)rf.n_seeds, (int)p->seed_count, (int)C2B_MAX_SEEDS});
        if (rf.n_seeds > 0 && std::min((int)rf.n_seeds, (int)p->seed_count) > C2B_MAX_SEEDS)
            return fail(e, C2B_E_LIMIT, "c2b_configure: more seeds than C2B_MAX_SEEDS");
        d.nseeds = ns; d.seed_len = 0;
        for (int s = 0; s < C2B_MAX_SEEDS; s++) { d.fw_seed[s] = ~0ull; d.rc_seed[s] = ~0ull; }
        for (int s = 0; s < ns; s++) {
            const std::string f = rf.fw_seeds[s], c = rf.rc_seeds[s];
            if (f.size() != c.size() || f.empty() || f.size() > C2B_MAX_SEED_LEN) return fail(e, C2B_E_LIMIT, "c2b_configure: seed length unsupported");
            if (s && (int)f.size() != d.seed_len) return fail(e, C2B_E_LIMIT, "c2b_configure: seeds of unequal length");
            d.seed_len = (int)f.size();
            d.fw_seed[s] = pack_seed(*p, f); d.rc_seed[s] = pack_seed(*p, c);
        }
        d.vec = e->d_counts + (size_t)r * per_ref;
        d.hist = d.vec + C2B_NVEC * (size_t)e->vstride;
        d.scal = d.hist + C2B_NHIST * (size_t)e->hstride;
    }
    memcpy(blob.data() + refs_off, e->refdev.data(), sizeof(RefDev) * n_refs);
    e->d_refs = (RefDev *)((unsigned char *)e->d_tables + refs_off);
    RTCHK(rt_h2d(e->d_tables, blob.data(), bytes, e->stream));
    RTCHK(rt_sync(e->stream));
    {
        unsigned char lut[256];
        memset(lut, 255, sizeof lut);
        for (int q = 0; q < p->nq; q++) lut[(unsigned char)p->alphabet[q]] = (unsigned char)q;
        int rc2;
        if ((rc2 = ensure(e, e->lut, 256))) return rc2;
        RTCHK(rt_h2d(e->lut.p, lut, 256, e->stream));
        RTCHK(rt_sync(e->stream));
    }
    e->max_I = maxI; e->max_nrb = max_nrb;
    e->scratch_TS = 0;
    e->configured = true;
    return C2B_OK;
}

int c2b_set_edit_cap(c2b_engine *e, int32_t edit_cap)
{
    if (!e || !e->configured || edit_cap < 0) return fail(e, C2B_E_ARG, "c2b_set_edit_cap: bad argument");
    e->prm.edit_cap = edit_cap;
    return C2B_OK;
}

int c2b_string_width(const c2b_engine *e, int32_t max_read_len)
{
    if (!e || !e->configured) return C2B_E_STATE;
    return (e->max_I + max_read_len + 31) & ~31;
}

// work block (u64): [2..6] cumulative path statistics; set s: [8 + 8 s] work hand-out counter, [9 + 8 s] widest alignment
constexpr size_t WORK_BYTES = 24 * 8;

static int ensure_scratch(c2b_engine *e, int maxJ)
{
    const int TS = ((maxJ + 32 + 31) & ~31);
    if (TS <= e->scratch_TS) return C2B_OK;
    int rc;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    e->set_tb = al((size_t)e->n_warps * e->max_nrb * TS * 64 * 4);      // 64: a pair stores two words per lane
    e->set_tbb = al((size_t)e->n_warps * PK_BAND_SLOTS * 64 * 4);        // banded slabs (packed path)
    e->set_tbq = al((size_t)e->n_warps * TS * 64 * 4 + 64);              // ring-banded path: (step, lane) entries
    e->set_bnd = al((size_t)e->n_warps * 2 * 3 * TS * 4);
    e->set_ops = al((size_t)e->n_warps * std::min(e->n_refs, (int)C2B_MAX_REFS) * 32 * 8);
    if ((rc = ensure(e, e->tb, 2 * e->set_tb))) return rc;
    if ((rc = ensure(e, e->tbb, 2 * e->set_tbb))) return rc;
    if ((rc = ensure(e, e->tbq, 2 * e->set_tbq))) return rc;
    if ((rc = ensure(e, e->bnd, 2 * e->set_bnd))) return rc;
    if ((rc = ensure(e, e->ops, 2 * e->set_ops))) return rc;
    e->set_rgo = al((size_t)e->n_warps * RG_MAX_REFS * 4 * RG_OPS_STRIDE * 8);
    if ((rc = ensure(e, e->rgo, 2 * e->set_rgo))) return rc;
    const bool fresh_work = !e->work.p;
    if ((rc = ensure(e, e->work, WORK_BYTES))) return rc;
    if (fresh_work) RTCHK(rt_zero(e->work.p, WORK_BYTES, e->stream));
    e->scratch_TS = TS;
#ifndef C2B_EMU
    {   // Keep the traceback slab (written once, read back by the same warp microseconds later) resident in L2:
        // persisting window over the slab, everything else on this stream streams through the rest of L2.
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, e->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 &&
            !getenv("C2B_NO_L2_PERSIST")) {
            const size_t slab = 2 * e->set_tbb;                                   // the banded slabs (both sets): the hot set
            const size_t carve = std::min((size_t)prop.persistingL2CacheMaxSize, slab);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
            cudaStreamAttrValue av; memset(&av, 0, sizeof av);
            av.accessPolicyWindow.base_ptr = e->tbb.p;
            av.accessPolicyWindow.num_bytes = std::min(slab, (size_t)prop.accessPolicyMaxWindowSize);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)std::max<size_t>(1, av.accessPolicyWindow.num_bytes));
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &av);
            if (e->stream2) cudaStreamSetAttribute(e->stream2, cudaStreamAttributeAccessPolicyWindow, &av);
            cudaGetLastError();
            if (getenv("C2B_VERBOSE"))
                fprintf(stderr, "[c2b] grid %d x %d warps, banded traceback slabs %.1f MB, persisting L2 max %.1f MB (L2 %.1f MB), window %.1f MB, hitRatio %.2f\n",
                        e->grid, e->wpc, slab / 1e6, prop.persistingL2CacheMaxSize / 1e6, prop.l2CacheSize / 1e6,
                        av.accessPolicyWindow.num_bytes / 1e6, av.accessPolicyWindow.hitRatio);
        }
    }
#endif
    return C2B_OK;
}

// One launch on compute stream `cs` using scratch set `set` (0 or 1).  Launches that may overlap in time must use
// different sets; launches on the same stream are ordered.
static int launch_on(c2b_engine *e, rt_stream cs, int set, const uint8_t *d_reads, const int64_t *d_offsets, int64_t n_reads,
                     int32_t max_read_len, const int32_t *d_count, const int32_t *d_qweight,
                     const int32_t *d_ref_id, c2b_read_rec *d_recs, c2b_aln_rec *d_alns,
                     uint8_t *d_strings, c2b_edit *d_edits)
{
    if (!e || !e->configured) return fail(e, C2B_E_STATE, "c2b_align_batch: engine not configured");
    if (n_reads < 0 || !d_recs || !d_alns || (n_reads && (!d_reads || !d_offsets))) return fail(e, C2B_E_ARG, "c2b_align_batch: bad argument");
    if (max_read_len < 1) max_read_len = 1;
    if (max_read_len > C2B_MAX_READ_LEN) return fail(e, C2B_E_LIMIT, "c2b_align_batch: read longer than C2B_MAX_READ_LEN");
    if ((int64_t)std::abs((long long)e->prm.gap_open) * max_read_len * e->max_I >= (1ll << 28))
        return fail(e, C2B_E_LIMIT, "c2b_align_batch: gap_open * lengths exceeds the int32 score range");
    if (!d_ref_id && e->n_refs > C2B_MAX_REFS)
        return fail(e, C2B_E_LIMIT, "c2b_align_batch: more than C2B_MAX_REFS references need a per-read ref_id");
    int rc = ensure_scratch(e, max_read_len);
    if (rc) return rc;
    if (n_reads == 0) return C2B_OK;
    KParams P;
    memset(&P, 0, sizeof P);
    P.reads = d_reads; P.offsets = d_offsets; P.n_reads = n_reads; P.count = d_count; P.qweight = d_qweight; P.ref_id = d_ref_id;
    P.recs = d_recs; P.alns = d_alns; P.strings = d_strings; P.edits = d_edits;
    P.W = (e->max_I + max_read_len + 31) & ~31; P.edit_cap = d_edits ? e->prm.edit_cap : 0;
    if (P.edit_cap == 0) P.edits = nullptr;
    P.refs = e->d_refs; P.n_refs = e->n_refs;
    P.out_refs = d_ref_id ? 1 : e->n_refs; P.ops_refs = std::min(e->n_refs, (int)C2B_MAX_REFS);
    P.go = e->prm.gap_open; P.ge = e->prm.gap_extend; P.seed_count = e->prm.seed_count; P.seed_min = e->prm.seed_min;
    P.flags = e->prm.flags; P.nq = e->prm.nq;
    memcpy(P.alpha, e->prm.alphabet, C2B_MAX_Q); memcpy(P.comp, e->prm.complement, C2B_MAX_Q);
    P.TS = e->scratch_TS;
    auto at = [&](const DevBuf &b, size_t per_set) { return (char *)b.p + (size_t)set * per_set; };
    P.tb = (uint32_t *)at(e->tb, e->set_tb); P.tb_words_per_warp = (int64_t)e->max_nrb * P.TS * 64;
    P.tbb = getenv("C2B_NO_BAND") ? nullptr : (uint32_t *)at(e->tbb, e->set_tbb); P.tbb_words_per_warp = (int64_t)PK_BAND_SLOTS * 64;
    P.tbq = getenv("C2B_NO_RING") ? nullptr : (uint32_t *)at(e->tbq, e->set_tbq);
    P.bnd = (int32_t *)at(e->bnd, e->set_bnd); P.bnd_words_per_warp = 2 * 3 * (int64_t)P.TS;
    P.opsbuf = (uint64_t *)at(e->ops, e->set_ops);
    P.rgops = getenv("C2B_NO_MULTI_RING") ? nullptr : (uint64_t *)at(e->rgo, e->set_rgo);
    P.stats = (unsigned long long *)e->work.p;
    P.work_counter = P.stats + 8 + 8 * set;
    P.vstride = e->vstride; P.hstride = e->hstride;
    P.forced_ops = e->forced_ops; P.forced_n = e->forced_n;
    P.avail = e->k_avail; P.chunk_end = e->k_chunk_end; P.chunk_done = e->k_chunk_done; P.n_chunks = e->k_n_chunks;
    P.phase_sync = 4;                                     // warps per phase set (C2B_PHASE_WARPS: 0/1 = free-running, 2, 4, 8)
    if (const char *v = getenv("C2B_PHASE_WARPS")) { const int k = atoi(v), a = k < 0 ? -k : k; P.phase_sync = (a == 2 || a == 4 || a == 8 || a == 16) ? k : 0; }
    {   // grp_sync wants |g| and the number of sets to be powers of two
        const int a = P.phase_sync < 0 ? -P.phase_sync : P.phase_sync;
        const int nsets = a ? e->wpc / a : 0;
        if (a > e->wpc || (a && e->wpc % a) || (P.phase_sync < 0 && (nsets & (nsets - 1)))) P.phase_sync = 0;
    }
    if (e->n_refs > 1 && !d_ref_id && getenv("C2B_NO_MULTI_PHASE")) P.phase_sync = 0;     // A/B switch: free-running warps in multi-reference mode
    P.pair_order = e->pair_order;
    P.lut = (const uint8_t *)e->lut.p;
    P.stage_bytes = 0; P.stage_src = nullptr;
#ifndef C2B_EMU
    {   // stage reference 0's packed profile when it exists and fits beside two CTAs' worth of per-warp state
        const RefDev &r0 = e->refdev[0];
        const size_t bytes = (size_t)e->prm.nq * e->prm.nq * r0.Ipad * 4;
        if (r0.pk_maxJ > 0 && !e->forced_ops && bytes <= (size_t)e->stage_cap && !getenv("C2B_NO_TMA_STAGE")) {
            P.stage_bytes = (int32_t)bytes; P.stage_src = r0.prof2;
        }
    }
#endif
    RTCHK(rt_zero(P.work_counter, 16, cs));               // this set's work counter and widest alignment
#ifndef C2B_EMU
    cudaEventRecord(e->ev0, cs);
    {
        const size_t smem = (sizeof(WarpSmem) + sizeof(QuadSmem)) * e->wpc + 128 + (size_t)P.stage_bytes;
        const bool one = (e->n_refs == 1 || d_ref_id != nullptr) && !getenv("C2B_GENERIC_KERNEL");      // one candidate reference per read
        const bool stream = P.avail != nullptr;
        if (one && !stream) c2b_align_classify_kernel<true, false><<<e->grid, e->wpc * 32, smem, cs>>>(P);
        else if (!stream) c2b_align_classify_kernel<false, false><<<e->grid, e->wpc * 32, smem, cs>>>(P);
        else if (one) c2b_align_classify_kernel<true, true><<<e->grid, e->wpc * 32, smem, cs>>>(P);
        else c2b_align_classify_kernel<false, true><<<e->grid, e->wpc * 32, smem, cs>>>(P);
    }
    cudaEventRecord(e->ev1, cs);
    RTCHK(cudaGetLastError());
#else
    {
        static WarpSmem S; static QuadSmem Q;
        for (int64_t w = 0; 8 * w < n_reads; w++) {
            wp::g_grp_syncs = 0;
            const bool one = (e->n_refs == 1 || d_ref_id != nullptr) && !getenv("C2B_GENERIC_KERNEL");
            if (one) emu::run_warp([&]() { process_group<true>(P, S, Q, nullptr, w, 0); });
            else emu::run_warp([&]() { process_group<false>(P, S, Q, nullptr, w, 0); });
            // every path through a work group must execute the same number of phase barriers (a mismatch deadlocks the GPU)
            if (w == 0 && getenv("C2B_EMU_VERBOSE")) fprintf(stderr, "warp_emu: phase_sync %d, %ld barriers in group 0 (expected %d)\n", P.phase_sync, wp::g_grp_syncs, group_phases(P));
            if (P.phase_sync && wp::g_grp_syncs != group_phases(P)) {
                fprintf(stderr, "warp_emu: work group %lld executed %ld phase barriers, expected %d\n", (long long)w, wp::g_grp_syncs, group_phases(P));
                abort();
            }
        }
    }
#endif
    e->launches++;
    return C2B_OK;
}

int c2b_align_batch_device(c2b_engine *e, const uint8_t *d_reads, const int64_t *d_offsets, int64_t n_reads,
                           int32_t max_read_len, const int32_t *d_count, const int32_t *d_qweight,
                           const int32_t *d_ref_id, c2b_read_rec *d_recs, c2b_aln_rec *d_alns,
                           uint8_t *d_strings, c2b_edit *d_edits)
{
    return launch_on(e, e ? e->stream : 0, 0, d_reads, d_offsets, n_reads, max_read_len, d_count, d_qweight, d_ref_id, d_recs,
                     d_alns, d_strings, d_edits);
}

int64_t c2b_band_reruns(c2b_engine *e) { return e ? e->band_reruns : 0; }

int c2b_set_pair_order(c2b_engine *e, const int32_t *d_order)
{
    if (!e) return C2B_E_ARG;
    e->pair_order = d_order;
    return C2B_OK;
}

int c2b_sync(c2b_engine *e)
{
    if (!e) return C2B_E_ARG;
    RTCHK(rt_sync(e->stream));
    return C2B_OK;
}

void *c2b_stream(c2b_engine *e) { return e ? (void *)(uintptr_t)e->stream : nullptr; }

double c2b_last_kernel_ms(c2b_engine *e)
{
#ifndef C2B_EMU
    if (!e || !e->launches) return 0.0;
    float ms = 0.f;
    if (cudaEventSynchronize(e->ev1) != cudaSuccess) return 0.0;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) != cudaSuccess) return 0.0;
    return (double)ms;
#else
    (void)e; return 0.0;
#endif
}

int64_t c2b_launch_count(const c2b_engine *e) { return e ? e->launches : 0; }

int c2b_path_counts(c2b_engine *e, int64_t *pair_items, int64_t *single_items)
{
    if (!e || !e->work.p) return fail(e, C2B_E_STATE, "c2b_path_counts: nothing launched yet");
    int64_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    RTCHK(rt_d2h(v, e->work.p, 64, e->stream));
    RTCHK(rt_sync(e->stream));
    e->band_reruns = v[4]; e->ring_pairs = v[5]; e->ring_fallbacks = v[6];
    if (pair_items) *pair_items = v[2];
    if (single_items) *single_items = v[3];
    return C2B_OK;
}

int c2b_ring_counts(c2b_engine *e, int64_t *ring_pairs, int64_t *ring_fallbacks)
{
    if (!e) return C2B_E_ARG;
    if (ring_pairs) *ring_pairs = e->ring_pairs;
    if (ring_fallbacks) *ring_fallbacks = e->ring_fallbacks;
    return C2B_OK;
}

#ifndef C2B_EMU
// C2B_STREAMED=1 / =0 forces the streamed launch on / off; default on (r01k: e2e 30.6 ms against 34.5 ms per 1 M reads).
static bool streamed_default(const c2b_engine *) { const char *v = getenv("C2B_STREAMED"); return v ? atoi(v) != 0 : true; }

// Host batch through ONE persistent launch: the kernel starts at once and takes work groups as their read bytes arrive
// (chunked H2D on the copy stream, each followed by an 8-byte update of the "groups resident" mark the kernel polls);
// every finished group bumps its chunk's counter, the host polls those and queues each chunk's D2H as soon as the chunk is
// complete.  No launch head/tail per chunk, copies of both directions overlap the kernel.
static int align_batch_streamed(c2b_engine *e, const uint8_t *reads, const int64_t *offsets, int64_t n_reads, int64_t maxJ,
                                const int32_t *count, const int32_t *qweight, const int32_t *ref_id,
                                c2b_read_rec *recs, c2b_aln_rec *alns, uint8_t *strings, c2b_edit *edits)
{
    c2b_engine::Streamed &m = e->sm;
    int rc;
    if (!m.ready) { RTCHK(rt_stream_create(&m.s_poll)); m.ready = true; }
    const int W = (e->max_I + (int)maxJ + 31) & ~31;
    const int cap = edits ? e->prm.edit_cap : 0;
    const int nr = ref_id ? 1 : e->n_refs;
    const int64_t b0 = offsets[0], nbytes = offsets[n_reads] - b0;
    // chunk boundaries at multiples of 64 reads (whole hand-outs of a phase set); small first and last chunks
    int64_t chunk = 1 << 16;
    if (const char *v = getenv("C2B_STREAM_CHUNK")) chunk = std::max<int64_t>(64, atoll(v) / 64 * 64);
    std::vector<int64_t> cuts;
    cuts.push_back(0);
    {
        const int64_t edge = std::max<int64_t>(64, chunk / 8 / 64 * 64);
        int64_t pos = 0;
        if (n_reads > 4 * edge) { pos = edge; cuts.push_back(pos); }
        const int64_t tail = (n_reads - pos > 2 * edge) ? edge : 0;
        const int64_t body_end = (n_reads - tail) / 64 * 64;
        while (pos < body_end) { pos = std::min(pos + chunk, body_end); cuts.push_back(pos); }
        if (pos < n_reads) cuts.push_back(n_reads);
    }
    const int nc = (int)cuts.size() - 1;
    if ((rc = ensure_scratch(e, (int)maxJ))) return rc;
    if ((rc = ensure(e, m.reads, (size_t)nbytes + 256))) return rc;
    if ((rc = ensure(e, m.off, (size_t)(n_reads + 1) * 8))) return rc;
    if ((rc = ensure(e, m.recs, (size_t)n_reads * sizeof(c2b_read_rec)))) return rc;
    if ((rc = ensure(e, m.alns, (size_t)n_reads * nr * sizeof(c2b_aln_rec)))) return rc;
    if (strings && (rc = ensure(e, m.str, (size_t)n_reads * nr * 2 * W))) return rc;
    if (cap && (rc = ensure(e, m.ed, (size_t)n_reads * nr * cap * sizeof(c2b_edit)))) return rc;
    if (count && (rc = ensure(e, m.cnt, (size_t)n_reads * 4))) return rc;
    if (qweight && (rc = ensure(e, m.qw, (size_t)n_reads * 4))) return rc;
    if (ref_id && (rc = ensure(e, m.rid, (size_t)n_reads * 4))) return rc;
    // control block (u64): [0] groups resident, [8 .. 8+nc) chunk end groups, [8+nc .. 8+2nc) groups done, [8+2nc .. 8+3nc) widest alignment
    const size_t ctl_n = 8 + 3 * (size_t)nc;
    if ((rc = ensure(e, m.ctl, ctl_n * 8))) return rc;
    const size_t pin_n = 2 * ctl_n + (size_t)nc + 8;       // pinned: initial image | per-chunk "resident" marks | poll buffer
    if (m.h_ctl_cap < pin_n) {
        if (m.h_ctl) rt_host_free(m.h_ctl);
        m.h_ctl = (unsigned long long *)rt_host_alloc(pin_n * 8); m.h_ctl_cap = m.h_ctl ? pin_n : 0;
        if (!m.h_ctl) return fail(e, C2B_E_CUDA, "c2b_align_batch: pinned allocation failed");
    }
    if (m.h_off_cap < (size_t)(n_reads + 1)) {
        if (m.h_off) rt_host_free(m.h_off);
        m.h_off = (int64_t *)rt_host_alloc((size_t)(n_reads + 1) * 8); m.h_off_cap = m.h_off ? (size_t)(n_reads + 1) : 0;
        if (!m.h_off) return fail(e, C2B_E_CUDA, "c2b_align_batch: pinned allocation failed");
    }
    bool need_order = false;
    {
        const int64_t L0 = offsets[1] - offsets[0];
        for (int64_t k = 0; k <= n_reads; k++) m.h_off[k] = offsets[k] - b0;
        for (int64_t k = 1; k < n_reads && !need_order; k++)
            need_order = (offsets[k + 1] - offsets[k] != L0) || (ref_id && ref_id[k] != ref_id[0]);
    }
    // everything but the read bytes is small (8 + 12 bytes per read) and goes up before the launch
    RTCHK(rt_h2d(m.off.p, m.h_off, (size_t)(n_reads + 1) * 8, e->stream));
    if (count) RTCHK(rt_h2d(m.cnt.p, count, (size_t)n_reads * 4, e->stream));
    if (qweight) RTCHK(rt_h2d(m.qw.p, qweight, (size_t)n_reads * 4, e->stream));
    if (ref_id) RTCHK(rt_h2d(m.rid.p, ref_id, (size_t)n_reads * 4, e->stream));
    e->pair_order = nullptr;
    if (need_order) {                                      // per chunk: counting sort by (reference id, length); global indices
        if ((rc = ensure(e, m.ord, (size_t)n_reads * 4))) return rc;
        if (m.h_ord_cap < (size_t)n_reads) {
            if (m.h_ord) rt_host_free(m.h_ord);
            m.h_ord = (int32_t *)rt_host_alloc((size_t)n_reads * 4); m.h_ord_cap = m.h_ord ? (size_t)n_reads : 0;
            if (!m.h_ord) return fail(e, C2B_E_CUDA, "c2b_align_batch: pinned allocation failed");
        }
        const int64_t nb = (int64_t)(C2B_MAX_READ_LEN + 1) * (ref_id ? e->n_refs : 1);
        std::vector<int64_t> start((size_t)nb + 1);
        auto key = [&](int64_t k) -> int64_t {
            const int64_t L = offsets[k + 1] - offsets[k];
            const int64_t r = ref_id ? std::min<int64_t>(std::max<int32_t>(ref_id[k], 0), e->n_refs - 1) : 0;
            return r * (C2B_MAX_READ_LEN + 1) + L;
        };
        for (int c = 0; c < nc; c++) {
            std::fill(start.begin(), start.end(), 0);
            for (int64_t k = cuts[c]; k < cuts[c + 1]; k++) start[(size_t)key(k) + 1]++;
            for (int64_t b = 0; b < nb; b++) start[(size_t)b + 1] += start[(size_t)b];
            for (int64_t k = cuts[c]; k < cuts[c + 1]; k++) m.h_ord[cuts[c] + start[(size_t)key(k)]++] = (int32_t)k;
        }
        RTCHK(rt_h2d(m.ord.p, m.h_ord, (size_t)n_reads * 4, e->stream));
        e->pair_order = (const int32_t *)m.ord.p;
    }
    unsigned long long *h = m.h_ctl;
    for (size_t k = 0; k < ctl_n; k++) h[k] = 0;
    for (int c = 0; c < nc; c++) h[8 + c] = (unsigned long long)((cuts[c + 1] + 7) / 8);
    RTCHK(rt_h2d(m.ctl.p, h, ctl_n * 8, e->stream));
    unsigned long long *d_ctl = (unsigned long long *)m.ctl.p;
    // the copy stream starts after the control block, offsets and per-read arrays are in place (queued above on `stream`)
    RTCHK(rt_record(e->fork_ev, e->stream));
    RTCHK(rt_wait(e->s_in, e->fork_ev));
    // All copies are queued BEFORE the launch: with pinned host buffers they are asynchronous and overlap the kernel just the
    // same, and nothing the kernel waits for depends on host code that runs after the launch call -- a launch that blocks
    // the host (CUDA_LAUNCH_BLOCKING, a profiler serialising kernels) would otherwise leave the kernel waiting for
    // copies that are never issued.  (Pageable host buffers make cudaMemcpyAsync synchronous: correct, but the upload then
    // precedes the kernel instead of overlapping it -- use c2b_host_alloc.)
    unsigned long long *h_avail = h + ctl_n;               // pinned, one slot per chunk
    for (int c = 0; c < nc; c++) {
        const int64_t a = m.h_off[cuts[c]];
        int64_t b = m.h_off[cuts[c + 1]];
        if (c + 1 < nc) b = std::min<int64_t>((b + 127) & ~(int64_t)127, nbytes);    // whole 128-byte lines: no line is half-written when first read
        RTCHK(rt_h2d((uint8_t *)m.reads.p + a, reads + b0 + a, (size_t)(b - a), e->s_in));
        h_avail[c] = (unsigned long long)((cuts[c + 1] + 7) / 8);
        RTCHK(rt_h2d(d_ctl, &h_avail[c], 8, e->s_in));
    }
    // launch: work groups whose bytes have not arrived yet are waited for inside the kernel
    e->k_avail = d_ctl; e->k_chunk_end = d_ctl + 8; e->k_chunk_done = d_ctl + 8 + nc; e->k_n_chunks = nc;
    rc = launch_on(e, e->stream, 0, (const uint8_t *)m.reads.p, (const int64_t *)m.off.p, n_reads, (int32_t)maxJ,
                   count ? (const int32_t *)m.cnt.p : nullptr, qweight ? (const int32_t *)m.qw.p : nullptr,
                   ref_id ? (const int32_t *)m.rid.p : nullptr, (c2b_read_rec *)m.recs.p, (c2b_aln_rec *)m.alns.p,
                   strings ? (uint8_t *)m.str.p : nullptr, cap ? (c2b_edit *)m.ed.p : nullptr);
    e->k_avail = nullptr; e->k_chunk_end = nullptr; e->k_chunk_done = nullptr; e->k_n_chunks = 0;
    e->pair_order = nullptr;
    if (rc) return rc;
    // completion: poll the per-chunk counters, copy each chunk out as soon as it is whole
    unsigned long long *h_poll = h + ctl_n + nc + 8;        // third part of the pinned block
    for (int c = 0; c < nc; c++) {
        const unsigned long long want = (unsigned long long)((cuts[c + 1] + 7) / 8 - (cuts[c] + 7) / 8);
        int idle = 0;
        const auto t_wait = std::chrono::steady_clock::now();
        for (;;) {
            RTCHK(rt_d2h(h_poll, d_ctl, ctl_n * 8, m.s_poll));
            RTCHK(rt_sync(m.s_poll));
            if (h_poll[8 + nc + c] >= want) break;
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_wait).count() > 120.0)
                return fail(e, C2B_E_CUDA, "c2b_align_batch: streamed launch made no progress for 120 s");
            if (cudaStreamQuery(e->stream) == cudaSuccess && ++idle > 2)
                return fail(e, C2B_E_CUDA, "c2b_align_batch: streamed launch ended before all chunks were complete");
        }
        const int64_t c0 = cuts[c], n = cuts[c + 1] - cuts[c];
        RTCHK(rt_d2h(recs + c0, (c2b_read_rec *)m.recs.p + c0, (size_t)n * sizeof(c2b_read_rec), e->s_out));
        RTCHK(rt_d2h(alns + c0 * nr, (c2b_aln_rec *)m.alns.p + c0 * nr, (size_t)n * nr * sizeof(c2b_aln_rec), e->s_out));
        if (cap) RTCHK(rt_d2h(edits + c0 * nr * cap, (c2b_edit *)m.ed.p + c0 * nr * cap, (size_t)n * nr * cap * sizeof(c2b_edit), e->s_out));
        if (strings) {
            size_t Wt = ((size_t)h_poll[8 + 2 * nc + c] + 31) & ~(size_t)31;
            if (Wt > (size_t)W) Wt = W;
            RTCHK(rt_d2h_2d(strings + c0 * nr * 2 * W + (W - Wt), (const uint8_t *)m.str.p + c0 * nr * 2 * W + (W - Wt), W, Wt, (size_t)n * nr * 2, e->s_out));
        }
    }
    RTCHK(rt_sync(e->s_out));
    RTCHK(rt_sync(e->stream));
    {   // stats[7]: a wait inside the kernel timed out
        unsigned long long flag = 0;
        RTCHK(rt_d2h(&flag, (const char *)e->work.p + 7 * 8, 8, e->stream));
        RTCHK(rt_sync(e->stream));
        if (flag) return fail(e, C2B_E_CUDA, "c2b_align_batch: streamed launch timed out waiting for input");
    }
    return C2B_OK;
}
#endif

int c2b_align_batch(c2b_engine *e, const uint8_t *reads, const int64_t *offsets, int64_t n_reads,
                    const int32_t *count, const int32_t *qweight, const int32_t *ref_id,
                    c2b_read_rec *recs, c2b_aln_rec *alns, uint8_t *strings, c2b_edit *edits)
{
    if (!e || !e->configured) return fail(e, C2B_E_STATE, "c2b_align_batch: engine not configured");
    if (n_reads < 0 || !recs || !alns || (n_reads && (!reads || !offsets))) return fail(e, C2B_E_ARG, "c2b_align_batch: bad argument");
    if (n_reads == 0) return C2B_OK;
    int64_t maxJ = 1;
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t L = offsets[r + 1] - offsets[r];
        if (L < 0) return fail(e, C2B_E_ARG, "c2b_align_batch: offsets not monotone");
        maxJ = std::max(maxJ, L);
    }
    if (maxJ > C2B_MAX_READ_LEN) return fail(e, C2B_E_LIMIT, "c2b_align_batch: read longer than C2B_MAX_READ_LEN");
    if (!e->pipe_ready) {
        RTCHK(rt_stream_create(&e->s_in));
        RTCHK(rt_stream_create(&e->s_out));
        for (auto &st : e->stage) { RTCHK(rt_event_create(&st.in_done)); RTCHK(rt_event_create(&st.k_done)); RTCHK(rt_event_create(&st.out_done)); }
        e->pipe_ready = true;
    }
#ifndef C2B_EMU
    if (n_reads >= (1 << 16) && streamed_default(e)) {
        // one persistent launch per slice of up to 4 Mi reads (whole-slice device buffers: about 1.2 KB per read)
        const int64_t slice = 4 << 20;
        const int64_t Wb = (e->max_I + (int)maxJ + 31) & ~31, nrb = ref_id ? 1 : e->n_refs, capb = edits ? e->prm.edit_cap : 0;
        for (int64_t k = 0; k < n_reads; k += slice) {
            const int64_t n = std::min(slice, n_reads - k);
            const int rc2 = align_batch_streamed(e, reads, offsets + k, n, maxJ, count ? count + k : nullptr, qweight ? qweight + k : nullptr,
                                                 ref_id ? ref_id + k : nullptr, recs + k, alns + k * nrb,
                                                 strings ? strings + k * nrb * 2 * Wb : nullptr, edits ? edits + k * nrb * capb : nullptr);
            if (rc2) return rc2;
        }
        return C2B_OK;
    }
#endif
    const int W = (e->max_I + (int)maxJ + 31) & ~31;
    const int cap = edits ? e->prm.edit_cap : 0;
    const int nr = ref_id ? 1 : e->n_refs;                 // output slots per read (compact when ref_id is given)
    // Chunks pipeline through two staging sets: H2D of chunk c+1 and D2H of chunk c-1 overlap the kernel of chunk c.
    const int64_t per_read = (int64_t)nr * (2 * (int64_t)W * (strings ? 1 : 0) + (int64_t)cap * 8 + 32) + 16 + maxJ + 24;
    int64_t chunk = std::max<int64_t>(4096, std::min<int64_t>((int64_t)(768ll << 20) / per_read, 1 << 17));
    if (n_reads < 4 * chunk) chunk = std::max<int64_t>(4096, (n_reads + 3) / 4);
    if (const char *v = getenv("C2B_CHUNK")) chunk = std::max<int64_t>(2, atoll(v));     // test hook: force many small chunks
    for (auto &st : e->stage) st.used = false;
    int rc = C2B_OK;
    // D2H of a chunk is queued one iteration late: by then its kernel has finished and the widest alignment of the
    // chunk is known, so only the right-hand `Wt` bytes of every W-byte string slot cross PCIe.
    struct Pending { bool any = false; int64_t c0 = 0, n = 0; int set = 0; } pend;
    auto flush = [&](const Pending &q) -> int {
        c2b_engine::Stage &st = e->stage[q.set];
        RTCHK(rt_wait(e->s_out, st.k_done));
        RTCHK(rt_d2h(recs + q.c0, st.recs.p, (size_t)q.n * sizeof(c2b_read_rec), e->s_out));
        RTCHK(rt_d2h(alns + q.c0 * nr, st.alns.p, (size_t)q.n * nr * sizeof(c2b_aln_rec), e->s_out));
        if (cap) RTCHK(rt_d2h(edits + q.c0 * nr * cap, st.ed.p, (size_t)q.n * nr * cap * sizeof(c2b_edit), e->s_out));
        if (strings) {
            RTCHK(rt_event_sync(st.k_done));
            long long wmax = 0;
            RTCHK(rt_d2h(&wmax, (const char *)st.maxlen.p, 8, e->s_out));
            RTCHK(rt_sync(e->s_out));
            size_t Wt = ((size_t)wmax + 31) & ~(size_t)31;
            if (Wt > (size_t)W) Wt = W;
            RTCHK(rt_d2h_2d(strings + q.c0 * nr * 2 * W + (W - Wt), (const uint8_t *)st.str.p + (W - Wt), W, Wt, (size_t)q.n * nr * 2, e->s_out));
        }
        RTCHK(rt_record(st.out_done, e->s_out));
        return C2B_OK;
    };
    // chunk boundaries: a small first chunk (its H2D is exposed) and a small last chunk (its D2H is exposed)
    std::vector<int64_t> cuts;
    {
        const int64_t edge = getenv("C2B_CHUNK") ? std::max<int64_t>(2, chunk / 2) : std::max<int64_t>(4096, chunk / 8);
        int64_t pos = 0;
        cuts.push_back(0);
        if (n_reads > 4 * edge) { pos = edge; cuts.push_back(pos); }
        const int64_t tail = (n_reads - pos > 2 * edge) ? edge : 0;
        while (n_reads - tail - pos > 0) { pos += std::min(chunk, n_reads - tail - pos); cuts.push_back(pos); }
        if (tail) cuts.push_back(n_reads);
    }
    // Optional (C2B_TWO_STREAMS=1): kernels of consecutive chunks on two compute streams.  Measured r01k: 42.7 ms per
    // 1 M reads against 35.7 ms on one stream -- two resident persistent grids slow each other more than the
    // launch tails cost -- so one stream is the default.
    bool two_streams = false;
#ifndef C2B_EMU
    two_streams = e->stream2 && getenv("C2B_TWO_STREAMS");
#endif
    if ((rc = ensure_scratch(e, (int)maxJ))) return rc;       // allocations / memsets on `stream` happen before the fork
    if (two_streams) { RTCHK(rt_record(e->fork_ev, e->stream)); RTCHK(rt_wait(e->stream2, e->fork_ev)); }
    for (int ci = 0; ci + 1 < (int)cuts.size(); ci++) {
        c2b_engine::Stage &st = e->stage[ci & 1];
        const int64_t c0 = cuts[ci], n = cuts[ci + 1] - cuts[ci];
        const int64_t b0 = offsets[c0], b1 = offsets[c0 + n];
        if (st.used) RTCHK(rt_event_sync(st.out_done));        // set is being reused: the D2H of chunk ci-2 must be done
        if ((rc = ensure(e, st.reads, (size_t)(b1 - b0) + 16))) return rc;
        if ((rc = ensure(e, st.off, (size_t)(n + 1) * 8))) return rc;
        if ((rc = ensure(e, st.recs, (size_t)n * sizeof(c2b_read_rec)))) return rc;
        if ((rc = ensure(e, st.alns, (size_t)n * nr * sizeof(c2b_aln_rec)))) return rc;
        if ((rc = ensure(e, st.maxlen, 8))) return rc;
        if (strings && (rc = ensure(e, st.str, (size_t)n * nr * 2 * W))) return rc;
        if (cap && (rc = ensure(e, st.ed, (size_t)n * nr * cap * sizeof(c2b_edit)))) return rc;
        if (count && (rc = ensure(e, st.cnt, (size_t)n * 4))) return rc;
        if (qweight && (rc = ensure(e, st.qw, (size_t)n * 4))) return rc;
        if (ref_id && (rc = ensure(e, st.rid, (size_t)n * 4))) return rc;
        if (st.h_off_cap < (size_t)(n + 1)) {
            if (st.h_off) rt_host_free(st.h_off);
            st.h_off = (int64_t *)rt_host_alloc((size_t)(n + 1) * 8); st.h_off_cap = st.h_off ? (size_t)(n + 1) : 0;
            if (!st.h_off) return fail(e, C2B_E_CUDA, "c2b_align_batch: pinned allocation failed");
        }
        for (int64_t k = 0; k <= n; k++) st.h_off[k] = offsets[c0 + k] - b0;      // chunk-relative offsets
        RTCHK(rt_h2d(st.reads.p, reads + b0, (size_t)(b1 - b0), e->s_in));
        RTCHK(rt_h2d(st.off.p, st.h_off, (size_t)(n + 1) * 8, e->s_in));
        if (count) RTCHK(rt_h2d(st.cnt.p, count + c0, (size_t)n * 4, e->s_in));
        if (qweight) RTCHK(rt_h2d(st.qw.p, qweight + c0, (size_t)n * 4, e->s_in));
        if (ref_id) RTCHK(rt_h2d(st.rid.p, ref_id + c0, (size_t)n * 4, e->s_in));
        // pairing order: counting sort of the chunk by (reference id, length) when reads differ, so equal ones are adjacent
        bool need_order = false;
        for (int64_t k = 1; k < n && !need_order; k++)
            need_order = (st.h_off[k + 1] - st.h_off[k] != st.h_off[1] - st.h_off[0]) || (ref_id && ref_id[c0 + k] != ref_id[c0]);
        e->pair_order = nullptr;
        if (need_order) {
            if ((rc = ensure(e, st.ord, (size_t)n * 4))) return rc;
            if (st.h_ord_cap < (size_t)n) {
                if (st.h_ord) rt_host_free(st.h_ord);
                st.h_ord = (int32_t *)rt_host_alloc((size_t)n * 4); st.h_ord_cap = st.h_ord ? (size_t)n : 0;
                if (!st.h_ord) return fail(e, C2B_E_CUDA, "c2b_align_batch: pinned allocation failed");
            }
            const int64_t nb = (int64_t)(C2B_MAX_READ_LEN + 1) * (ref_id ? e->n_refs : 1);
            std::vector<int64_t> start((size_t)nb + 1, 0);
            auto key = [&](int64_t k) -> int64_t {
                const int64_t L = st.h_off[k + 1] - st.h_off[k];
                const int64_t r = ref_id ? std::min<int64_t>(std::max<int32_t>(ref_id[c0 + k], 0), e->n_refs - 1) : 0;
                return r * (C2B_MAX_READ_LEN + 1) + L;
            };
            for (int64_t k = 0; k < n; k++) start[(size_t)key(k) + 1]++;
            for (int64_t b = 0; b < nb; b++) start[(size_t)b + 1] += start[(size_t)b];
            for (int64_t k = 0; k < n; k++) st.h_ord[start[(size_t)key(k)]++] = (int32_t)k;
            RTCHK(rt_h2d(st.ord.p, st.h_ord, (size_t)n * 4, e->s_in));
            e->pair_order = (const int32_t *)st.ord.p;
        }
        RTCHK(rt_record(st.in_done, e->s_in));
        // kernels of consecutive chunks go to two compute streams with their own scratch sets: chunk ci+1's CTAs fill
        // the SMs as chunk ci's CTAs run out of work, instead of waiting for its slowest warp (same-stream launches
        // of one chunk parity stay ordered, so at most two launches -- of different sets -- are ever resident)
        const int set = two_streams ? (ci & 1) : 0;
        rt_stream cs = set ? e->stream2 : e->stream;
        RTCHK(rt_wait(cs, st.in_done));
        rc = launch_on(e, cs, set, (const uint8_t *)st.reads.p, (const int64_t *)st.off.p, n, (int32_t)maxJ,
                                    count ? (const int32_t *)st.cnt.p : nullptr, qweight ? (const int32_t *)st.qw.p : nullptr,
                                    ref_id ? (const int32_t *)st.rid.p : nullptr, (c2b_read_rec *)st.recs.p,
                                    (c2b_aln_rec *)st.alns.p, strings ? (uint8_t *)st.str.p : nullptr,
                                    cap ? (c2b_edit *)st.ed.p : nullptr);
        e->pair_order = nullptr;
        if (rc) return rc;
        // keep this launch's "widest alignment" before the next launch resets it
        RTCHK(cudaMemcpyAsyncOrCopy(st.maxlen.p, (const char *)e->work.p + (9 + 8 * set) * 8, 8, cs));
        RTCHK(rt_record(st.k_done, cs));
        st.used = true;
        if (pend.any && (rc = flush(pend))) return rc;           // chunk ci-1: overlaps this chunk's kernel
        pend.any = true; pend.c0 = c0; pend.n = n; pend.set = ci & 1;
    }
    if (pend.any && (rc = flush(pend))) return rc;
    RTCHK(rt_sync(e->s_out));
    RTCHK(rt_sync(e->stream));
    if (two_streams) RTCHK(rt_sync(e->stream2));
    return C2B_OK;
}

int c2b_counts_layout(const c2b_engine *e, int32_t *n_refs, int32_t *n_vec, int32_t *stride, int32_t *n_scal)
{
    if (!e || !e->configured) return C2B_E_STATE;
    if (n_refs) *n_refs = e->n_refs;
    if (n_vec) *n_vec = C2B_NVEC;
    if (stride) *stride = e->vstride;
    if (n_scal) *n_scal = C2B_NSCAL;
    return C2B_OK;
}

int c2b_counts_hist_layout(const c2b_engine *e, int32_t *n_hist, int32_t *hstride, int32_t *hist_zero)
{
    if (!e || !e->configured) return C2B_E_STATE;
    if (n_hist) *n_hist = C2B_NHIST;
    if (hstride) *hstride = e->hstride;
    if (hist_zero) *hist_zero = e->hist_zero;
    return C2B_OK;
}

int c2b_counts_reset(c2b_engine *e)
{
    if (!e || !e->configured) return fail(e, C2B_E_STATE, "c2b_counts_reset: engine not configured");
    RTCHK(rt_zero(e->d_counts, e->counts_n * 8, e->stream));
    if (e->work.p) RTCHK(rt_zero(e->work.p, WORK_BYTES, e->stream));
    RTCHK(rt_sync(e->stream));
    return C2B_OK;
}

int c2b_counts_read(c2b_engine *e, int64_t *out, size_t n_int64)
{
    if (!e || !e->configured || !out) return fail(e, C2B_E_STATE, "c2b_counts_read: engine not configured");
    if (n_int64 < e->counts_n) return fail(e, C2B_E_ARG, "c2b_counts_read: buffer too small");
    RTCHK(rt_d2h(out, e->d_counts, e->counts_n * 8, e->stream));
    RTCHK(rt_sync(e->stream));
    return C2B_OK;
}

int c2b_counts_device(c2b_engine *e, void **d_ptr, size_t *n_int64)
{
    if (!e || !e->configured) return C2B_E_STATE;
    if (d_ptr) *d_ptr = e->d_counts;
    if (n_int64) *n_int64 = e->counts_n;
    return C2B_OK;
}

int c2b_global_align(c2b_engine *e, const char *read, int32_t read_len, const char *ref, int32_t ref_len,
                     const char *alphabet, int32_t nq, const int64_t *score_rows, const int64_t *gap_incentive,
                     int32_t gap_open, int32_t gap_extend,
                     char *out_read, char *out_ref, int32_t *aln_len, int32_t *n_match)
{
    if (!e || !read || !ref || !alphabet || !score_rows || !gap_incentive || !out_read || !out_ref || !aln_len || !n_match)
        return fail(e, C2B_E_ARG, "c2b_global_align: bad argument");
    c2b_params p; memset(&p, 0, sizeof p);
    p.gap_open = gap_open; p.gap_extend = gap_extend; p.flags = C2B_F_NO_STRAND_SEARCH; p.nq = nq;
    if (nq < 1 || nq > C2B_MAX_Q) return fail(e, C2B_E_LIMIT, "c2b_global_align: alphabet larger than C2B_MAX_Q");
    memcpy(p.alphabet, alphabet, nq);
    for (int q = 0; q < nq; q++) p.complement[q] = (uint8_t)q;
    c2b_ref r; memset(&r, 0, sizeof r);
    r.seq = ref; r.len = ref_len; r.gap_incentive = gap_incentive; r.score_rows = score_rows; r.min_aln_score = -1.0;
    int rc = c2b_configure(e, &p, 1, &r);
    if (rc) return rc;
    const int W = c2b_string_width(e, read_len);
    std::vector<uint8_t> str((size_t)2 * W);
    int64_t off[2] = {0, read_len};
    c2b_read_rec rec; c2b_aln_rec a;
    rc = c2b_align_batch(e, (const uint8_t *)read, off, 1, nullptr, nullptr, nullptr, &rec, &a, str.data(), nullptr);
    if (rc) return rc;
    if (a.status) { e->err = "c2b_global_align: alignment status " + std::to_string(a.status); return 100 + a.status; }
    memcpy(out_read, str.data() + W - a.aln_len, a.aln_len);
    memcpy(out_ref, str.data() + 2 * W - a.aln_len, a.aln_len);
    *aln_len = a.aln_len; *n_match = a.n_match;
    return C2B_OK;
}

int c2b_classify_aligned(c2b_engine *e, const char *read_al, const char *ref_al, int32_t n_cols,
                         const char *alphabet, int32_t nq, const int64_t *include_idx, int32_t n_include,
                         c2b_aln_rec *out, c2b_edit *edits)
{
    if (!e || !read_al || !ref_al || !alphabet || !out || !edits || n_cols < 1) return fail(e, C2B_E_ARG, "c2b_classify_aligned: bad argument");
    if (n_cols > C2B_MAX_ALN_LEN) return fail(e, C2B_E_LIMIT, "c2b_classify_aligned: alignment longer than C2B_MAX_ALN_LEN");
    if (nq < 1 || nq > C2B_MAX_Q) return fail(e, C2B_E_LIMIT, "c2b_classify_aligned: alphabet larger than C2B_MAX_Q");
    std::string read, ref;
    std::vector<uint64_t> ops(32, ~0ull);
    int prev = -1;
    for (int c = n_cols - 1, n = 0; c >= 0; c--, n++) {           // op n = n-th column from the right
        const bool gq = read_al[c] == '-', gr = ref_al[c] == '-';
        if (gq && gr) return fail(e, C2B_E_ARG, "c2b_classify_aligned: column with two gaps");
        const int op = gq ? OP_J : gr ? OP_I : OP_M;
        if ((op == OP_I && prev == OP_J) || (op == OP_J && prev == OP_I))
            return fail(e, C2B_E_ARG, "c2b_classify_aligned: insertion column adjacent to a deletion column");
        prev = op;
        ops[n >> 5] &= ~(3ull << (2 * (n & 31)));
        ops[n >> 5] |= (uint64_t)op << (2 * (n & 31));
    }
    for (int c = 0; c < n_cols; c++) { if (read_al[c] != '-') read.push_back(read_al[c]); if (ref_al[c] != '-') ref.push_back(ref_al[c]); }
    if (read.empty() || ref.empty()) return fail(e, C2B_E_ARG, "c2b_classify_aligned: empty sequence");
    if ((int)read.size() > C2B_MAX_READ_LEN || (int)ref.size() > C2B_MAX_REF_LEN) return fail(e, C2B_E_LIMIT, "c2b_classify_aligned: sequence too long");
    c2b_params p; memset(&p, 0, sizeof p);
    p.gap_open = -1; p.gap_extend = -1; p.flags = C2B_F_NO_STRAND_SEARCH; p.nq = nq; p.edit_cap = n_cols + 1;
    memcpy(p.alphabet, alphabet, nq);
    for (int q = 0; q < nq; q++) p.complement[q] = (uint8_t)q;
    std::vector<int64_t> gi(ref.size() + 1, 0), rows((size_t)nq * ref.size(), 0);
    c2b_ref r; memset(&r, 0, sizeof r);
    r.seq = ref.c_str(); r.len = (int32_t)ref.size(); r.gap_incentive = gi.data(); r.score_rows = rows.data();
    r.include_idx = include_idx; r.n_include = n_include; r.min_aln_score = -1.0;
    int rc = c2b_configure(e, &p, 1, &r);
    if (rc) return rc;
    DevBuf d_ops, d_n;
    if ((rc = ensure(e, d_ops, 32 * 8)) || (rc = ensure(e, d_n, 4))) return rc;
    const int32_t n32 = n_cols;
    RTCHK(rt_h2d(d_ops.p, ops.data(), 32 * 8, e->stream));
    RTCHK(rt_h2d(d_n.p, &n32, 4, e->stream));
    e->forced_ops = (const uint64_t *)d_ops.p; e->forced_n = (const int32_t *)d_n.p;
    int64_t off[2] = {0, (int64_t)read.size()};
    c2b_read_rec rec;
    rc = c2b_align_batch(e, (const uint8_t *)read.data(), off, 1, nullptr, nullptr, nullptr, &rec, out, nullptr, edits);
    e->forced_ops = nullptr; e->forced_n = nullptr;
    rt_free(d_ops.p); rt_free(d_n.p);
    return rc;
}

// pinned host memory for callers that want full-speed copies (bench.py, the Python wrapper)
void *c2b_host_alloc(size_t n)
{
#ifndef C2B_EMU
    void *p = nullptr;
    if (cudaHostAlloc(&p, n ? n : 16, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
#else
    return malloc(n ? n : 16);
#endif
}
void c2b_host_free(void *p)
{
#ifndef C2B_EMU
    if (p) cudaFreeHost(p);
#else
    free(p);
#endif
}

}  // extern "C"
