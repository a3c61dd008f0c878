// phb200 internal helpers: error handling, per-system scalar records, deterministic fused reductions.
// sm_100a only.  See include/phb200.h for the public C ABI and DESIGN.md for the kernel inventory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include "../../include/phb200.h"

namespace phb {

// ------------------------------------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

// Temporaries of the set-up calls come from the device's stream-ordered pool.  By default that pool hands its memory back to
// the OS at every synchronisation, so each call paid fresh allocations (milliseconds on a process holding GBs); the release
// threshold is raised once per device so the temporaries are recycled.
inline cudaError_t phb_malloc_async(void** p, size_t bytes, cudaStream_t stream) {
    static std::atomic<unsigned> configured{0};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 32 && !(configured.load(std::memory_order_relaxed) & (1u << dev))) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t threshold = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        }
        cudaGetLastError();
        configured.fetch_or(1u << dev, std::memory_order_relaxed);
    }
    return cudaMallocAsync(p, bytes, stream);
}
template <typename P>
inline cudaError_t phb_malloc_async(P** p, size_t bytes, cudaStream_t stream) { return phb_malloc_async((void**)p, bytes, stream); }

#define PHB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            phb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PHB200_ERR_CUDA;                                                             \
        }                                                                                       \
    } while (0)

#define PHB_ARG(cond, ...)                                                                      \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            phb::set_error(__VA_ARGS__);                                                        \
            return PHB200_ERR_ARG;                                                              \
        }                                                                                       \
    } while (0)

#define PHB_TRY(expr)                                                                           \
    do {                                                                                        \
        int _r = (expr);                                                                        \
        if (_r != PHB200_OK) return _r;                                                         \
    } while (0)

// launch + count + check.  PHB200_DEBUG_SYNC=1 synchronises after every launch and names the kernel that faulted.
bool debug_sync();
#define PHB_LAUNCH(kernel, grid, block, smem, stream, ...)                                      \
    do {                                                                                        \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                             \
        phb::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        cudaError_t _e = cudaPeekAtLastError();                                                 \
        if (_e == cudaSuccess && phb::debug_sync()) _e = cudaStreamSynchronize(stream);         \
        if (_e != cudaSuccess) {                                                                \
            phb::set_error("kernel %s failed: %s (%s:%d)", #kernel, cudaGetErrorString(_e), __FILE__, __LINE__); \
            if (phb::debug_sync()) fprintf(stderr, "[phb200] kernel %s failed: %s (%s:%d)\n", #kernel, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PHB200_ERR_CUDA;                                                             \
        }                                                                                       \
    } while (0)

// Launch with programmatic dependent launch (PDL): the grid may become resident while the previous kernel of the stream is
// still draining; the kernel orders itself behind it with griddepcontrol.wait (pdl_wait) before touching its results.
bool pdl_enabled();
// Optional L2 residency hint for the launches that go through PHB_LAUNCH_PDL (the two kernels of the fused CG iteration):
// accesses inside [base, base + bytes) are marked persisting with probability `hit`, everything else streaming
// (cudaLaunchAttributeAccessPolicyWindow).  base == nullptr: no hint.  Set by the solver when PHB200_L2_PERSIST=1 (experimental,
// off by default: not measured yet); the window is q, which the stencil kernel writes and the residual kernel reads next.
struct L2Window {
    void* base;
    size_t bytes;
    float hit;
};
L2Window& l2_window();
L2Window l2_persist_window(void* base, size_t bytes);
int l2_persist_mode();   // 0 off, 1 window on q, 2 window on r
#define PHB_LAUNCH_PDL(kernel, grid, block, smem, strm, ...)                                  \
    do {                                                                                        \
        cudaLaunchConfig_t _cfg = {};                                                           \
        _cfg.gridDim = (grid);                                                                  \
        _cfg.blockDim = (block);                                                                \
        _cfg.dynamicSmemBytes = (smem);                                                         \
        _cfg.stream = (strm);                                                                   \
        cudaLaunchAttribute _at[2];                                                             \
        unsigned _na = 0;                                                                       \
        if (phb::pdl_enabled()) {                                                               \
            _at[_na].id = cudaLaunchAttributeProgrammaticStreamSerialization;                   \
            _at[_na].val.programmaticStreamSerializationAllowed = 1;                            \
            ++_na;                                                                              \
        }                                                                                       \
        const phb::L2Window& _win = phb::l2_window();                                           \
        if (_win.base) {                                                                        \
            _at[_na].id = cudaLaunchAttributeAccessPolicyWindow;                                \
            _at[_na].val.accessPolicyWindow.base_ptr = _win.base;                               \
            _at[_na].val.accessPolicyWindow.num_bytes = _win.bytes;                             \
            _at[_na].val.accessPolicyWindow.hitRatio = _win.hit;                                \
            _at[_na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;             \
            _at[_na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;             \
            ++_na;                                                                              \
        }                                                                                       \
        _cfg.attrs = _at;                                                                       \
        _cfg.numAttrs = _na;                                                                    \
        cudaError_t _e = cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__);                        \
        phb::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        if (_e == cudaSuccess && phb::debug_sync()) _e = cudaStreamSynchronize(strm);           \
        if (_e != cudaSuccess) {                                                                \
            phb::set_error("kernel %s failed: %s (%s:%d)", #kernel, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PHB200_ERR_CUDA;                                                             \
        }                                                                                       \
    } while (0)

// cudaFuncSetAttribute and occupancy queries are PER DEVICE: "done once" flags of the launchers are kept per device (a process may hold
// operands on several GPUs; one process per GPU is the common case).
constexpr int kMaxDevices = 32;
inline int device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
    return d;
}

constexpr int kSMs = 148;          // B200
constexpr int kBlock = 256;        // threads per CTA for all streaming kernels
constexpr int kMaxDots = 4;        // fused dot products per kernel
constexpr int kMaxStencil = 16;
constexpr int kTargetCtas = kSMs * 8;   // upper bound on CTAs of one grid-stride launch (sizes the reduction scratch)

// ------------------------------------------------------------------------------------------------ records

// One record per system (batch entry).  Scalars are kept as doubles that hold values already rounded to the
// arithmetic type T where the reference computes them in T.
struct Sys {
    double s[40];      // algorithm-specific scalar slots (alpha, beta, delta, rho, omega, tau[], sigma[], gamma[] ...)
    double tol;        // this system's squared tolerance before the batch minimum
    double rsq0;       // first monitored value seen by the progress check
    double rsq;        // monitored value of the current pass (|r|^2 or r.M^-1 r)
    int its, fev, go, conv, div, max_iter, pad0, pad1;
};

constexpr int kSlotYY = 11;   // Sys::s slot holding |y_b|^2 (S_YY in solver.cu)

struct Glob {
    double tol_min;    // minimum squared tolerance over the batch (the (batch,batch) broadcast of stop_on_l2)
    double yy_sum;     // sum over the batch of |y_b|^2 (generic cg_adaptive / bicg: the reference's tolerance is relative to it)
    int any_go;        // 1 while at least one system iterates
    int checks;        // progress checks performed so far (0 -> first check semantics)
    int pass;          // loop passes executed (torch fast path: true-residual refresh every 20th pass)
    int refresh;       // 1 if this pass refreshes the residual
    unsigned ticket;   // global "last block" counter
    int pad[3];
};

// All-reduce over the slab ranks INSIDE the kernel that produced the local sums (one process per GPU, peer memory mapped through
// CUDA IPC): the finishing thread of a system stores its sums into every rank's mailbox (NVLink peer stores), waits until the
// mailboxes of its own GPU hold the sums of all ranks for this reduction, and adds them in rank order — the same value on every
// rank — before the ordinary scalar step.  The messages are SELF-VALIDATING (the "LL" scheme of collective libraries): every
// 8-byte word carries 32 bits of payload and the 32-bit sequence number of the reduction, and an 8-byte store is single-copy
// atomic, so no fence or release/acquire pair orders payload against flag: all words of all peers leave back to back (one NVLink
// latency in total), the receiver polls until every word shows the expected sequence number.  Two slots per (rank, system) are
// enough: nobody can be two reductions ahead of a peer.
struct PeerMail {
    unsigned long long w[2 * kMaxDots];   // w[2d] = lo32(v_d) << 32 | seq32, w[2d+1] = hi32(v_d) << 32 | seq32
};
struct PeerReduce {
    PeerMail* mail[8];             // mailbox array of every rank: [2][world][batch]
    unsigned long long* seq;       // [batch] reductions made so far (this GPU)
    int world, rank, batch;
};

struct Reduce {
    double* partials;      // [batch][max_blocks][kMaxDots]
    unsigned* tickets;     // [batch]
    int max_blocks;
    const PeerReduce* peer;   // non-null: multi-GPU slab mode with in-kernel all-reduce
    int peer_data;            // this kernel also stores data (halo planes) into peer memory: fences at system scope around the all-reduce
};

// ------------------------------------------------------------------------------------------------ device helpers

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

// separately rounded multiply / add (never contracted into an FMA): keeps row sums bit-identical to scipy's csr_matvec
// and the vector updates on the reference's rounding sequence
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename T> __device__ __forceinline__ T safe_div(T a, T b) { return b == T(0) ? T(0) : a / b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum `ND` per-thread accumulators over the CTA; result valid in thread 0.  `sh` needs ND*8 doubles (256 threads).
template <int ND>
__device__ __forceinline__ void block_sum(double (&acc)[ND], double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        double v = warp_sum(acc[d]);
        if (lane == 0) sh[d * 8 + warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            double v = lane < (int)(blockDim.x >> 5) ? sh[d * 8 + lane] : 0.0;
            v = warp_sum(v);
            if (lane == 0) acc[d] = v;
        }
    }
    __syncthreads();
}

// Deterministic cross-CTA reduction for system `b`.  Thread 0 of every CTA holds that CTA's sums in `tot`; every CTA
// deposits them, the CTA that draws the last ticket adds all partials in a fixed order.  Returns true (in every
// thread of that CTA) for the last CTA; `tot` then holds the totals in thread 0.  Tickets reset themselves, so
// back-to-back kernels on one stream can share `red`.  nbx == 1 needs neither scratch nor tickets.
template <int ND>
__device__ __forceinline__ bool system_reduce(double (&tot)[ND], const Reduce& red, int b, int bx, int nbx) {
    if (nbx == 1) return true;
    __shared__ double sh[kMaxDots * 8];
    __shared__ int s_last;
    double* mine = red.partials + ((size_t)b * red.max_blocks) * kMaxDots;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) mine[(size_t)bx * kMaxDots + d] = tot[d];
        __threadfence();     // (stores into peer GPUs were fenced at system scope by the threads that made them, phase.cuh)
        unsigned t = atomicAdd(&red.tickets[b], 1u);
        s_last = (t == (unsigned)nbx - 1u);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    double part[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) part[d] = 0.0;
    for (int i = threadIdx.x; i < nbx; i += blockDim.x) {
#pragma unroll
        for (int d = 0; d < ND; ++d) part[d] += __ldcg(&mine[(size_t)i * kMaxDots + d]);
    }
    block_sum<ND>(part, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) tot[d] = part[d];
        red.tickets[b] = 0u;
    }
    return true;
}

// thread 0 of the finishing CTA of system b: tot (local sums) -> sums over all ranks.  `ordered_data`: this kernel also stored
// DATA into peer memory (halo planes) that the peers read after the reduction: one system-scope fence orders it (cumulatively:
// the other CTAs fenced before drawing their tickets) before the messages, and the receiver fences after the last poll.
template <int ND>
__device__ __forceinline__ void peer_allreduce(const PeerReduce& pr, int b, double (&tot)[ND], bool ordered_data = true) {
    const unsigned long long s = pr.seq[b] + 1ull;
    pr.seq[b] = s;
    const unsigned long long tag = s & 0xffffffffull;
    const size_t par = (size_t)(s & 1ull);
    const size_t mine = (par * pr.world + pr.rank) * pr.batch + b;
    unsigned long long word[2 * ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(tot[d]);
        word[2 * d] = ((bits & 0xffffffffull) << 32) | tag;
        word[2 * d + 1] = (bits & 0xffffffff00000000ull) | tag;
    }
    if (ordered_data) __threadfence_system();
    for (int p = 0; p < pr.world; ++p) {
        unsigned long long* m = (pr.mail[p] + mine)->w;
#pragma unroll
        for (int k = 0; k < 2 * ND; ++k) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(m + k), "l"(word[k]) : "memory");
    }
    double sum[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) sum[d] = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < pr.world; ++r) {
        const unsigned long long* m = (pr.mail[pr.rank] + (par * pr.world + r) * pr.batch + b)->w;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            unsigned long long lo = 0, hi = 0;
            while (true) {
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(lo) : "l"(m + 2 * d) : "memory");
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(hi) : "l"(m + 2 * d + 1) : "memory");
                if ((lo & 0xffffffffull) == tag && (hi & 0xffffffffull) == tag) break;
                if (clock64() - t0 > 8000000000ll) __trap();
            }
            sum[d] += __longlong_as_double((long long)((hi & 0xffffffff00000000ull) | (lo >> 32)));
        }
    }
    if (ordered_data) __threadfence_system();
#pragma unroll
    for (int d = 0; d < ND; ++d) tot[d] = sum[d];
}

// Per-thread accumulators -> CTA sum -> system_reduce.
template <int ND>
__device__ __forceinline__ bool reduce_system(double (&acc)[ND], const Reduce& red, int b, int bx, int nbx) {
    __shared__ double sh[kMaxDots * 8];
    block_sum<ND>(acc, sh);
    return system_reduce<ND>(acc, red, b, bx, nbx);
}

// After the per-system finaliser ran in thread 0 of a system's last CTA: draw the global ticket.  Returns true in
// every thread of the one CTA that finished last over the whole batch.
__device__ __forceinline__ bool last_of_batch(Glob* g, int batch) {
    __shared__ int s_glast;
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicAdd(&g->ticket, 1u);
        s_glast = (t == (unsigned)batch - 1u);
        if (s_glast) g->ticket = 0u;
    }
    __syncthreads();
    if (s_glast) __threadfence();
    return s_glast != 0;
}

// Progress check of stop_on_l2 (phiml/backend/_linalg.py:23-40) for every system; run by ALL threads of one CTA.
// torch_rules: torch_sparse_cg's variant (threshold 100, no non-finite test inside the loop).
template <typename T>
__device__ __forceinline__ void progress_check(Sys* sys, Glob* g, int batch, bool torch_rules, bool freeze_finished) {
    const bool first = g->checks == 0;
    const T tol = (T)g->tol_min;
    int any = 0;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        volatile Sys* s = sys + b;
        if (freeze_finished && !first && !s->go) continue;   // finished systems are frozen: same inputs, same flags
        T rsq = (T)fabs(s->rsq);
        bool conv = rsq <= tol;
        bool nonfin = !isfinite(rsq);
        bool div;
        if (first) {
            s->rsq0 = (double)rsq;
            // generic loops: non-finite monitored value (stop_on_l2).  torch_sparse_cg(_adaptive) test x0 instead — any(~isfinite(x0)), _torch_backend.py:1362 —
            // so a non-finite right-hand side does NOT stop the system (it iterates on NaNs up to max_iter); |r0|^2 non-finite while |y|^2 is finite = x0's doing
            div = torch_rules ? (nonfin && isfinite(s->s[kSlotYY])) : nonfin;
        } else {
            T ratio = rsq / (T)s->rsq0;
            if (torch_rules) div = (ratio > T(100)) && s->its >= 8;
            else div = ((ratio > T(1e5)) && s->its >= 8) || nonfin;
        }
        int go = (!conv && !div && s->its < s->max_iter) ? 1 : 0;
        s->conv = conv;
        s->div = div;
        s->go = go;
        any |= go;
    }
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    if (any) atomicOr(&s_any, 1);
    __syncthreads();
    any = s_any;
    if (threadIdx.x == 0) {
        g->any_go = any;
        g->checks = g->checks + 1;
        __threadfence();
    }
}

template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }

// PDL: wait until every kernel launched before this one on the stream has completed and its writes are visible (a no-op
// for an ordinary launch); then let the NEXT kernel of the stream start getting resident.
// L2 eviction-priority policies (createpolicy) and 16-byte global accesses that carry one.  kind: 0 plain, 1 evict_last (keep), 2 evict_first
// (stream).  Used by the fused CG pair to keep the residual in the 126 MB L2 between its three accesses per iteration.
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ld16_hint(const void* ptr, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void st16_hint(void* ptr, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

}  // namespace phb
