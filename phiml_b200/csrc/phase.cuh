// Element-wise "phase" kernel: one pass over up to 8 vectors of a batched system with fused dot products, the
// per-system scalar update and (optionally) the batch-wide progress check in the last CTA.  128-bit accesses.
#pragma once
#include "spmv_kernels.cuh"

#include <utility>
#include <type_traits>

namespace phb {

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// Optional traversal order of a phase kernel.  plane == 0: plain index order.  Otherwise the vector is seen as x-planes of
// `plane` elements grouped into chunks of `xchunk` planes (the decomposition of the plane-marching stencil kernel that ran
// just before), and every chunk is walked from its LAST plane to its first, all chunks side by side: the kernel first touches
// what the stencil kernel touched last (still in the 126 MB L2) and leaves the chunk heads — where the next stencil pass
// starts — for the end.
struct PhaseOrder {
    int64_t plane;
    int xchunk, chunks;
    int64_t planes;
    unsigned keep_mask = 0, stream_mask = 0;   // vectors whose 16-byte accesses carry an L2 evict_last / evict_first policy
};

// Slab mode, fused halo exchange: a phase functor that defines `peer` (PeerHalo) makes k_phase ALSO store array 0 on the first /
// last owned plane into the slab neighbours' halo planes (their memory, mapped through CUDA IPC: plain coalesced stores over
// NVLink), so no separate exchange step follows the kernel.
struct PeerHalo {
    void* lo;          // neighbour below: its upper halo plane of the same vector (null: physical boundary / not a slab)
    void* hi;          // neighbour above: its lower halo plane
    int64_t plane;     // elements per plane
    int64_t ld_lo, ld_hi;   // leading dimensions (elements per system) of the neighbours' arrays
};
// a phase functor that defines `per_index` gets the element's index: elem(e, acc, sc, i) (per-row table look-ups, e.g. the inverse diagonal
// of a pattern-coded star through the row's code byte)
template <class P, class = void> struct HasIndex : std::false_type {};
template <class P> struct HasIndex<P, std::void_t<decltype(P::per_index)>> : std::true_type {};
template <class P, class = void> struct HasPeer : std::false_type {};
template <class P> struct HasPeer<P, std::void_t<decltype(std::declval<P>().peer)>> : std::true_type {};

struct Ctx {
    Sys* sys;
    Glob* glob;
    Reduce red;
    int batch;
    double* defer;   // non-null: store the per-system sums here instead of finalising (multi-GPU all-reduce in between)
};

// Phase protocol (in addition to the epilogue protocol of spmv_kernels.cuh)
//   NA                 number of vectors; v[a] base pointers (row b at v[a] + b*ld unless bit a of SHARED is set)
//   RMASK / WMASK      which vectors are loaded / stored
//   Scal load(b)       per-system scalars, read once per thread
//   elem(e, acc, sc)   one element: e[a] in/out
template <typename T, class P, int V>
__global__ void __launch_bounds__(kBlock) k_phase(P p, int64_t n, int64_t ld, int64_t off, PhaseOrder ord) {
    pdl_wait();
    if (p.idle()) return;
    const int b = blockIdx.y;
    constexpr int ND = AccSize<P>::N;
    constexpr int NA = P::NA;
    double acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[d] = 0.0;
    bool wrote_peer = false;      // this thread stored into a slab neighbour's memory
    if (!p.skip(b)) {
        const typename P::Scal sc = p.load(b);
        T* base[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) base[a] = (T*)p.v[a] + off + (((P::SHARED >> a) & 1u) ? (int64_t)0 : (int64_t)b * ld);
        using VT = typename Vec<T>::type;
        const bool hinted = (ord.keep_mask | ord.stream_mask) != 0 && sizeof(VT) == 16;
        const uint64_t pol_keep = hinted ? l2_policy_keep() : 0, pol_stream = hinted ? l2_policy_stream() : 0;
        const int64_t n_log = ord.plane == 0 ? n : (int64_t)ord.chunks * ord.xchunk * ord.plane;
        const int64_t per_step = (int64_t)ord.chunks * ord.plane;
        // two groups of V elements per trip: all loads of both groups are issued before the first use (bytes in flight),
        // the groups are then processed in index order, so the per-thread summation order is that of the plain loop
        constexpr int U = 1;   // groups per trip (2 measured slower: registers 32 -> 40 cost a resident CTA per SM)
        const int64_t stride = (int64_t)gridDim.x * kBlock * V;
        for (int64_t j0 = ((int64_t)blockIdx.x * kBlock + threadIdx.x) * V; j0 < n_log; j0 += U * stride) {
            int64_t idx[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t j = j0 + u * stride;
                ok[u] = j < n_log;
                idx[u] = j;
                if (ord.plane != 0 && ok[u]) {
                    const int64_t tau = j / per_step, rem = j - tau * per_step;
                    const int64_t c = rem / ord.plane, o = rem - c * ord.plane;
                    const int64_t x0 = c * ord.xchunk, x1 = min(ord.planes, x0 + ord.xchunk);
                    const int64_t px = x1 - 1 - tau;
                    ok[u] = px >= x0;
                    idx[u] = px * ord.plane + o;
                }
            }
            alignas(16) T e[U][NA][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                static_for<NA>([&](auto A) {
                    constexpr int a = decltype(A)::value;
                    if constexpr (((P::RMASK >> a) & 1u) != 0) {
                        if (ok[u]) {
                            if constexpr (V == 1) e[u][a][0] = base[a][idx[u]];
                            else if (hinted && (((ord.keep_mask | ord.stream_mask) >> a) & 1u))
                                *reinterpret_cast<uint4*>(&e[u][a][0]) = ld16_hint(base[a] + idx[u], ((ord.keep_mask >> a) & 1u) ? pol_keep : pol_stream);
                            else *reinterpret_cast<VT*>(&e[u][a][0]) = *reinterpret_cast<const VT*>(base[a] + idx[u]);
                        }
                    } else {
#pragma unroll
                        for (int l = 0; l < V; ++l) e[u][a][l] = T(0);
                    }
                });
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
#pragma unroll
                for (int l = 0; l < V; ++l) {
                    T el[NA];
#pragma unroll
                    for (int a = 0; a < NA; ++a) el[a] = e[u][a][l];
                    if constexpr (HasIndex<P>::value) p.elem(el, acc, sc, off + idx[u] + l);
                    else p.elem(el, acc, sc);
#pragma unroll
                    for (int a = 0; a < NA; ++a) e[u][a][l] = el[a];
                }
                static_for<NA>([&](auto A) {
                    constexpr int a = decltype(A)::value;
                    if constexpr (((P::WMASK >> a) & 1u) != 0) {
                        if constexpr (V == 1) base[a][idx[u]] = e[u][a][0];
                        else if (hinted && (((ord.keep_mask | ord.stream_mask) >> a) & 1u))
                            st16_hint(base[a] + idx[u], *reinterpret_cast<const uint4*>(&e[u][a][0]), ((ord.keep_mask >> a) & 1u) ? pol_keep : pol_stream);
                        else *reinterpret_cast<VT*>(base[a] + idx[u]) = *reinterpret_cast<const VT*>(&e[u][a][0]);
                    }
                });
                if constexpr (HasPeer<P>::value) {     // first / last owned plane of array 0 -> the neighbours' halo planes
                    const PeerHalo& ph = p.peer;
                    const int64_t i = idx[u];
                    if (ph.lo && i < ph.plane) {
                        wrote_peer = true;
                        T* dst = (T*)ph.lo + (int64_t)b * ph.ld_lo + i;
                        if constexpr (V == 1) dst[0] = e[u][0][0];
                        else *reinterpret_cast<VT*>(dst) = *reinterpret_cast<const VT*>(&e[u][0][0]);
                    }
                    if (ph.hi && i >= n - ph.plane) {
                        wrote_peer = true;
                        T* dst = (T*)ph.hi + (int64_t)b * ph.ld_hi + (i - (n - ph.plane));
                        if constexpr (V == 1) dst[0] = e[u][0][0];
                        else *reinterpret_cast<VT*>(dst) = *reinterpret_cast<const VT*>(&e[u][0][0]);
                    }
                }
            }
        }
    }
    // peer stores are ordered before the in-kernel all-reduce by the threads that made them (system scope); the CTA barrier of the
    // block reduction and the gpu-scope fence + ticket of thread 0 carry the order on (fences are cumulative), the finishing
    // thread adds one system-scope fence before it sends its messages (peer_allreduce, ordered_data)
    if (wrote_peer) __threadfence_system();
    epilogue_tail(p, acc, b, blockIdx.x, gridDim.x);
}

template <typename T, class P>
int launch_phase(const P& p, int64_t n, int64_t ld, int batch, cudaStream_t stream, int64_t off = 0, PhaseOrder ord = PhaseOrder{}, bool pdl = false) {
    constexpr int V = Vec<T>::N;
    bool aligned = (n % V == 0) && (ld % V == 0) && (off % V == 0);
    for (int a = 0; a < P::NA; ++a) aligned = aligned && (((uintptr_t)p.v[a]) % 16 == 0);
    if (ord.plane % V != 0) ord.plane = 0;
    if (sizeof(T) * V != 16) ord.keep_mask = ord.stream_mask = 0;
    if (aligned) {
        dim3 grid((unsigned)grid_x((n / V + kBlock - 1) / kBlock, batch, occupancy(k_phase<T, P, V>)), (unsigned)batch);
        if (pdl) PHB_LAUNCH_PDL((k_phase<T, P, V>), grid, dim3(kBlock), 0, stream, p, n, ld, off, ord);
        else PHB_LAUNCH((k_phase<T, P, V>), grid, kBlock, 0, stream, p, n, ld, off, ord);
    } else {
        dim3 grid((unsigned)grid_x((n + kBlock - 1) / kBlock, batch, occupancy(k_phase<T, P, 1>)), (unsigned)batch);
        PHB_LAUNCH((k_phase<T, P, 1>), grid, kBlock, 0, stream, p, n, ld, off, PhaseOrder{});
    }
    return PHB200_OK;
}

// Scalar step of a phase / epilogue functor applied to (all-reduced) sums: one CTA, thread b handles system b.
template <class F>
__global__ void __launch_bounds__(kBlock) k_finalize(F f, const double* __restrict__ totals) {
    if (f.idle()) return;
    constexpr int ND = AccSize<F>::N;
    for (int b = threadIdx.x; b < f.batch; b += blockDim.x) {
        double tot[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) tot[d] = totals[(size_t)b * kMaxDots + d];
        f.finalize(b, tot);
    }
    __threadfence();
    __syncthreads();
    if constexpr (F::GLOBAL) f.finalize_all();
}

}  // namespace phb
