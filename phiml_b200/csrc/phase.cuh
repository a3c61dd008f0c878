// Element-wise "phase" kernel: one pass over up to 8 vectors of a batched system with fused dot products, the
// per-system scalar update and (optionally) the batch-wide progress check in the last CTA.  128-bit accesses.
#pragma once
#include "spmv_kernels.cuh"

#include <utility>

namespace phb {

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

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
__global__ void __launch_bounds__(kBlock) k_phase(P p, int64_t n, int64_t ld, int64_t off) {
    if (p.idle()) return;
    const int b = blockIdx.y;
    constexpr int ND = AccSize<P>::N;
    constexpr int NA = P::NA;
    double acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[d] = 0.0;
    if (!p.skip(b)) {
        const typename P::Scal sc = p.load(b);
        T* base[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) base[a] = (T*)p.v[a] + off + (((P::SHARED >> a) & 1u) ? (int64_t)0 : (int64_t)b * ld);
        using VT = typename Vec<T>::type;
        for (int64_t i = ((int64_t)blockIdx.x * kBlock + threadIdx.x) * V; i < n; i += (int64_t)gridDim.x * kBlock * V) {
            alignas(16) T e[NA][V];
            static_for<NA>([&](auto A) {
                constexpr int a = decltype(A)::value;
                if constexpr (((P::RMASK >> a) & 1u) != 0) {
                    if constexpr (V == 1) e[a][0] = base[a][i];
                    else *reinterpret_cast<VT*>(&e[a][0]) = *reinterpret_cast<const VT*>(base[a] + i);
                } else {
#pragma unroll
                    for (int l = 0; l < V; ++l) e[a][l] = T(0);
                }
            });
#pragma unroll
            for (int l = 0; l < V; ++l) {
                T el[NA];
#pragma unroll
                for (int a = 0; a < NA; ++a) el[a] = e[a][l];
                p.elem(el, acc, sc);
#pragma unroll
                for (int a = 0; a < NA; ++a) e[a][l] = el[a];
            }
            static_for<NA>([&](auto A) {
                constexpr int a = decltype(A)::value;
                if constexpr (((P::WMASK >> a) & 1u) != 0) {
                    if constexpr (V == 1) base[a][i] = e[a][0];
                    else *reinterpret_cast<VT*>(base[a] + i) = *reinterpret_cast<const VT*>(&e[a][0]);
                }
            });
        }
    }
    epilogue_tail(p, acc, b, blockIdx.x, gridDim.x);
}

template <typename T, class P>
int launch_phase(const P& p, int64_t n, int64_t ld, int batch, cudaStream_t stream, int64_t off = 0) {
    constexpr int V = Vec<T>::N;
    bool aligned = (n % V == 0) && (ld % V == 0) && (off % V == 0);
    for (int a = 0; a < P::NA; ++a) aligned = aligned && (((uintptr_t)p.v[a]) % 16 == 0);
    if (aligned) {
        dim3 grid((unsigned)grid_x((n / V + kBlock - 1) / kBlock, batch, occupancy(k_phase<T, P, V>)), (unsigned)batch);
        PHB_LAUNCH((k_phase<T, P, V>), grid, kBlock, 0, stream, p, n, ld, off);
    } else {
        dim3 grid((unsigned)grid_x((n + kBlock - 1) / kBlock, batch, occupancy(k_phase<T, P, 1>)), (unsigned)batch);
        PHB_LAUNCH((k_phase<T, P, 1>), grid, kBlock, 0, stream, p, n, ld, off);
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
