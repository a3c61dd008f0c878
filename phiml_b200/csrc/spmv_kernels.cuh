// SpMV kernels, templated on an epilogue so that the Krylov loops can fuse dot products, per-system scalar updates
// and the convergence test into the product (see solver.cu).  phb200_spmv instantiates them with NoEpi.
#pragma once
#include "matrix.cuh"

namespace phb {

// Epilogue protocol ------------------------------------------------------------------------------------------
//   ND            number of fused dot products (0: no reduction, no tickets)
//   GLOBAL        whether a batch-wide finaliser (progress check) follows the per-system one
//   idle()        true if the whole launch is a no-op (all systems finished) — uniform over the grid
//   skip(b)       true if system b is frozen — uniform over the CTA
//   row(b,i,y,acc) called once per produced element y_i; may read other vectors at i and add into acc
//   finalize(b,tot)  thread 0 of the last CTA of system b, totals in fixed summation order
//   finalize_all()   all threads of the last CTA of the batch
template <typename T>
struct NoEpi {
    static constexpr int ND = 0;
    static constexpr bool GLOBAL = false;
    __device__ bool idle() const { return false; }
    __device__ bool skip(int) const { return false; }
    __device__ void row(int, int64_t, T, double (&)[1]) const {}
    __device__ void finalize(int, const double (&)[1]) const {}
    __device__ void finalize_all() const {}
    Reduce red;
    Glob* glob;
    Sys* sys;
    int batch;
    double* defer;
};

// acc holds per-thread partial sums
template <class Epi, int NDA>
__device__ __forceinline__ void epilogue_tail(const Epi& epi, double (&acc)[NDA], int b, int bx, int nbx) {
    if constexpr (Epi::ND > 0) {
        bool last = reduce_system<Epi::ND>(acc, epi.red, b, bx, nbx);
        if (last) {
            if (threadIdx.x == 0) {
                if (epi.defer) {   // multi-GPU: hand the local sums to the host-driven all-reduce, finalise later
#pragma unroll
                    for (int d = 0; d < Epi::ND; ++d) epi.defer[(size_t)b * kMaxDots + d] = acc[d];
                } else {
                    if (epi.red.peer) peer_allreduce<Epi::ND>(*epi.red.peer, b, acc, epi.red.peer_data != 0);   // multi-GPU: sums over the ranks, in the kernel
                    epi.finalize(b, acc);
                }
            }
            if constexpr (Epi::GLOBAL) {
                if (!epi.defer && last_of_batch(epi.glob, epi.batch)) epi.finalize_all();
            }
        }
    }
}

// tot holds the CTA's sums in thread 0
template <class Epi, int NDA>
__device__ __forceinline__ void epilogue_tail_cta(const Epi& epi, double (&tot)[NDA], int b, int bx, int nbx) {
    if constexpr (Epi::ND > 0) {
        bool last = system_reduce<Epi::ND>(tot, epi.red, b, bx, nbx);
        if (last) {
            if (threadIdx.x == 0) {
                if (epi.defer) {
#pragma unroll
                    for (int d = 0; d < Epi::ND; ++d) epi.defer[(size_t)b * kMaxDots + d] = tot[d];
                } else {
                    if (epi.red.peer) peer_allreduce<Epi::ND>(*epi.red.peer, b, tot, epi.red.peer_data != 0);
                    epi.finalize(b, tot);
                }
            }
            if constexpr (Epi::GLOBAL) {
                if (!epi.defer && last_of_batch(epi.glob, epi.batch)) epi.finalize_all();
            }
        }
    }
}

template <class Epi> struct AccSize { static constexpr int N = Epi::ND > 0 ? Epi::ND : 1; };

constexpr int kCsrBt = 8;          // right-hand sides that share one staged matrix chunk

// ------------------------------------------------------------------------------------------------ CSR, streaming
// A CTA walks over chunks of 256 consecutive rows (grid-stride).  The chunk's entries (<= kCsrCap) are staged in
// shared memory with coalesced loads and reused for up to kCsrBt right-hand sides.  Each thread then adds up its own
// row sequentially with separately rounded products — the summation order of scipy's csr_matvec, bit for bit.
// Fused dot products are collected per warp in shared memory (fixed order -> deterministic).
template <typename T, bool PRESCALE, class Epi>
__global__ void __launch_bounds__(kBlock) k_spmv_csr_stream(CsrView A, const T* __restrict__ X, int64_t ldx,
                                                            T* __restrict__ Y, int64_t ldy, const T* __restrict__ scale,
                                                            int batch, int bt, Epi epi) {
    if (epi.idle()) return;
    constexpr int ND = AccSize<Epi>::N;
    __shared__ int s_col[kCsrCap];
    __shared__ T s_val[kCsrCap];
    __shared__ T s_prod[kCsrCap];
    __shared__ double s_tot[kBlock / 32][kCsrBt][ND];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int b0 = blockIdx.y * bt;
    const int nb = min(batch - b0, bt);
    for (int k = t; k < (kBlock / 32) * kCsrBt * ND; k += kBlock) (&s_tot[0][0][0])[k] = 0.0;
    const T* __restrict__ val = (const T*)A.val;
    const int64_t nchunks = (A.n_rows + kCsrRows - 1) / kCsrRows;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t r0 = chunk * kCsrRows;
        const int nrows = (int)min((int64_t)kCsrRows, A.n_rows - r0);
        const int e0 = A.rowptr[r0], e1 = A.rowptr[r0 + nrows];
        const int cnt = e1 - e0;
        __syncthreads();
        const bool direct = nb == 1 && !epi.skip(b0);     // one right-hand side: products straight from registers, no staging
        if (direct) {
            const T* __restrict__ x = X + (int64_t)b0 * ldx;
            for (int k = t; k < cnt; k += kBlock) {     // (two chains per trip and fully unrolled loads measured slower: L1/LSU bound)
                const int c = A.col[e0 + k];
                T xv = x[c];
                if (PRESCALE) xv = mul_rn(xv, scale[c]);
                s_prod[k] = mul_rn(val[e0 + k], xv);
            }
        } else {
            for (int k = t; k < cnt; k += kBlock) {
                s_col[k] = A.col[e0 + k];
                s_val[k] = val[e0 + k];
            }
        }
        int my0 = 0, my1 = 0;
        if (t < nrows) {
            my0 = A.rowptr[r0 + t] - e0;
            my1 = A.rowptr[r0 + t + 1] - e0;
        }
        __syncthreads();
        for (int bl = 0; bl < nb; ++bl) {
            const int b = b0 + bl;
            if (epi.skip(b)) continue;
            const T* __restrict__ x = X + (int64_t)b * ldx;
            if (!direct) {
                for (int k = t; k < cnt; k += kBlock) {
                    const int c = s_col[k];
                    T xv = x[c];
                    if (PRESCALE) xv = mul_rn(xv, scale[c]);
                    s_prod[k] = mul_rn(s_val[k], xv);
                }
                __syncthreads();
            }
            double acc[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) acc[d] = 0.0;
            if (t < nrows) {
                T sum = T(0);
                for (int k = my0; k < my1; ++k) sum = add_rn(sum, s_prod[k]);
                Y[(int64_t)b * ldy + r0 + t] = sum;
                epi.row(b, r0 + t, sum, acc);
            }
            if constexpr (Epi::ND > 0) {
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    double v = warp_sum(acc[d]);
                    if (lane == 0) s_tot[warp][bl][d] += v;
                }
            }
            __syncthreads();
        }
    }
    if constexpr (Epi::ND > 0) {
        for (int bl = 0; bl < nb; ++bl) {
            __syncthreads();
            double tot[ND];
            if (t == 0) {
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    double v = 0.0;
                    for (int w = 0; w < kBlock / 32; ++w) v += s_tot[w][bl][d];
                    tot[d] = v;
                }
            }
            epilogue_tail_cta(epi, tot, b0 + bl, blockIdx.x, gridDim.x);
        }
    }
}

// ------------------------------------------------------------------------------------------------ CSR, warp per row
// Fallback for matrices with long or very uneven rows (grid-stride over groups of 8 rows).
template <typename T, bool PRESCALE, class Epi>
__global__ void __launch_bounds__(kBlock) k_spmv_csr_vector(CsrView A, const T* __restrict__ X, int64_t ldx,
                                                            T* __restrict__ Y, int64_t ldy, const T* __restrict__ scale,
                                                            int batch, Epi epi) {
    if (epi.idle()) return;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[AccSize<Epi>::N];
#pragma unroll
    for (int d = 0; d < AccSize<Epi>::N; ++d) acc[d] = 0.0;
    if (!epi.skip(b)) {
        const T* __restrict__ val = (const T*)A.val;
        const T* __restrict__ x = X + (int64_t)b * ldx;
        for (int64_t row = (int64_t)blockIdx.x * (kBlock / 32) + warp; row < A.n_rows; row += (int64_t)gridDim.x * (kBlock / 32)) {
            T sum = T(0);
            for (int k = A.rowptr[row] + lane; k < A.rowptr[row + 1]; k += 32) {
                const int c = A.col[k];
                T xv = x[c];
                if (PRESCALE) xv = mul_rn(xv, scale[c]);
                sum += val[k] * xv;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
            if (lane == 0) {
                Y[(int64_t)b * ldy + row] = sum;
                epi.row(b, row, sum, acc);
            }
        }
    }
    epilogue_tail(epi, acc, b, blockIdx.x, gridDim.x);
}

// ------------------------------------------------------------------------------------------------ generic stencil
// One thread per grid point (grid-stride), any rank / point set.  (The 5/7-point fast paths live in stencil_fast.cuh.)
template <typename T, bool PRESCALE, class Epi>
__global__ void __launch_bounds__(kBlock) k_spmv_stencil(StencilDesc st, int64_t n, const T* __restrict__ X, int64_t ldx,
                                                         T* __restrict__ Y, int64_t ldy, const T* __restrict__ scale,
                                                         int batch, Epi epi) {
    if (epi.idle()) return;
    const int b = blockIdx.y;
    double acc[AccSize<Epi>::N];
#pragma unroll
    for (int d = 0; d < AccSize<Epi>::N; ++d) acc[d] = 0.0;
    if (!epi.skip(b)) {
        const int64_t nz = st.dims[2], ny = st.dims[1];
        for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
            const int64_t iz = i % nz, iy = (i / nz) % ny, ix = i / (nz * ny);
            T sum = stencil_row<T, PRESCALE>(st, X + (int64_t)b * ldx, scale, ix, iy, iz, i);
            Y[(int64_t)b * ldy + i] = sum;
            epi.row(b, i, sum, acc);
        }
    }
    epilogue_tail(epi, acc, b, blockIdx.x, gridDim.x);
}

// CTAs along x for a grid-stride kernel over `units` work units with `rows` batch rows in gridDim.y:
// one resident wave (148 SMs x occupancy), never more CTAs than work units.
inline int grid_x(int64_t units, int64_t rows, int ctas_per_sm = 8) {
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    int64_t want = ((int64_t)kSMs * ctas_per_sm) / (rows < 1 ? 1 : rows);
    if (want > units) want = units;
    if (want < 1) want = 1;
    return (int)want;
}

// resident CTAs per SM of a kernel at kBlock threads (cached per kernel address)
int occupancy_of(const void* kernel);
template <class K>
int occupancy(K kernel) { return occupancy_of((const void*)kernel); }

}  // namespace phb
