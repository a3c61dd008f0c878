// General CSR product, matrix stream staged by the TMA unit (sm_100a).
//
// The CSR product of a matrix_from_function matrix is a pure stream: 8 bytes of (col, val) per entry make up > 80 % of the
// traffic (8 nnz + 4 (N+1) + 2 N w), every entry is used once.  k_spmv_csr_stream pulled that stream through the LSU
// (entry-parallel loads, products parked in shared memory, one more pass to add the rows up): L1/LSU-bound at 0.5 of the HBM
// peak.  Here the stream never touches a register before it is needed:
//   * a CTA walks over chunks of 256 consecutive rows (grid-stride); ONE thread asks the TMA unit for the chunk's col and val
//     segments (cp.async.bulk, 1-D, mbarrier completion, L2 evict-first: the stream must not push x out of the L2), kStagesB - 1
//     chunks ahead of the one being added up;
//   * the other threads own one row each: they read their entries from shared memory, gather x (lanes = consecutive rows, so a
//     warp's gathers fall on neighbouring addresses for banded matrices) with up to 8 loads in flight, and add the products
//     sequentially — sum += val * x with separately rounded multiply and add, i.e. SciPy's csr_matvec order, bit for bit.
// Bulk copies need 16-byte aligned addresses and sizes: the copy covers the aligned superset [e0 & ~3, (e1 + 3) & ~3) of the
// chunk's entries; a chunk whose superset would reach past the end of the arrays (at most the last one) is loaded with plain
// loads instead.  Up to kCsrBt right-hand sides share one staged chunk.
#pragma once
#include "star_tma.cuh"

namespace phb {

constexpr int kStagesB = 3;
constexpr int kBulkPad = 8;        // superset slack (<= 3 leading + <= 3 trailing entries)

template <typename T> constexpr int bulk_stage_bytes() { return (kCsrCap + kBulkPad) * (4 + (int)sizeof(T)); }
template <typename T> constexpr int bulk_smem_bytes() { return kStagesB * bulk_stage_bytes<T>() + 8 * kStagesB + 16; }

__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar), "l"(policy)
                 : "memory");
}

template <typename T, bool PRESCALE, int BT, class Epi>
__global__ void __launch_bounds__(kBlock) k_spmv_csr_bulk(CsrView A, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y, int64_t ldy,
                                                          const T* __restrict__ scale, int batch, Epi epi) {
    extern __shared__ __align__(16) unsigned char bulk_smem[];
    if (epi.idle()) return;
    constexpr int ND = AccSize<Epi>::N;
    constexpr int SB = bulk_stage_bytes<T>();
    constexpr int CAPP = kCsrCap + kBulkPad;
    __shared__ double s_tot[kBlock / 32][BT][ND];
    uint64_t* bars = reinterpret_cast<uint64_t*>(bulk_smem + (size_t)kStagesB * SB);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int b0 = blockIdx.y * BT;
    const int nb = min(batch - b0, BT);
    for (int k = t; k < (kBlock / 32) * BT * ND; k += kBlock) (&s_tot[0][0][0])[k] = 0.0;
    if (t == 0) {
#pragma unroll
        for (int k = 0; k < kStagesB; ++k) mbar_init(smem_u32(&bars[k]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const T* __restrict__ val = (const T*)A.val;
    const int32_t* __restrict__ rowptr = A.rowptr;
    const int64_t nchunks = (A.n_rows + kCsrRows - 1) / kCsrRows;
    const int64_t my_chunks = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint64_t policy = l2_evict_first_policy();

    // superset [ea, eb) of chunk j's entries; plain == true: not bulk-copyable (reaches past the arrays)
    auto bounds = [&](int64_t j, int& e0, int& e1) {
        const int64_t r0 = ((int64_t)blockIdx.x + j * gridDim.x) * kCsrRows;
        const int64_t r1 = min(A.n_rows, r0 + kCsrRows);
        e0 = rowptr[r0];
        e1 = rowptr[r1];
    };
    auto issue = [&](int64_t j, int e0, int e1) {     // thread 0 only
        const int stage = (int)(j % kStagesB);
        const uint32_t bar = smem_u32(&bars[stage]);
        const int ea = e0 & ~3;
        const int64_t eb = ((int64_t)e1 + 3) & ~(int64_t)3;
        if (e1 > e0 && eb <= A.nnz) {
            const uint32_t n = (uint32_t)(eb - ea);
            mbar_expect_tx(bar, n * (4u + (uint32_t)sizeof(T)));
            const uint32_t dst = smem_u32(bulk_smem + (size_t)stage * SB);
            bulk_load_1d(dst, A.col + ea, n * 4u, bar, policy);
            bulk_load_1d(dst + CAPP * 4u, val + ea, n * (uint32_t)sizeof(T), bar, policy);
        } else {
            mbar_expect_tx(bar, 0u);                  // nothing in flight: the phase completes at once, consumers load plainly
        }
    };

    int ne0 = 0, ne1 = 0;                             // thread 0: bounds of the next chunk to issue
    if (t == 0) {
        for (int64_t j = 0; j < kStagesB - 1 && j < my_chunks; ++j) {
            bounds(j, ne0, ne1);
            issue(j, ne0, ne1);
        }
        if (kStagesB - 1 < my_chunks) bounds(kStagesB - 1, ne0, ne1);
    }

    double acc1[ND];                                  // BT == 1: per-thread sums over all chunks (one block reduction at the end)
#pragma unroll
    for (int d = 0; d < ND; ++d) acc1[d] = 0.0;

    for (int64_t j = 0; j < my_chunks; ++j) {
        const int64_t r0 = ((int64_t)blockIdx.x + j * gridDim.x) * kCsrRows;
        const int nrows = (int)min((int64_t)kCsrRows, A.n_rows - r0);
        const int stage = (int)(j % kStagesB);
        if (t == 0 && j + kStagesB - 1 < my_chunks) {
            issue(j + kStagesB - 1, ne0, ne1);         // its stage was drained in iteration j - 1 (barrier at the end of the loop body)
            if (j + kStagesB < my_chunks) bounds(j + kStagesB, ne0, ne1);
        }
        // what does not depend on the staged data is fetched before the wait
        const int e0 = rowptr[r0], e1 = rowptr[r0 + nrows];
        int p0 = 0, p1 = 0;
        if (t < nrows) {
            p0 = rowptr[r0 + t];
            p1 = rowptr[r0 + t + 1];
        }
        const int ea = e0 & ~3;
        const bool plain = !(e1 > e0 && (((int64_t)e1 + 3) & ~(int64_t)3) <= A.nnz);
        int* s_col = reinterpret_cast<int*>(bulk_smem + (size_t)stage * SB);
        T* s_val = reinterpret_cast<T*>(bulk_smem + (size_t)stage * SB + CAPP * 4);
        mbar_wait(smem_u32(&bars[stage]), (uint32_t)((j / kStagesB) & 1));
        if (plain) {
            for (int k = e0 + t; k < e1; k += kBlock) {
                s_col[k - ea] = A.col[k];
                s_val[k - ea] = val[k];
            }
            __syncthreads();
        }
        const int k0 = p0 - ea, k1 = p1 - ea;
#pragma unroll 1
        for (int bl = 0; bl < nb; ++bl) {
            const int b = b0 + bl;
            if (epi.skip(b)) continue;
            const T* __restrict__ x = X + (int64_t)b * ldx;
            double acc[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) acc[d] = 0.0;
            if (t < nrows) {
                T sum = T(0);
                for (int k = k0; k < k1; k += 8) {
                    int c[8];
                    T v[8], xv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const bool in = k + u < k1;
                        c[u] = in ? s_col[k + u] : -1;
                        v[u] = in ? s_val[k + u] : T(0);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        xv[u] = T(0);
                        if (c[u] >= 0) {
                            xv[u] = x[c[u]];
                            if (PRESCALE) xv[u] = mul_rn(xv[u], scale[c[u]]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (c[u] >= 0) sum = add_rn(sum, mul_rn(v[u], xv[u]));
                }
                Y[(int64_t)b * ldy + r0 + t] = sum;
                if constexpr (BT == 1) epi.row(b, r0 + t, sum, acc1);
                else epi.row(b, r0 + t, sum, acc);
            }
            if constexpr (Epi::ND > 0 && BT > 1) {
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const double w = warp_sum(acc[d]);
                    if (lane == 0) s_tot[warp][bl][d] += w;
                }
            }
        }
        __syncthreads();                               // the stage may be refilled; s_tot rows are final for this chunk
    }
    if constexpr (Epi::ND > 0) {
        if constexpr (BT == 1) {
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double w = warp_sum(acc1[d]);
                if (lane == 0) s_tot[warp][0][d] = w;
            }
        }
        for (int bl = 0; bl < nb; ++bl) {
            __syncthreads();
            double tot[ND];
            if (t == 0) {
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    double w = 0.0;
                    for (int q = 0; q < kBlock / 32; ++q) w += s_tot[q][bl][d];
                    tot[d] = w;
                }
            }
            epilogue_tail_cta(epi, tot, b0 + bl, blockIdx.x, gridDim.x);
        }
    }
}

// returns false (nothing launched) when the arrays are not 16-byte aligned or the environment asks for the LSU kernel
template <typename T, bool PRESCALE, int BT, class Epi>
int launch_csr_bulk_t(const phb200_matrix* A, const T* X, int64_t ldx, T* Y, int64_t ldy, const T* scale, int batch, const Epi& epi, cudaStream_t stream) {
    auto kernel = k_spmv_csr_bulk<T, PRESCALE, BT, Epi>;
    constexpr int smem = bulk_smem_bytes<T>();
    static int occ_dev[kMaxDevices] = {};
    int& occ = occ_dev[device_slot()];
    if (occ == 0) {
        PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kBlock, smem) != cudaSuccess || v < 1) { cudaGetLastError(); v = 1; }
        occ = v;
    }
    const int rows = (batch + BT - 1) / BT;
    dim3 grid((unsigned)grid_x((A->n_rows + kCsrRows - 1) / kCsrRows, rows, occ), (unsigned)rows);
    PHB_LAUNCH(kernel, grid, kBlock, (size_t)smem, stream, A->csr, X, ldx, Y, ldy, scale, batch, epi);
    return PHB200_OK;
}

inline bool csr_bulk_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PHB200_CSR_BULK"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

template <typename T, class Epi>
int launch_csr_bulk(const phb200_matrix* A, const T* X, int64_t ldx, T* Y, int64_t ldy, const T* scale, int batch, const Epi& epi, cudaStream_t stream, bool* used) {
    *used = false;
    if (!csr_bulk_enabled() || ((uintptr_t)A->csr.col % 16) != 0 || ((uintptr_t)A->csr.val % 16) != 0 || A->n_rows == 0) return PHB200_OK;
    *used = true;
    const bool tiles = batch >= 8 * kCsrBt;          // many right-hand sides: tiles of kCsrBt share the staged chunk
    if (scale) return tiles ? launch_csr_bulk_t<T, true, kCsrBt, Epi>(A, X, ldx, Y, ldy, scale, batch, epi, stream) : launch_csr_bulk_t<T, true, 1, Epi>(A, X, ldx, Y, ldy, scale, batch, epi, stream);
    return tiles ? launch_csr_bulk_t<T, false, kCsrBt, Epi>(A, X, ldx, Y, ldy, scale, batch, epi, stream) : launch_csr_bulk_t<T, false, 1, Epi>(A, X, ldx, Y, ldy, scale, batch, epi, stream);
}

}  // namespace phb
