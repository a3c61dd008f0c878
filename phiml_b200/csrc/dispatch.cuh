// Host-side choice of the SpMV kernel for a matrix handle (used by phb200_spmv and by the Krylov loops).
#pragma once
#include "star_tma.cuh"
#include "csr_bulk.cuh"
#include "coded.cuh"

namespace phb {

inline bool coded_star_tma() {        // PHB200_CODED_STAR_TMA=0: coded stars through the flat coded product (A/B)
    static int v = -1;
    if (v < 0) { const char* e = getenv("PHB200_CODED_STAR_TMA"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

// Host-side dispatch used by phb200_spmv and by the solver.  `scale` != nullptr selects y = A (scale .* x).
template <typename T, class Epi>
int launch_spmv(const phb200_matrix* A, const T* X, int64_t ldx, T* Y, int64_t ldy, const T* scale, int batch,
                const Epi& epi, cudaStream_t stream) {
    if (A->kind == PHB200_MAT_STENCIL && A->is_star && !scale) {
        bool used = false;     // TMA-staged product first; falls back when no tensor map can be built for X
        PHB_TRY((launch_star_spmv_tma<T, Epi>(A->star, X, ldx, Y, ldy, batch, A->dtype, epi, stream, &used)));
        if (used) return PHB200_OK;
        PlainLoad<T> loader{X, ldx};
        return launch_star<T, PlainLoad<T>, Epi, false>(A->star, loader, Y, ldy, batch, epi, stream);
    }
    if (A->kind == PHB200_MAT_STENCIL) {
        const int64_t n = A->n_rows;
        const int occ = scale ? occupancy(k_spmv_stencil<T, true, Epi>) : occupancy(k_spmv_stencil<T, false, Epi>);
        dim3 grid((unsigned)grid_x((n + kBlock - 1) / kBlock, batch, occ), (unsigned)batch);
        if (scale) PHB_LAUNCH((k_spmv_stencil<T, true, Epi>), grid, kBlock, 0, stream, A->st, n, X, ldx, Y, ldy, scale, batch, epi);
        else PHB_LAUNCH((k_spmv_stencil<T, false, Epi>), grid, kBlock, 0, stream, A->st, n, X, ldx, Y, ldy, scale, batch, epi);
        return PHB200_OK;
    }
    if (A->kind == PHB200_MAT_CODED && !scale) {
        if (A->coded.star_diag && A->is_star && coded_star_tma()) {
            // star with a per-row diagonal (ZERO_GRADIENT-type boundaries): the TMA-staged stencil product with the diagonal looked up per row
            bool used = false;
            const StarDiag dg{A->coded.code, A->coded.diag_tab, A->coded.inv_tab, A->coded.npat, nullptr, 0};
            PHB_TRY((launch_star_spmv_tma_diag<T, Epi>(A->star, dg, X, ldx, Y, ldy, batch, A->dtype, epi, stream, &used)));
            if (used) return PHB200_OK;
        }
        return launch_coded_spmv<T, Epi>(A->coded, A->n_rows, X, ldx, Y, ldy, batch, epi, stream);
    }
    if (A->transposed) {
        set_error("fused product with a transposed (CSC-native) matrix is not supported; use phb200_spmv");
        return PHB200_ERR_UNSUPPORTED;
    }
    PHB_TRY(csr_analyse(A, stream));
    if (A->max_block_nnz <= kCsrCap) {
        bool used = false;     // matrix stream staged by the TMA unit (csr_bulk.cuh); falls back when the arrays are not 16-byte aligned
        PHB_TRY((launch_csr_bulk<T, Epi>(A, X, ldx, Y, ldy, scale, batch, epi, stream, &used)));
        if (used) return PHB200_OK;
        // few right-hand sides: one per CTA row for parallelism; many: tiles of kCsrBt share the staged matrix chunk
        const int bt = batch >= 8 * kCsrBt ? kCsrBt : 1;
        const int rows = (batch + bt - 1) / bt;
        const int occ = scale ? occupancy(k_spmv_csr_stream<T, true, Epi>) : occupancy(k_spmv_csr_stream<T, false, Epi>);
        dim3 grid((unsigned)grid_x((A->n_rows + kCsrRows - 1) / kCsrRows, rows, occ), (unsigned)rows);
        if (scale) PHB_LAUNCH((k_spmv_csr_stream<T, true, Epi>), grid, kBlock, 0, stream, A->csr, X, ldx, Y, ldy, scale, batch, bt, epi);
        else PHB_LAUNCH((k_spmv_csr_stream<T, false, Epi>), grid, kBlock, 0, stream, A->csr, X, ldx, Y, ldy, scale, batch, bt, epi);
    } else {
        const int occ = scale ? occupancy(k_spmv_csr_vector<T, true, Epi>) : occupancy(k_spmv_csr_vector<T, false, Epi>);
        dim3 grid((unsigned)grid_x((A->n_rows + 7) / 8, batch, occ), (unsigned)batch);
        if (scale) PHB_LAUNCH((k_spmv_csr_vector<T, true, Epi>), grid, kBlock, 0, stream, A->csr, X, ldx, Y, ldy, scale, batch, epi);
        else PHB_LAUNCH((k_spmv_csr_vector<T, false, Epi>), grid, kBlock, 0, stream, A->csr, X, ldx, Y, ldy, scale, batch, epi);
    }
    return PHB200_OK;
}

}  // namespace phb
