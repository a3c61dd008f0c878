// Matrix views and the SpMV device code shared by phb200_spmv and the fused Krylov kernels.
#pragma once
#include "common.cuh"

namespace phb {

struct CsrView {
    const int32_t* rowptr;
    const int32_t* col;
    const void* val;
    int64_t n_rows, n_cols, nnz;
};

// Constant-coefficient stencil with ZERO boundaries on a C-ordered grid (dims[0] slowest).  Points are sorted by
// flattened offset so that a row accumulates in the same order as the canonical (sorted-column) CSR row does.
struct StarDesc {       // centre + nearest neighbours per axis; see stencil_fast.cuh
    int64_t nx, ny, nz;     // extent of the arrays (for a slab: owned planes + one halo plane on each side)
    int64_t xb, xe;         // planes [xb, xe) are produced; reads reach planes xb-1 and xe when they lie inside the array
    int store_halo;         // slab mode: the fused CG kernel also stores d_new on the halo planes xb-1 and xe
    double c0, cxm, cxp, cym, cyp, czm, czp;
    int wrap;               // coded stars: PERIODIC axes (bit 0: x, 1: y, 2: z) — the neighbour across the boundary is the cell at the other end
};

// Pattern-coded stencil (coded.cuh): distinct flat offsets, distinct rows, one code byte per row
constexpr int kCodedMaxPat = 256;
struct CodedView {
    int K, npat;
    int k_m1, k_p1;                    // index of the offsets -1 / +1 in off[] or -1
    int fast_na;                       // >= 0: the interior pattern is [na aligned offsets, -1, 0, +1, na aligned offsets] (star stencil with 16-byte aligned rows)
    int64_t off[kMaxStencil];
    const uint8_t* code;               // [n_rows rounded up to a multiple of 4]
    const void* table;                 // [npat][kMaxStencil] of T
    // "star with a per-row diagonal": every row is the interior star with the neighbours outside the grid cut off (ZERO_GRADIENT-type boundaries);
    // only the diagonal differs between the patterns.  Such a matrix runs the fused TMA CG kernels (phb200_matrix.star holds the star, the TMA zero
    // fill makes the cut-off neighbours harmless) with the diagonal / its inverse looked up per row.
    int star_diag;
    const void* diag_tab;              // [npat] of T: diagonal of every pattern
    const void* inv_tab;               // [npat] of T: 1 / diagonal (Jacobi)
};

struct StencilDesc {
    int rank, npts;
    int64_t dims[3];        // padded with leading 1s: always (nx, ny, nz)
    int off[kMaxStencil][3];
    double coef[kMaxStencil];
    int diag;               // index of the (0,0,0) point or -1
};

}  // namespace phb

struct phb200_matrix {
    int kind;               // PHB200_MAT_*
    int dtype;
    int transposed;         // CSR arrays describe A^T
    phb::CsrView csr;
    phb::StencilDesc st;
    int is_star;            // stencil qualifies for the plane-marching kernels (star shape, nz % 4 == 0)
    phb::StarDesc star;
    phb::CodedView coded;   // PHB200_MAT_CODED: the csr view stays valid next to it
    void* owned[4];         // device arrays freed with the handle (code bytes, pattern table, diagonal tables)
    // lazily computed CSR analysis
    mutable int analysed;
    mutable int64_t max_block_nnz;   // max entries of any 256-row block
    int64_t n_rows, n_cols, nnz;
};

namespace phb {

constexpr int kCsrRows = 256;      // rows per CTA in the streaming CSR kernel
constexpr int kCsrCap = 2048;      // matrix entries staged per CTA (avg <= 8 per row)

int csr_analyse(const phb200_matrix* m, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------- stencil row
// y_i = sum_p coef[p] * x[i + off[p]] over in-domain neighbours, accumulated in ascending column order with
// separately rounded products (matches scipy's csr_matvec: sum += a*x without contraction).
template <typename T, bool PRESCALE>
__device__ __forceinline__ T stencil_row(const StencilDesc& st, const T* __restrict__ x, const T* __restrict__ scale,
                                         int64_t ix, int64_t iy, int64_t iz, int64_t i) {
    T sum = T(0);
    const int64_t ny = st.dims[1], nz = st.dims[2];
#pragma unroll 1
    for (int p = 0; p < st.npts; ++p) {
        int64_t jx = ix + st.off[p][0], jy = iy + st.off[p][1], jz = iz + st.off[p][2];
        if (jx < 0 || jx >= st.dims[0] || jy < 0 || jy >= ny || jz < 0 || jz >= nz) continue;
        int64_t j = (jx * ny + jy) * nz + jz;
        T xv = x[j];
        if (PRESCALE) xv = mul_rn(xv, scale[j]);
        sum = add_rn(sum, mul_rn((T)st.coef[p], xv));
    }
    (void)i;
    return sum;
}

}  // namespace phb
