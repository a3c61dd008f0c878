// Pattern-coded stencils: variable-coefficient / non-ZERO-boundary operators without a matrix stream.
//
// matrix_from_function emits, besides the constant-coefficient ZERO-boundary stencils of stencil_fast.cuh, operators whose rows are
// NOT all alike: ZERO_GRADIENT (Neumann) boundaries change the diagonal and drop neighbours, PERIODIC boundaries wrap them, masks /
// obstacles switch entries off.  As CSR such a matrix costs 8 nnz + 4 (N+1) + 2 N w bytes per product (68 N for a 7-point fp32 system)
// although it holds only a handful of DISTINCT rows: a row is fully described by the flat offsets of its entries (col - row) and
// their values, and a 256^3 Neumann Laplacian has 27 such descriptions.  The coded form stores
//     off[K]        the distinct flat offsets of the whole matrix, ascending            (K <= 16)
//     table[P][16]  the P distinct rows: coefficient per offset, 0 where a row has none  (P <= 256)
//     code[N]       one byte per row: which of the P patterns it is
// so a product moves N + 2 N w bytes (9 N for fp32) — the matrix has become one byte per row.
//
// Recognition happens on the device from the CSR native (three passes over the arrays, set-up only):
//   k_coded_scan    every row: offsets must ascend (sorted columns), <= 16 entries, no stored zeros; the distinct offsets go into a 64-slot set,
//                   a 64-bit hash of the row's (offset, value bits) sequence into a 1024-slot set together with the smallest row
//                   index that shows it (CTA-local sets in shared memory first, merged once per CTA);
//   k_coded_table   one thread per pattern copies its representative row into the table;
//   k_coded_assign  every row: looks its hash up, VERIFIES its entries bit for bit against the table row (a hash collision or any
//                   mismatch fails the recognition — the matrix then simply stays CSR) and writes its code byte.
// Pattern 0 is the most frequent pattern (the interior stencil): quads whose four rows all carry code 0 take the fast path of the product.
//
// Product (k_coded_spmv): a thread owns four consecutive rows; per offset (ascending = ascending column = the order of SciPy's
// csr_matvec) the four x values come from ONE 16-byte load when the offset keeps the alignment (y / x neighbours with nz % 4 == 0),
// from the neighbouring lanes' centre quads by warp shuffle for offsets -1 / +1, and from scalar loads otherwise (periodic wraps:
// only the boundary rows have a non-zero coefficient there, all other lanes are predicated off).  Products and sums are rounded
// separately; absent entries are skipped (coefficient 0), so the row sums are bit-identical to the CSR product.
#pragma once
#include "stencil_fast.cuh"

namespace phb {

constexpr int kCodedOffSlots = 64;
constexpr int kCodedFlush = 64;        // parked quads that trigger a flush (kBlock / 4: one row per thread)
constexpr int kCodedHashSlots = 1024;
constexpr long long kCodedEmptyOff = (long long)0x8000000000000000ull;

__device__ __forceinline__ unsigned long long coded_mix(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    return h;
}
template <typename T> __device__ __forceinline__ unsigned long long coded_bits(T v);
template <> __device__ __forceinline__ unsigned long long coded_bits<float>(float v) { return (unsigned long long)__float_as_uint(v); }
template <> __device__ __forceinline__ unsigned long long coded_bits<double>(double v) { return (unsigned long long)__double_as_longlong(v); }

// hash of row i (0 is reserved for "empty slot"); ok = false when the row cannot be coded
template <typename T>
__device__ __forceinline__ unsigned long long coded_row_hash(const CsrView& A, int64_t i, bool& ok) {
    const T* __restrict__ val = (const T*)A.val;
    const int k0 = A.rowptr[i], k1 = A.rowptr[i + 1];
    if (k1 - k0 > kMaxStencil) ok = false;
    unsigned long long h = 0x243f6a8885a308d3ull;
    long long prev = kCodedEmptyOff;
    for (int k = k0; k < k1 && k < k0 + kMaxStencil; ++k) {
        const long long d = (long long)A.col[k] - (long long)i;
        if (d <= prev) ok = false;                  // unsorted or duplicate columns: the CSR product order is not the offset order
        if (val[k] == T(0)) ok = false;             // a STORED zero still multiplies its x (0 * Inf = NaN in the CSR product); the table's 0 means "no entry"
        prev = d;
        h = coded_mix(h, (unsigned long long)d);
        h = coded_mix(h, coded_bits<T>(val[k]));
    }
    return h == 0ull ? 1ull : h;
}

// set insertion with compare-and-swap; returns the slot or -1 when the set is full
template <typename W>
__device__ __forceinline__ int coded_set_insert(W* set, int slots, W key, W empty, unsigned start) {
    for (int probe = 0; probe < slots; ++probe) {
        const int s = (int)((start + (unsigned)probe) % (unsigned)slots);
        W cur = *((volatile W*)&set[s]);
        if (cur == key) return s;
        if (cur == empty) {
            cur = atomicCAS(&set[s], empty, key);
            if (cur == empty || cur == key) return s;
        }
    }
    return -1;
}

struct CodedScan {
    unsigned long long offs[kCodedOffSlots];      // distinct offsets (as bit patterns), kCodedEmptyOff = empty
    unsigned long long hash[kCodedHashSlots];     // distinct row hashes, 0 = empty
    int rep[kCodedHashSlots];                     // smallest row index with that hash
    int cnt[kCodedHashSlots];                     // rows with that hash
    int fail;
};

template <typename T>
__global__ void __launch_bounds__(kBlock) k_coded_scan(CsrView A, CodedScan* out) {
    __shared__ unsigned long long s_offs[kCodedOffSlots];
    __shared__ unsigned long long s_hash[kCodedHashSlots];
    __shared__ int s_rep[kCodedHashSlots];
    __shared__ int s_cnt[kCodedHashSlots];
    __shared__ int s_fail;
    const int t = threadIdx.x;
    for (int k = t; k < kCodedOffSlots; k += kBlock) s_offs[k] = (unsigned long long)kCodedEmptyOff;
    for (int k = t; k < kCodedHashSlots; k += kBlock) { s_hash[k] = 0ull; s_rep[k] = 0x7fffffff; s_cnt[k] = 0; }
    if (t == 0) s_fail = 0;
    __syncthreads();
    const T* __restrict__ val = (const T*)A.val;
    (void)val;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + t; i < A.n_rows; i += (int64_t)gridDim.x * kBlock) {
        bool ok = true;
        const unsigned long long h = coded_row_hash<T>(A, i, ok);
        if (ok) {
            const int k0 = A.rowptr[i], k1 = A.rowptr[i + 1];
            for (int k = k0; k < k1; ++k) {
                const unsigned long long d = (unsigned long long)((long long)A.col[k] - (long long)i);
                if (coded_set_insert<unsigned long long>(s_offs, kCodedOffSlots, d, (unsigned long long)kCodedEmptyOff, (unsigned)(d * 0x9e3779b1u)) < 0) ok = false;
            }
            const int s = coded_set_insert<unsigned long long>(s_hash, kCodedHashSlots, h, 0ull, (unsigned)(h >> 20));
            if (s < 0) ok = false;
            else {
                atomicMin(&s_rep[s], (int)i);
                const unsigned peers = __match_any_sync(__activemask(), s);      // one shared-memory add per distinct pattern and warp
                if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_cnt[s], __popc(peers));
            }
        }
        if (!ok) s_fail = 1;
    }
    __syncthreads();
    if (s_fail) { if (t == 0) out->fail = 1; return; }
    for (int k = t; k < kCodedOffSlots; k += kBlock) {
        const unsigned long long d = s_offs[k];
        if (d != (unsigned long long)kCodedEmptyOff && coded_set_insert<unsigned long long>(out->offs, kCodedOffSlots, d, (unsigned long long)kCodedEmptyOff, (unsigned)(d * 0x9e3779b1u)) < 0) out->fail = 1;
    }
    for (int k = t; k < kCodedHashSlots; k += kBlock) {
        const unsigned long long h = s_hash[k];
        if (h != 0ull) {
            const int s = coded_set_insert<unsigned long long>(out->hash, kCodedHashSlots, h, 0ull, (unsigned)(h >> 20));
            if (s < 0) out->fail = 1;
            else {
                atomicMin(&out->rep[s], s_rep[k]);
                atomicAdd(&out->cnt[s], s_cnt[k]);
            }
        }
    }
}

struct CodedDict {                 // built on the host from CodedScan
    int K, npat;
    long long off[kMaxStencil];
    unsigned long long hash[kCodedMaxPat];
    int rep[kCodedMaxPat];
    int cnt[kCodedMaxPat];         // entries of the representative row (filled by k_coded_table)
};

template <typename T>
__global__ void k_coded_table(CsrView A, CodedDict* dict, T* __restrict__ table) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dict->npat) return;
    const T* __restrict__ val = (const T*)A.val;
    const int64_t i = dict->rep[p];
    for (int k = 0; k < kMaxStencil; ++k) table[p * kMaxStencil + k] = T(0);
    const int k0 = A.rowptr[i], k1 = A.rowptr[i + 1];
    dict->cnt[p] = k1 - k0;
    for (int k = k0; k < k1; ++k) {
        const long long d = (long long)A.col[k] - (long long)i;
        for (int q = 0; q < dict->K; ++q)
            if (dict->off[q] == d) table[p * kMaxStencil + q] = val[k];
    }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_coded_assign(CsrView A, const CodedDict* __restrict__ dict, const T* __restrict__ table, uint8_t* __restrict__ code, int* fail) {
    __shared__ unsigned long long s_hash[kCodedMaxPat];
    __shared__ long long s_off[kMaxStencil];
    __shared__ int s_cnt[kCodedMaxPat];
    const int t = threadIdx.x;
    const int npat = dict->npat, K = dict->K;
    for (int k = t; k < npat; k += kBlock) { s_hash[k] = dict->hash[k]; s_cnt[k] = dict->cnt[k]; }
    if (t < kMaxStencil) s_off[t] = t < K ? dict->off[t] : 0;
    __syncthreads();
    const T* __restrict__ val = (const T*)A.val;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + t; i < A.n_rows; i += (int64_t)gridDim.x * kBlock) {
        bool ok = true;
        const unsigned long long h = coded_row_hash<T>(A, i, ok);
        int p = -1;
        if (ok) {
            if (s_hash[0] == h) p = 0;                  // the interior pattern, nearly every row
            else
                for (int q = 1; q < npat; ++q)
                    if (s_hash[q] == h) { p = q; break; }
        }
        if (p < 0) { *fail = 1; continue; }
        const int k0 = A.rowptr[i], k1 = A.rowptr[i + 1];
        if (k1 - k0 != s_cnt[p]) ok = false;
        for (int k = k0; k < k1 && ok; ++k) {          // verification: same entries, bit for bit
            const long long d = (long long)A.col[k] - (long long)i;
            int q = 0;
            while (q < K && s_off[q] != d) ++q;
            if (q == K || coded_bits<T>(table[p * kMaxStencil + q]) != coded_bits<T>(val[k])) ok = false;
        }
        if (!ok) { *fail = 1; continue; }
        code[i] = (uint8_t)p;
    }
}

// Is the coded matrix a star whose rows only lose the neighbours that lie outside the grid (and differ in the diagonal)?  Every row is checked:
// for each of the 2 * rank directions the table entry of the row's pattern must be the interior coefficient when the neighbour is inside the grid and
// absent (0) when it is outside.  Writes the diagonal of every pattern and its inverse (1 / d in T, as Jacobi computes it).
struct StarCheck {
    int64_t nx, ny, nz;
    int k_diag;
    int k_dir[6];      // table index of the direct offset per direction (x-, x+, y-, y+, z-, z+) or -1 (axis absent)
    int k_wrap[6];     // table index of the wrapped offset per direction (PERIODIC axis) or -1
};

template <typename T>
__global__ void __launch_bounds__(kBlock) k_coded_star_check(CodedView cv, int64_t n, StarCheck sc, int* fail) {
    const T* __restrict__ tab = (const T*)cv.table;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        const int64_t iz = i % sc.nz, iy = (i / sc.nz) % sc.ny, ix = i / (sc.nz * sc.ny);
        const T* row = tab + (int)cv.code[i] * kMaxStencil;
        bool ok = true;
        // direction d: inside -> the direct entry carries the interior coefficient, the wrapped one is absent; outside -> the other way round on a
        // PERIODIC axis, both absent otherwise
        auto dir = [&](int d, bool inside) {
            const int kd = sc.k_dir[d], kw = sc.k_wrap[d];
            if (kd < 0) return;
            const T want = tab[kd];                          // interior coefficient of this direction (pattern 0)
            const T direct = row[kd], wrapped = kw >= 0 ? row[kw] : T(0);
            // NOTE: a wrapped offset of one direction may coincide with nothing else (axes with >= 3 cells), so the two entries are independent
            if (inside) { if (coded_bits<T>(direct) != coded_bits<T>(want) || wrapped != T(0)) ok = false; }
            else if (kw >= 0) { if (coded_bits<T>(wrapped) != coded_bits<T>(want) || direct != T(0)) ok = false; }
            else if (direct != T(0)) ok = false;
        };
        dir(0, ix > 0); dir(1, ix + 1 < sc.nx);
        dir(2, iy > 0); dir(3, iy + 1 < sc.ny);
        dir(4, iz > 0); dir(5, iz + 1 < sc.nz);
        if (sc.k_diag < 0 || row[sc.k_diag] == T(0)) ok = false;
        if (!ok) *fail = 1;
    }
}

template <typename T>
__global__ void k_coded_diag_tables(CodedView cv, int k_diag, T* __restrict__ diag, T* __restrict__ inv) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cv.npat) return;
    const T d = ((const T*)cv.table)[p * kMaxStencil + k_diag];
    diag[p] = d;
    inv[p] = T(1) / d;
}

// ---------------------------------------------------------------------------------------------- product
template <typename T> constexpr size_t coded_smem(int npat) { return (size_t)npat * kMaxStencil * sizeof(T); }

template <typename T, bool VEC, int NA, class Epi>
__global__ void __launch_bounds__(kBlock) k_coded_spmv(CodedView cv, int64_t n, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y, int64_t ldy, Epi epi) {
    extern __shared__ __align__(16) unsigned char coded_smem_raw[];
    T* s_tab = reinterpret_cast<T*>(coded_smem_raw);
    if (epi.idle()) return;
    const int b = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31;
    constexpr int NDA = AccSize<Epi>::N;
    double acc[NDA];
#pragma unroll
    for (int d = 0; d < NDA; ++d) acc[d] = 0.0;
    if (!epi.skip(b)) {
        const T* __restrict__ tab = (const T*)cv.table;
        for (int k = t; k < cv.npat * kMaxStencil; k += kBlock) s_tab[k] = tab[k];
        __syncthreads();
        const T* __restrict__ x = X + (int64_t)b * ldx;
        T* __restrict__ y = Y + (int64_t)b * ldy;
        const int K = cv.K;
        if constexpr (VEC) {
            constexpr int VA = 16 / (int)sizeof(T);                 // elements per 16-byte access
            constexpr int G = 16 / (int)sizeof(T);                  // offsets whose loads are in flight together (fast path)
            // the interior pattern (code 0) as a list of its non-zero entries: (offset, coefficient) — what nearly every thread runs
            __shared__ int64_t s_d0[kMaxStencil];       // offset of the entry
            __shared__ int64_t s_dv0[kMaxStencil];      // offset of its 16-byte load (0 when the values come from the centre quad / scalar loads)
            __shared__ T s_c0[kMaxStencil];
            __shared__ int s_cls0[kMaxStencil];         // 0 take the loaded quad, 1 offset -1, 2 offset +1, 3 scalar loads (unaligned offset)
            __shared__ int s_n0;
            if (t == 0) {
                int m = 0;
                for (int k = 0; k < K; ++k)
                    if (s_tab[k] != T(0)) {
                        const int64_t d = cv.off[k];
                        s_d0[m] = d;
                        s_c0[m] = s_tab[k];
                        s_cls0[m] = d == -1 ? 1 : d == 1 ? 2 : (d % VA == 0 ? 0 : 3);
                        s_dv0[m] = (d % VA == 0) ? d : 0;
                        ++m;
                    }
                s_n0 = m;
            }
            __syncthreads();
            const int n0 = s_n0;
            int64_t f_dlo[NA > 0 ? NA : 1], f_dhi[NA > 0 ? NA : 1];
            T f_clo[NA > 0 ? NA : 1], f_chi[NA > 0 ? NA : 1], f_cm = T(0), f_c0 = T(0), f_cp = T(0);
            if constexpr (NA >= 0) {
#pragma unroll
                for (int u = 0; u < NA; ++u) {
                    f_dlo[u] = s_d0[u]; f_clo[u] = s_c0[u];
                    f_dhi[u] = s_d0[NA + 3 + u]; f_chi[u] = s_c0[NA + 3 + u];
                }
                f_cm = s_c0[NA]; f_c0 = s_c0[NA + 1]; f_cp = s_c0[NA + 2];
            }
            (void)f_dlo; (void)f_dhi; (void)f_clo; (void)f_chi; (void)f_cm; (void)f_c0; (void)f_cp;
            const int64_t nquads = (n + 3) / 4;
            // Quads that are not four interior rows (boundary / obstacle rows, the ragged end) are NOT computed where they are met — one such
            // lane would make its whole warp walk the per-row code path — but parked in a shared-memory list; whenever the list holds work
            // for every thread, the CTA computes the parked rows one row per thread (dense warps again).
            __shared__ int s_list[kBlock + kCodedFlush];
            __shared__ int s_cnt;
            if (t == 0) s_cnt = 0;
            __syncthreads();
            auto flush = [&](int cnt) {
                for (int base = 0; base < cnt; base += kBlock / 4) {
                    const int idx = base + (t >> 2);
                    if (idx < cnt) {
                        const int64_t i = 4 * (int64_t)s_list[idx] + (t & 3);
                        if (i < n) {
                            const T* rt = s_tab + (int)cv.code[i] * kMaxStencil;
                            T sum = T(0);
                            for (int k = 0; k < K; ++k) {
                                const T c = rt[k];
                                if (c != T(0)) sum = add_rn(sum, mul_rn(c, x[i + cv.off[k]]));
                            }
                            y[i] = sum;
                            epi.row(b, i, sum, acc);
                        }
                    }
                }
                __syncthreads();
                if (t == 0) s_cnt = 0;
                __syncthreads();
            };
            for (int64_t cq = (int64_t)blockIdx.x * kBlock; cq < nquads; cq += (int64_t)gridDim.x * kBlock) {
                const int64_t q = cq + t;
                const int64_t i0 = 4 * q;
                const bool valid = q < nquads;
                const bool full = valid && i0 + 4 <= n;
                unsigned cw = 0;
                if (valid) cw = *reinterpret_cast<const unsigned*>(cv.code + i0);
                const bool fast = full && cw == 0u;
                Quad<T> xc{{T(0), T(0), T(0), T(0)}};
                if (full) xc = ld4<T>(x + i0);
                // z neighbours inside the warp's 128 consecutive rows travel by shuffle; the edge lanes load theirs
                T zl = __shfl_up_sync(0xffffffffu, xc.v[3], 1), zr = __shfl_down_sync(0xffffffffu, xc.v[0], 1);
                if (fast) {
                    // ---- four interior rows.  The loads of G offsets are issued before their products are added.
                    if (lane == 0 && i0 > 0) zl = x[i0 - 1];
                    if (lane == 31 && i0 + 4 < n) zr = x[i0 + 4];
                    T sum[4] = {T(0), T(0), T(0), T(0)};
                    if constexpr (NA >= 0) {
                        // ---- star stencil with aligned rows (2-D: NA = 1, 3-D: NA = 2): entries [NA aligned, -1, 0, +1, NA aligned], coefficients and
                        // offsets in registers, all 2 NA neighbour quads requested before the first product
                        Quad<T> lo[NA > 0 ? NA : 1], hi[NA > 0 ? NA : 1];
#pragma unroll
                        for (int u = 0; u < NA; ++u) {
                            lo[u] = ld4<T>(x + i0 + f_dlo[u]);
                            hi[u] = ld4<T>(x + i0 + f_dhi[u]);
                        }
#pragma unroll
                        for (int u = 0; u < NA; ++u) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) sum[e] = add_rn(sum[e], mul_rn(f_clo[u], lo[u].v[e]));
                        }
                        sum[0] = add_rn(sum[0], mul_rn(f_cm, zl));
                        sum[1] = add_rn(sum[1], mul_rn(f_cm, xc.v[0]));
                        sum[2] = add_rn(sum[2], mul_rn(f_cm, xc.v[1]));
                        sum[3] = add_rn(sum[3], mul_rn(f_cm, xc.v[2]));
#pragma unroll
                        for (int e = 0; e < 4; ++e) sum[e] = add_rn(sum[e], mul_rn(f_c0, xc.v[e]));
                        sum[0] = add_rn(sum[0], mul_rn(f_cp, xc.v[1]));
                        sum[1] = add_rn(sum[1], mul_rn(f_cp, xc.v[2]));
                        sum[2] = add_rn(sum[2], mul_rn(f_cp, xc.v[3]));
                        sum[3] = add_rn(sum[3], mul_rn(f_cp, zr));
#pragma unroll
                        for (int u = 0; u < NA; ++u) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) sum[e] = add_rn(sum[e], mul_rn(f_chi[u], hi[u].v[e]));
                        }
                    } else {
    #pragma unroll 1
                        for (int k0 = 0; k0 < n0; k0 += G) {
                            T c[G];
                            Quad<T> v[G];
                            int cls[G];
                            // (a) G unconditional 16-byte loads back to back (entries that need none re-read the centre quad: an L1 hit)
    #pragma unroll
                            for (int u = 0; u < G; ++u) {
                                const int kk = min(k0 + u, n0 - 1);
                                c[u] = k0 + u < n0 ? s_c0[kk] : T(0);
                                cls[u] = s_cls0[kk];
                                v[u] = ld4<T>(x + i0 + s_dv0[kk]);
                            }
                            // (b) z neighbours from registers, unaligned offsets by scalar loads (uniform over the CTA: same list for every thread)
    #pragma unroll
                            for (int u = 0; u < G; ++u) {
                                if (cls[u] == 1) v[u] = Quad<T>{{zl, xc.v[0], xc.v[1], xc.v[2]}};
                                else if (cls[u] == 2) v[u] = Quad<T>{{xc.v[1], xc.v[2], xc.v[3], zr}};
                                else if (cls[u] == 3) {
                                    const int64_t j0 = i0 + s_d0[min(k0 + u, n0 - 1)];
    #pragma unroll
                                    for (int e = 0; e < 4; ++e) v[u].v[e] = x[j0 + e];
                                }
                            }
                            // (c) products in ascending offset order
    #pragma unroll
                            for (int u = 0; u < G; ++u)
                                if (c[u] != T(0)) {
    #pragma unroll
                                    for (int e = 0; e < 4; ++e) sum[e] = add_rn(sum[e], mul_rn(c[u], v[u].v[e]));
                                }
                        }
                    }
                    st4(y + i0, Quad<T>{{sum[0], sum[1], sum[2], sum[3]}});
#pragma unroll
                    for (int e = 0; e < 4; ++e) epi.row(b, i0 + e, sum[e], acc);
                } else if (valid) {
                    s_list[atomicAdd(&s_cnt, 1)] = (int)q;
                }
                __syncthreads();
                const int cnt = s_cnt;
                __syncthreads();                             // nobody parks the next quad before everybody has read the count
                if (cnt >= kCodedFlush) flush(cnt);          // uniform decision
            }
            __syncthreads();
            if (s_cnt > 0) flush(s_cnt);
        } else {
            for (int64_t i = (int64_t)blockIdx.x * kBlock + t; i < n; i += (int64_t)gridDim.x * kBlock) {
                const T* rt = s_tab + (int)cv.code[i] * kMaxStencil;
                T sum = T(0);
                for (int k = 0; k < K; ++k) {
                    const T c = rt[k];
                    if (c != T(0)) sum = add_rn(sum, mul_rn(c, x[i + cv.off[k]]));
                }
                y[i] = sum;
                epi.row(b, i, sum, acc);
            }
        }
    }
    epilogue_tail(epi, acc, b, blockIdx.x, gridDim.x);
}

template <typename T, int NA, class Epi>
int launch_coded_vec(const CodedView& cv, int64_t n, const T* X, int64_t ldx, T* Y, int64_t ldy, int batch, const Epi& epi, cudaStream_t stream) {
    auto kernel = k_coded_spmv<T, true, NA, Epi>;
    static int occ_dev[kMaxDevices] = {};
    int& occ = occ_dev[device_slot()];
    if (occ == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kBlock, coded_smem<T>(32)) != cudaSuccess || v < 1) { cudaGetLastError(); v = 4; }
        occ = v;
    }
    dim3 grid((unsigned)grid_x((n + 4 * kBlock - 1) / (4 * kBlock), batch, occ), (unsigned)batch);
    PHB_LAUNCH(kernel, grid, kBlock, coded_smem<T>(cv.npat), stream, cv, n, X, ldx, Y, ldy, epi);
    return PHB200_OK;
}

template <typename T, class Epi>
int launch_coded_spmv(const CodedView& cv, int64_t n, const T* X, int64_t ldx, T* Y, int64_t ldy, int batch, const Epi& epi, cudaStream_t stream) {
    const size_t smem = coded_smem<T>(cv.npat);
    const bool vec = ((uintptr_t)X % 16) == 0 && ((uintptr_t)Y % 16) == 0 && (ldx * sizeof(T)) % 16 == 0 && (ldy * sizeof(T)) % 16 == 0;
    if (vec) {
        static int generic_only = -1;       // PHB200_CODED_FAST=0: the list-driven fast path for every pattern (A/B)
        if (generic_only < 0) { const char* e = getenv("PHB200_CODED_FAST"); generic_only = (e && e[0] == '0') ? 1 : 0; }
        if (cv.fast_na == 2 && !generic_only) return launch_coded_vec<T, 2, Epi>(cv, n, X, ldx, Y, ldy, batch, epi, stream);
        if (cv.fast_na == 1 && !generic_only) return launch_coded_vec<T, 1, Epi>(cv, n, X, ldx, Y, ldy, batch, epi, stream);
        return launch_coded_vec<T, -1, Epi>(cv, n, X, ldx, Y, ldy, batch, epi, stream);
    } else {
        auto kernel = k_coded_spmv<T, false, -1, Epi>;
        dim3 grid((unsigned)grid_x((n + kBlock - 1) / kBlock, batch, 8), (unsigned)batch);
        PHB_LAUNCH(kernel, grid, kBlock, smem, stream, cv, n, X, ldx, Y, ldy, epi);
    }
    return PHB200_OK;
}

}  // namespace phb
