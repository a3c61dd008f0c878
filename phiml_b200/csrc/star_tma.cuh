// Fused stencil CG iteration, first half, with TMA staging (sm_100a):
//     d_new = M^-1 r + beta d_old   (tile + halo recomputed on chip, stored once)
//     x    += alpha_prev d_old      (the previous iteration's pending update)
//     q     = A d_new               (5/7-point star, ZERO boundaries)
//     d_new . q  -> alpha           (deterministic cross-CTA reduction, scalar step in the last CTA)
//
// A CTA owns an (SY x SZ) tile of the y-z plane and marches over a chunk of x planes.  The r and d_old tiles INCLUDING
// their one-cell halos are fetched by the TMA unit (cp.async.bulk.tensor.4d, one elected thread, mbarrier completion,
// kStages planes in flight); out-of-domain halo cells are zero-filled by the TMA, which is exactly the ZERO boundary.
// d_new planes live in a 4-deep ring in shared memory; x-neighbours of the product come from that ring.
#pragma once
#include <cuda.h>
#include "stencil_fast.cuh"

namespace phb {

constexpr int kStages = 4;   // measured 2..8: 4 is best (3 CTAs per SM, 75.2 us; 3 stages / 4 CTAs 77.5 us; >= 5 stages / 2 CTAs 84 us)

template <typename T> constexpr int tile_bytes() { return ((SROWS * SROW * (int)sizeof(T) + 127) / 128) * 128; }
template <typename T> constexpr int star_tma_smem() { return (2 * kStages + 4) * tile_bytes<T>() + 8 * kStages + 128; }

// ---- host: tensor maps -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}

// (batch, nx, ny, nz) row-major vector set with leading dimension ld -> 4-D map with box (SROW, SROWS, 1, 1)
inline bool make_star_tmap(CUtensorMap* out, const void* base, int dtype, const StarDesc& st, int64_t batch, int64_t ld, int box_z = SROW, int box_y = SROWS) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t w = dtype == PHB200_F32 ? 4 : 8;
    cuuint64_t dims[4] = {(cuuint64_t)st.nz, (cuuint64_t)st.ny, (cuuint64_t)st.nx, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)st.nz * w, (cuuint64_t)st.ny * st.nz * w, (cuuint64_t)ld * w};
    cuuint32_t box[4] = {(cuuint32_t)box_z, (cuuint32_t)box_y, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int k = 0; k < 3; ++k)
        if (strides[k] % 16 != 0 || strides[k] >= (1ull << 40)) return false;
    if (((uintptr_t)base) % 16 != 0) return false;
    static int promo = -1;            // PHB200_TMA_PROMO: 0 none, 1 64 B, 2 128 B (default), 3 256 B
    if (promo < 0) { const char* e = getenv("PHB200_TMA_PROMO"); promo = e ? atoi(e) : 2; }
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                      : promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult r = fn(out, dtype == PHB200_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ---- device: mbarrier / TMA wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a transfer that never completes traps (error surfaces on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 8000000000ll) __trap();
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// Work decomposition of the plane-marching kernels.  xchunk > 0: items = (x-chunk of xchunk planes) x (y-z tile), grid-stride.
// xchunk <= 0 (balanced): the tiles x planes space, ordered tile-major, is cut into gridDim.x equal ranges; a CTA walks its range as one
// or two x-segments (the range may run over the end of a tile's column into the next tile).  With 64 tiles and 444 resident CTAs
// (256^3, 3 CTAs per SM) equal chunks can only use 6 x 64 = 384 CTAs — 60 SM slots stay empty; equal ranges fill every slot.
struct SegIter {
    int64_t x0, x1;
    int ty, tz;
    int64_t u, u_end, planes, xb;
    int64_t item, items;
    int xchunk, tiles_y, tiles_z;
    __device__ SegIter(const StarDesc& st, int xchunk_, int tiles_y_, int tiles_z_) : xchunk(xchunk_), tiles_y(tiles_y_), tiles_z(tiles_z_) {
        planes = st.xe - st.xb;
        xb = st.xb;
        if (xchunk <= 0) {
            const int64_t W = (int64_t)tiles_y * tiles_z * planes;
            u = W * (int64_t)blockIdx.x / gridDim.x;
            u_end = W * ((int64_t)blockIdx.x + 1) / gridDim.x;
        } else {
            item = blockIdx.x;
            items = ((planes + xchunk - 1) / xchunk) * tiles_y * tiles_z;
        }
    }
    __device__ bool next() {
        if (xchunk <= 0) {
            if (u >= u_end) return false;
            const int64_t tile = u / planes, xs = u - tile * planes;
            const int64_t len = min(planes - xs, u_end - u);
            tz = (int)(tile % tiles_z);
            ty = (int)(tile / tiles_z);
            x0 = xb + xs;
            x1 = x0 + len;
            u += len;
            return true;
        }
        if (item >= items) return false;
        tz = (int)(item % tiles_z);
        ty = (int)((item / tiles_z) % tiles_y);
        const int64_t ch = item / ((int64_t)tiles_z * tiles_y);
        x0 = xb + ch * xchunk;
        x1 = min(xb + planes, x0 + xchunk);
        item += gridDim.x;
        return true;
    }
};

__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(dst),
                 "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
                 : "memory");
}
template <typename T> __device__ __forceinline__ Quad<T> ld4_hint(const T* p, uint64_t pol) {
    Quad<T> q;
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<uint4*>(&q.v[0]) = ld16_hint(p, pol);
    } else {
        *reinterpret_cast<uint4*>(&q.v[0]) = ld16_hint(p, pol);
        *reinterpret_cast<uint4*>(&q.v[2]) = ld16_hint(p + 2, pol);
    }
    return q;
}
template <typename T> __device__ __forceinline__ void st4_hint(T* p, const Quad<T>& q, uint64_t pol) {
    if constexpr (sizeof(T) == 4) {
        st16_hint(p, *reinterpret_cast<const uint4*>(&q.v[0]), pol);
    } else {
        st16_hint(p, *reinterpret_cast<const uint4*>(&q.v[0]), pol);
        st16_hint(p + 2, *reinterpret_cast<const uint4*>(&q.v[2]), pol);
    }
}

// L2 hint bits of k_star_cg_tma (two bits per stream: 0 plain, 1 keep = evict_last, 2 stream = evict_first):
//   bits 0-1 r (TMA loads), 2-3 d_old (TMA loads), 4-5 x (loads and stores), 6-7 d_new (stores), 8-9 q (stores)
// Per-row diagonal of a pattern-coded star (coded.cuh, CodedView::star_diag): code byte per row, diagonal / inverse diagonal per pattern
struct StarDiag {
    const uint8_t* code;
    const void* diag_tab;
    const void* inv_tab;
    int npat;
    const void* x;          // product kernel: the input vector itself (PERIODIC axes: neighbours across the boundary are loaded from it directly)
    int64_t ldx;
};

// PRE: 0 none, 3 Jacobi with constant inverse diagonal (DIAG: inverse diagonal of the row's pattern)
// ALG 0: CG direction d = M^-1 r + beta d ; ALG 1: CG-adaptive direction d = r - (r.q / d.q) d with the second dot d.r
// DIAG: the diagonal coefficient (and its inverse) is looked up per row through the pattern code — stars whose boundary rows only lose the neighbours
//       outside the grid (ZERO_GRADIENT-type): the TMA zero fill makes the cut-off neighbours contribute a signed zero, exactly like ZERO boundaries
template <typename T, int PRE, class Epi, bool UNIT, int ALG = 0, bool DIAG = false>
__global__ void __launch_bounds__(kBlock) k_star_cg_tma(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_a,
                                                        const __grid_constant__ CUtensorMap tm_b, StarDesc st, T* __restrict__ X, T* buf_a, T* buf_b,
                                                        T* __restrict__ Q, T cinv, int64_t ld, int xchunk, int apply_x, Epi epi, unsigned l2h = 0, StarDiag dg = StarDiag{}) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TB = tile_bytes<T>();
    constexpr uint32_t kTxBytes = 2u * SROWS * SROW * sizeof(T);
    // 128-byte aligned start inside the dynamic window, by pointer arithmetic on the shared array (an integer round trip makes
    // the compiler forget the address space: every tile access became a generic LD.E / ST.E with 64-bit address arithmetic)
    unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)(2 * kStages + 4) * TB);

    const int b = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int NDA = AccSize<Epi>::N;
    double acc[NDA];
#pragma unroll
    for (int dd = 0; dd < NDA; ++dd) acc[dd] = 0.0;
    __shared__ T s_c0[DIAG ? 256 : 1], s_inv[DIAG ? 256 : 1];
    if constexpr (DIAG) {
        if (t < dg.npat) {
            s_c0[t] = ((const T*)dg.diag_tab)[t];
            s_inv[t] = ((const T*)dg.inv_tab)[t];
        }
    }
    if (t == 0) {
#pragma unroll
        for (int k = 0; k < kStages; ++k) mbar_init(smem_u32(&bars[k]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();                    // everything above overlaps the tail of the previous kernel (PDL launch)
    if (epi.idle()) return;

    if (!epi.skip(b)) {
        const bool odd = (epi.glob->pass & 1) != 0;
        const CUtensorMap* tm_old = odd ? &tm_b : &tm_a;
        T* __restrict__ d_new = (odd ? buf_a : buf_b) + (int64_t)b * ld;
        T* __restrict__ xb = X + (int64_t)b * ld;
        T* __restrict__ qb = Q + (int64_t)b * ld;
        const T beta = (T)epi.sys[b].s[1 /*S_BETA*/], alpha = (T)epi.sys[b].s[0 /*S_ALPHA*/];
        const int64_t nx = st.nx, ny = st.ny, nz = st.nz;
        const int tiles_z = (int)((nz + SZ - 1) / SZ), tiles_y = (int)((ny + SY - 1) / SY);
        const T c0 = (T)st.c0, cxm = (T)st.cxm, cxp = (T)st.cxp, cym = (T)st.cym, cyp = (T)st.cyp, czm = (T)st.czm, czp = (T)st.czp;

        const uint64_t pol_keep = l2h ? l2_policy_keep() : 0, pol_stream = l2h ? l2_policy_stream() : 0;
        auto pol_of = [&](unsigned sh) { return ((l2h >> sh) & 3u) == 1u ? pol_keep : pol_stream; };
        auto has = [&](unsigned sh) { return ((l2h >> sh) & 3u) != 0u; };
        const T ad_rq = (T)epi.sys[b].s[5 /*S_RQ*/], ad_dq = (T)epi.sys[b].s[3 /*S_DQ*/];
        auto dir = [&](T rv, T dv, T ci) {
            if constexpr (ALG == 1) {      // AdDirection: d = r - safe_div(rq * d, dq)
                const T tt = mul_rn(ad_rq, dv);
                const T u = ad_dq == T(0) ? T(0) : tt / ad_dq;
                return add_rn(rv, -u);
            } else {
                const T s = PRE == 3 ? mul_rn(rv, ci) : rv;
                return add_rn(s, mul_rn(beta, dv));
            }
        };
        constexpr bool DINV = DIAG && PRE == 3 && ALG == 0;      // the direction needs the inverse diagonal of every cell it is formed on (tile + halo)
        const T cinv0 = DIAG ? s_inv[0] : cinv;
        const T c00 = DIAG ? s_c0[0] : c0;
        auto inv_of = [&](unsigned cw, int e) { return DINV ? (cw == 0u ? cinv0 : s_inv[(cw >> (8 * e)) & 0xffu]) : cinv; };

        // running state instead of per-plane index arithmetic: TMA stage + its mbarrier phase bits, byte offsets of the
        // tile pair and of the three live planes of the d_new ring, element offset of the current plane in global memory
        unsigned char* const tiles0 = base;
        unsigned char* const planes0 = base + (size_t)(2 * kStages) * TB;
        const int64_t pstride = ny * nz;
        const int c = 4 + 4 * lane;            // interior column of this lane in a shared-memory row
        const int sr = warp + 1;               // interior row of this warp
        uint32_t stg = 0, phase_bits = 0;      // next stage to be consumed; bit s = parity the next wait on stage s expects
        uint32_t issue_stg = 0;                // next stage to be filled (thread 0)

        SegIter seg(st, xchunk, tiles_y, tiles_z);
        while (seg.next()) {
            const int tz = seg.tz, ty = seg.ty;
            const int64_t x0 = seg.x0, x1 = seg.x1;
            const int y0 = ty * SY, z0 = tz * SZ;
            const int jy = y0 + warp, jz = z0 + 4 * lane;
            const bool mine = jy < ny && jz < nz;
            const int P = (int)(x1 - x0) + 2;   // planes x0-1 .. x1
            const bool halo_lo = st.store_halo && mine && x0 == st.xb && x0 - 1 >= 0;      // slab halo planes: keep the
            const bool halo_hi = st.store_halo && mine && x1 == st.xe && x1 < nx;          // neighbour's direction current

            auto issue = [&](int px) {
                const uint32_t bar = smem_u32(&bars[issue_stg]);
                unsigned char* tr = tiles0 + (size_t)(2 * issue_stg) * TB;
                mbar_expect_tx(bar, kTxBytes);
                if (has(0)) tma_load_4d_hint(smem_u32(tr), &tm_r, bar, z0 - 4, y0 - 1, px, b, pol_of(0));
                else tma_load_4d(smem_u32(tr), &tm_r, bar, z0 - 4, y0 - 1, px, b);
                if (has(2)) tma_load_4d_hint(smem_u32(tr + TB), tm_old, bar, z0 - 4, y0 - 1, px, b, pol_of(2));
                else tma_load_4d(smem_u32(tr + TB), tm_old, bar, z0 - 4, y0 - 1, px, b);
                issue_stg = issue_stg + 1 == kStages ? 0 : issue_stg + 1;
            };

            __syncthreads();   // everything of the previous item has been consumed
            if (t == 0)
                for (int k = 0; k < kStages && k < P; ++k) issue((int)x0 - 1 + k);
            Quad<T> x_cur{{T(0), T(0), T(0), T(0)}};
            Quad<T> r_prev{{T(0), T(0), T(0), T(0)}};          // r of the plane whose product is formed next (ALG 1: d.r)
            // the thread's own quad of d_new of the last two planes and the row's halo-column value (first / last lane) stay in registers:
            // only the y neighbours (other warps' rows) are read back from the plane ring; z neighbours inside a row travel by warp shuffle
            Quad<T> dn_m{{T(0), T(0), T(0), T(0)}}, dn_c{{T(0), T(0), T(0), T(0)}};
            T hz_c = T(0);
            const int hcol = lane == 0 ? 3 : 4 + SZ;
            int64_t off = ((x0 - 1) * ny + jy) * nz + jz;      // element offset of (plane x0-1, jy, jz); advanced per plane
            // DIAG: pattern codes of the cells this thread forms directions / products on, fetched one plane ahead: its own quad (one 32-bit
            // word), the halo row of warps 0 / 1 and the halo column of the edge lanes (inverse diagonal of the halo cells, Jacobi only)
            const int jy_h = warp == 0 ? y0 - 1 : y0 + SY;                  // halo row formed by warp 0 / 1
            const bool hrow_in = DINV && warp < 2 && jy_h >= 0 && jy_h < ny && jz < nz;
            const int64_t hrow_d = (int64_t)(jy_h - jy) * nz;
            const int jz_h = lane == 0 ? z0 - 1 : z0 + SZ;                  // halo column formed by lane 0 / 31
            const bool hcol_in = DINV && (lane == 0 || lane == 31) && jz_h >= 0 && jz_h < nz && jy < ny;
            const int64_t hcol_d = (int64_t)(jz_h - jz);
            unsigned cw_c = 0, cwh_c = 0, cb_c = 0;                        // own quad, halo row, halo column (codes of the plane being formed)
            T c0_p[4] = {c00, c00, c00, c00};                              // diagonals of the own quad of the previous plane (the product's rows)
            if constexpr (DIAG) {
                if (x0 - 1 >= 0) {
                    if (mine) cw_c = *reinterpret_cast<const unsigned*>(dg.code + off);
                    if (hrow_in) cwh_c = *reinterpret_cast<const unsigned*>(dg.code + off + hrow_d);
                    if (hcol_in) cb_c = dg.code[off + hcol_d];
                }
            }
            uint32_t pl_off = 0, pc_off = 3 * TB;              // byte offsets of planes k, k-1 in the ring
            for (int k = 0; k < P; ++k) {
                const bool owned = mine && k >= 1 && k <= P - 2;
                // the x values of the next owned plane travel while this plane is processed
                Quad<T> x_next{{T(0), T(0), T(0), T(0)}};
                if (apply_x && mine && k <= P - 3) x_next = has(4) ? ld4_hint<T>(xb + off + pstride, pol_of(4)) : ld4<T>(xb + off + pstride);
                unsigned cw_n = 0, cwh_n = 0, cb_n = 0;
                if constexpr (DIAG) {
                    if (k + 1 < P && x0 + k < nx) {       // plane x0 + k (the next one) lies inside the grid
                        if (mine) cw_n = *reinterpret_cast<const unsigned*>(dg.code + off + pstride);
                        if (hrow_in) cwh_n = *reinterpret_cast<const unsigned*>(dg.code + off + pstride + hrow_d);
                        if (hcol_in) cb_n = dg.code[off + pstride + hcol_d];
                    }
                }
                mbar_wait(smem_u32(&bars[stg]), (phase_bits >> stg) & 1u);
                phase_bits ^= 1u << stg;
                T(*rt)[SROW] = reinterpret_cast<T(*)[SROW]>(tiles0 + (size_t)(2 * stg) * TB);
                T(*dt)[SROW] = reinterpret_cast<T(*)[SROW]>(tiles0 + (size_t)(2 * stg + 1) * TB);
                T(*pl)[SROW] = reinterpret_cast<T(*)[SROW]>(planes0 + pl_off);
                stg = stg + 1 == kStages ? 0 : stg + 1;
                const Quad<T> rv = ld4<T>(&rt[sr][c]), dv = ld4<T>(&dt[sr][c]);
                Quad<T> dn;
                T hz_n = T(0);
                // diagonal and inverse diagonal of the own quad: interior rows (code 0) take the constants, the others ONE look-up each under a branch
                // that only boundary lanes enter; the diagonals travel to the next step's product in registers
                T ci_q[4] = {cinv0, cinv0, cinv0, cinv0}, c0_q[4] = {c00, c00, c00, c00};
                if constexpr (DIAG) {
                    if (cw_c != 0u) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const unsigned cd = (cw_c >> (8 * e)) & 0xffu;
                            c0_q[e] = s_c0[cd];
                            if constexpr (DINV) ci_q[e] = s_inv[cd];
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) dn.v[e] = dir(rv.v[e], dv.v[e], DINV ? ci_q[e] : cinv);
                st4(&pl[sr][c], dn);
                if (owned) {
                    if (has(6)) st4_hint(d_new + off, dn, pol_of(6));
                    else st4(d_new + off, dn);
                    if (apply_x) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) x_cur.v[e] = add_rn(x_cur.v[e], mul_rn(alpha, dv.v[e]));
                        if (has(4)) st4_hint(xb + off, x_cur, pol_of(4));
                        else st4(xb + off, x_cur);
                    }
                } else if ((k == 0 && halo_lo) || (k == P - 1 && halo_hi)) {
                    st4(d_new + off, dn);
                }
                if (warp < 2) {   // halo rows
                    const int hr = warp == 0 ? 0 : SY + 1;
                    const Quad<T> rh = ld4<T>(&rt[hr][c]), dh = ld4<T>(&dt[hr][c]);
                    Quad<T> hn;
                    T ci_h[4] = {cinv0, cinv0, cinv0, cinv0};
                    if constexpr (DINV) {
                        if (cwh_c != 0u) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) ci_h[e] = s_inv[(cwh_c >> (8 * e)) & 0xffu];
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) hn.v[e] = dir(rh.v[e], dh.v[e], DINV ? ci_h[e] : cinv);
                    st4(&pl[hr][c], hn);
                }
                if (lane == 0 || lane == 31) hz_n = dir(rt[sr][hcol], dt[sr][hcol], inv_of(cb_c, 0));   // halo columns of this row
                x_cur = x_next;
                __syncthreads();   // plane k complete; its stage fully read
                if (t == 0 && k + kStages < P) issue((int)x0 - 1 + k + kStages);
                if (k >= 2) {
                    T(*pc)[SROW] = reinterpret_cast<T(*)[SROW]>(planes0 + pc_off);
                    const Quad<T> ym = ld4<T>(&pc[sr - 1][c]), yp = ld4<T>(&pc[sr + 1][c]);
                    T left = __shfl_up_sync(0xffffffffu, dn_c.v[3], 1), right = __shfl_down_sync(0xffffffffu, dn_c.v[0], 1);
                    if (lane == 0) left = hz_c;
                    if (lane == 31) right = hz_c;
                    if (mine) {
                        const Quad<T>& ce = dn_c;
                        const Quad<T>& xm = dn_m;
                        const Quad<T>& xp = dn;
                        Quad<T> out;
                        T part = T(0), part_r = T(0);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const T zm = e == 0 ? left : ce.v[e - 1];
                            const T zp = e == 3 ? right : ce.v[e + 1];
                            const T c0 = DIAG ? c0_p[e] : c00;      // diagonal of this row (plane k-1)
                            T sum;
                            if constexpr (UNIT) {      // all six neighbour coefficients are -1: c*v == -v exactly, no multiply needed
                                sum = -xm.v[e];
                                sum = add_rn(sum, -ym.v[e]);
                                sum = add_rn(sum, -zm);
                                sum = add_rn(sum, mul_rn(c0, ce.v[e]));
                                sum = add_rn(sum, -zp);
                                sum = add_rn(sum, -yp.v[e]);
                                sum = add_rn(sum, -xp.v[e]);
                            } else {
                                sum = mul_rn(cxm, xm.v[e]);
                                sum = add_rn(sum, mul_rn(cym, ym.v[e]));
                                sum = add_rn(sum, mul_rn(czm, zm));
                                sum = add_rn(sum, mul_rn(c0, ce.v[e]));
                                sum = add_rn(sum, mul_rn(czp, zp));
                                sum = add_rn(sum, mul_rn(cyp, yp.v[e]));
                                sum = add_rn(sum, mul_rn(cxp, xp.v[e]));
                            }
                            out.v[e] = sum;
                            const T pr = mul_rn(ce.v[e], sum);
                            part = e == 0 ? pr : add_rn(part, pr);
                            if constexpr (ALG == 1) {
                                const T prr = mul_rn(ce.v[e], r_prev.v[e]);
                                part_r = e == 0 ? prr : add_rn(part_r, prr);
                            }
                        }
                        acc[0] += (double)part;        // the four products of a quad are added in T before they join the double sum
                        if constexpr (ALG == 1) acc[1] += (double)part_r;
                        if (has(8)) st4_hint(qb + off - pstride, out, pol_of(8));
                        else st4(qb + off - pstride, out);
                    }
                }
                off += pstride;
                r_prev = rv;
                dn_m = dn_c;
                dn_c = dn;
                hz_c = hz_n;
                cw_c = cw_n; cwh_c = cwh_n; cb_c = cb_n;
#pragma unroll
                for (int e = 0; e < 4; ++e) c0_p[e] = c0_q[e];
                pc_off = pl_off;   // rotate the ring: k -> k-1; the next plane takes the slot after k's (the slot of k-1 may still be read by slower warps)
                pl_off = pl_off + TB == 4 * TB ? 0 : pl_off + TB;
            }
        }
    }
    epilogue_tail(epi, acc, b, blockIdx.x, gridDim.x);
}


// Row sum of a star whose neighbours across a PERIODIC boundary sit at the other end of their axis: the CSR row is ordered by column, so a
// wrapped neighbour is added at its WRAPPED position.  Ascending offsets of everything a row can hold:
//   x+ wrapped, x-, y+ wrapped, y-, z+ wrapped, z-, centre, z+, z- wrapped, y+, y- wrapped, x+, x- wrapped
// wf bits: 0 x- wrapped, 1 x+ wrapped, 2 y- wrapped, 3 y+ wrapped, 4 z- wrapped, 5 z+ wrapped.  With wf == 0 this is the ordinary order.
template <typename T>
__device__ __forceinline__ T star_sum_wrapped(unsigned wf, T txm, T txp, T tym, T typ, T tzm, T tzp, T tc) {
    T sum = T(0);
    bool first = true;
    auto add = [&](T v) { sum = first ? v : add_rn(sum, v); first = false; };
    if (wf & 2u) add(txp);
    if (!(wf & 1u)) add(txm);
    if (wf & 8u) add(typ);
    if (!(wf & 4u)) add(tym);
    if (wf & 32u) add(tzp);
    if (!(wf & 16u)) add(tzm);
    add(tc);
    if (!(wf & 32u)) add(tzp);
    if (wf & 16u) add(tzm);
    if (!(wf & 8u)) add(typ);
    if (wf & 4u) add(tym);
    if (!(wf & 2u)) add(txp);
    if (wf & 1u) add(txm);
    return sum;
}

// ---------------------------------------------------------------------------------------------------------------------
// Plain product y = A x on a star stencil with TMA staging: the x tiles (with halo, zero-filled outside the domain) ARE the
// plane ring — kSpStages planes per CTA in flight, one barrier per plane, any epilogue (dots of BiCGstab, none for phb200_spmv).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSpStages = 5;
// WRAP: every stage also holds the rows / columns from the other end of a PERIODIC axis (filled by extra TMA boxes of the boundary tiles):
// [tile][y-low row][y-high row][z-low 4-column box][z-high 4-column box]
template <typename T> constexpr int wrap_row_bytes() { return ((SROW * (int)sizeof(T) + 127) / 128) * 128; }
template <typename T> constexpr int wrap_col_bytes() { return ((SROWS * 4 * (int)sizeof(T) + 127) / 128) * 128; }
template <typename T, bool WRAP> constexpr int spmv_stage_bytes() { return tile_bytes<T>() + (WRAP ? 2 * wrap_row_bytes<T>() + 2 * wrap_col_bytes<T>() : 0); }
template <typename T, bool WRAP = false> constexpr int star_spmv_smem() { return kSpStages * spmv_stage_bytes<T, WRAP>() + 8 * kSpStages + 128; }

// DIAG: the diagonal coefficient is looked up per row through the pattern code of a coded star (coded.cuh, CodedView::star_diag)
// WRAP (with DIAG): some axes are PERIODIC (st.wrap): the neighbour across the boundary is the cell at the other end of the axis
template <typename T, class Epi, bool UNIT, bool DIAG = false, bool WRAP = false>
__global__ void __launch_bounds__(kBlock) k_star_spmv_tma(const __grid_constant__ CUtensorMap tm_x, StarDesc st, T* __restrict__ Y, int64_t ldy, int xchunk, Epi epi,
                                                          StarDiag dg, const __grid_constant__ CUtensorMap tm_row, const __grid_constant__ CUtensorMap tm_col) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TB = spmv_stage_bytes<T, WRAP>();          // bytes per stage (tile + wrap buffers)
    constexpr int TILE = tile_bytes<T>(), WROW = wrap_row_bytes<T>(), WCOL = wrap_col_bytes<T>();
    (void)TILE; (void)WROW; (void)WCOL;
    constexpr int S = kSpStages;
    constexpr uint32_t kTx = (uint32_t)(SROWS * SROW * sizeof(T));
    // 128-byte aligned start inside the dynamic window, by pointer arithmetic on the shared array (an integer round trip makes
    // the compiler forget the address space: every tile access became a generic LD.E / ST.E with 64-bit address arithmetic)
    unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)S * TB);
    const int b = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int NDA = AccSize<Epi>::N;
    double acc[NDA];
#pragma unroll
    for (int d = 0; d < NDA; ++d) acc[d] = 0.0;
    __shared__ T s_c0[DIAG ? 256 : 1];
    if constexpr (DIAG) {
        if (t < dg.npat) s_c0[t] = ((const T*)dg.diag_tab)[t];
    }
    if (t == 0) {
#pragma unroll
        for (int k = 0; k < S; ++k) mbar_init(smem_u32(&bars[k]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    if (epi.idle()) return;
    if (!epi.skip(b)) {
        T* __restrict__ yb = Y + (int64_t)b * ldy;
        const int64_t ny = st.ny, nz = st.nz;
        const int tiles_z = (int)((nz + SZ - 1) / SZ), tiles_y = (int)((ny + SY - 1) / SY);
        const T c0 = (T)st.c0, cxm = (T)st.cxm, cxp = (T)st.cxp, cym = (T)st.cym, cyp = (T)st.cyp, czm = (T)st.czm, czp = (T)st.czp;
        const int64_t pstride = ny * nz;
        const int c = 4 + 4 * lane, sr = warp + 1;
        uint32_t stg = 0, phase_bits = 0, issue_stg = 0;
        SegIter seg(st, xchunk, tiles_y, tiles_z);
        while (seg.next()) {
            const int tz = seg.tz, ty = seg.ty;
            const int64_t x0 = seg.x0, x1 = seg.x1;
            const int y0 = ty * SY, z0 = tz * SZ;
            const int jy = y0 + warp, jz = z0 + 4 * lane;
            const bool mine = jy < ny && jz < nz;
            const int P = (int)(x1 - x0) + 2;   // planes x0-1 .. x1
            const int nxi = (int)st.nx;
            const bool wrap_x = WRAP && (st.wrap & 1), wrap_y = WRAP && (st.wrap & 2), wrap_z = WRAP && (st.wrap & 4);
            // PERIODIC y / z: tiles at the boundary of such an axis also fetch the row / the 4-column box from the other end of the axis into
            // the stage's wrap buffers (the main tile holds zeros there), with the same TMA transaction barrier
            const bool t_ylo = wrap_y && y0 == 0, t_yhi = wrap_y && y0 + SY >= ny, t_zlo = wrap_z && z0 == 0, t_zhi = wrap_z && z0 + SZ >= nz;
            const uint32_t tx_bytes = kTx + (WRAP ? ((t_ylo ? 1u : 0u) + (t_yhi ? 1u : 0u)) * (uint32_t)(SROW * sizeof(T)) +
                                                        ((t_zlo ? 1u : 0u) + (t_zhi ? 1u : 0u)) * (uint32_t)(SROWS * 4 * sizeof(T)) : 0u);
            auto issue = [&](int px) {
                const uint32_t bar = smem_u32(&bars[issue_stg]);
                mbar_expect_tx(bar, tx_bytes);
                if (wrap_x) px = px < 0 ? px + nxi : (px >= nxi ? px - nxi : px);      // PERIODIC x: the plane before the first is the last one
                unsigned char* stage = base + (size_t)issue_stg * TB;
                tma_load_4d(smem_u32(stage), &tm_x, bar, z0 - 4, y0 - 1, px, b);
                if constexpr (WRAP) {
                    if (t_ylo) tma_load_4d(smem_u32(stage + TILE), &tm_row, bar, z0 - 4, (int)ny - 1, px, b);
                    if (t_yhi) tma_load_4d(smem_u32(stage + TILE + WROW), &tm_row, bar, z0 - 4, 0, px, b);
                    if (t_zlo) tma_load_4d(smem_u32(stage + TILE + 2 * WROW), &tm_col, bar, (int)nz - 4, y0 - 1, px, b);
                    if (t_zhi) tma_load_4d(smem_u32(stage + TILE + 2 * WROW + WCOL), &tm_col, bar, 0, y0 - 1, px, b);
                }
                issue_stg = issue_stg + 1 == S ? 0 : issue_stg + 1;
            };
            const bool wy_lo = wrap_y && mine && jy == 0, wy_hi = wrap_y && mine && jy == ny - 1;
            const bool wz_lo = wrap_z && mine && jz == 0, wz_hi = wrap_z && mine && jz + 4 == nz;
            __syncthreads();   // every stage of the previous item has been consumed
            if (t == 0)
                for (int k = 0; k < S && k < P; ++k) issue((int)x0 - 1 + k);
            int64_t off = (x0 * ny + jy) * nz + jz;              // element offset of the first OUTPUT plane
            uint32_t s_c = 0;                                    // stage of plane k-1
            // every thread keeps its own quad of the last three planes in registers (a plane's quad is read from shared memory once, when
            // the plane arrives); only the y neighbours (other warps' rows) and the two halo columns of a row come from the staged tiles.
            // z neighbours inside the row travel by warp shuffle (the scalar loads at c-1 / c+4 were 4-way bank conflicts).
            Quad<T> q_m{{T(0), T(0), T(0), T(0)}}, q_c{{T(0), T(0), T(0), T(0)}};
            const int hcol = lane == 0 ? 3 : 4 + SZ;             // halo column read by the first / last lane of a row
            unsigned cw_c = 0;                                   // DIAG: pattern codes of this step's output quad, fetched one step ahead
            const T c00 = DIAG ? s_c0[0] : c0;
            for (int k = 0; k < P; ++k) {
                unsigned cw_n = 0;
                if constexpr (DIAG) {
                    if (mine && k >= 1 && k + 1 < P) cw_n = *reinterpret_cast<const unsigned*>(dg.code + off + (k >= 2 ? pstride : 0));
                }
                mbar_wait(smem_u32(&bars[stg]), (phase_bits >> stg) & 1u);
                phase_bits ^= 1u << stg;
                const uint32_t s_p = stg;
                stg = stg + 1 == S ? 0 : stg + 1;
                T(*pp)[SROW] = reinterpret_cast<T(*)[SROW]>(base + (size_t)s_p * TB);
                const Quad<T> q_p = ld4<T>(&pp[sr][c]);
                if (k >= 2) {
                    T(*pc)[SROW] = reinterpret_cast<T(*)[SROW]>(base + (size_t)s_c * TB);
                    const Quad<T> ym = ld4<T>(&pc[sr - 1][c]), yp = ld4<T>(&pc[sr + 1][c]);
                    T left = __shfl_up_sync(0xffffffffu, q_c.v[3], 1), right = __shfl_down_sync(0xffffffffu, q_c.v[0], 1);
                    if (lane == 0 || lane == 31) {
                        const T h = pc[sr][hcol];
                        if (lane == 0) left = h;
                        else right = h;
                    }
                    if (mine) {
                    const Quad<T>& ce = q_c;
                    const Quad<T>& xm = q_m;
                    const Quad<T>& xp = q_p;
                    Quad<T> out;
                    unsigned wf_row = 0;                   // WRAP: wrapped neighbours of this step's output rows (see star_sum_wrapped)
                    Quad<T> ymw = ym, ypw = yp;
                    if constexpr (WRAP) {
                        const int xo = (int)x0 + k - 2;    // output plane
                        if (wrap_x && xo == 0) wf_row |= 1u;
                        if (wrap_x && xo == nxi - 1) wf_row |= 2u;
                        const unsigned char* stage_c = base + (size_t)s_c * TB;      // wrap buffers of the product plane
                        if (wy_lo) { wf_row |= 4u; ymw = ld4<T>(reinterpret_cast<const T*>(stage_c + TILE) + c); }
                        if (wy_hi) { wf_row |= 8u; ypw = ld4<T>(reinterpret_cast<const T*>(stage_c + TILE + WROW) + c); }
                        if (wz_lo) left = reinterpret_cast<const T*>(stage_c + TILE + 2 * WROW)[sr * 4 + 3];
                        if (wz_hi) right = reinterpret_cast<const T*>(stage_c + TILE + 2 * WROW + WCOL)[sr * 4];
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const T zm = e == 0 ? left : ce.v[e - 1];
                        const T zp = e == 3 ? right : ce.v[e + 1];
                        const T c0 = DIAG ? (cw_c == 0u ? c00 : s_c0[(cw_c >> (8 * e)) & 0xffu]) : c00;      // diagonal of this row
                        T sum;
                        // the seven terms (UNIT: the six neighbour coefficients are -1, c*v == -v exactly), then their sum in ascending column order
                        const T txm = UNIT ? -xm.v[e] : mul_rn(cxm, xm.v[e]), txp = UNIT ? -xp.v[e] : mul_rn(cxp, xp.v[e]);
                        const T tym = UNIT ? -ymw.v[e] : mul_rn(cym, ymw.v[e]), typ = UNIT ? -ypw.v[e] : mul_rn(cyp, ypw.v[e]);
                        const T tzm = UNIT ? -zm : mul_rn(czm, zm), tzp = UNIT ? -zp : mul_rn(czp, zp), tc = mul_rn(c0, ce.v[e]);
                        if (WRAP && wf_row != 0u) {           // rows on a PERIODIC x / y boundary (uniform over the warp, rare): general order
                            const unsigned wf = wf_row | ((wz_lo && e == 0) ? 16u : 0u) | ((wz_hi && e == 3) ? 32u : 0u);
                            sum = star_sum_wrapped<T>(wf, txm, txp, tym, typ, tzm, tzp, tc);
                        } else {
                            // a z neighbour across a PERIODIC boundary changes the order of the three middle terms only (selects, no divergence):
                            // z- wrapped (first cell of a row): centre, z+, z- ; z+ wrapped (last cell): z+, z-, centre
                            T a1 = tzm, a2 = tc, a3 = tzp;
                            if constexpr (WRAP) {
                                const bool wl = wz_lo && e == 0, wh = wz_hi && e == 3;
                                a1 = wl ? tc : (wh ? tzp : tzm);
                                a2 = wl ? tzp : (wh ? tzm : tc);
                                a3 = wl ? tzm : (wh ? tc : tzp);
                            }
                            sum = add_rn(txm, tym);
                            sum = add_rn(sum, a1);
                            sum = add_rn(sum, a2);
                            sum = add_rn(sum, a3);
                            sum = add_rn(sum, typ);
                            sum = add_rn(sum, txp);
                        }
                        out.v[e] = sum;
                    }
                    st4(yb + off, out);
#pragma unroll
                    for (int e = 0; e < 4; ++e) epi.row(b, off + e, out.v[e], acc);
                    off += pstride;
                    }
                }
                q_m = q_c;
                q_c = q_p;
                cw_c = cw_n;
                __syncthreads();   // the stage of plane k-1 is free (its rows were read as y neighbours above); plane k stays for the next step
                // planes k+1 .. k+S-2 are in flight; the freed stage takes plane k+S-1
                if (t == 0 && k >= 1 && k + S - 1 < P) issue((int)x0 - 1 + k + S - 1);
                s_c = s_p;
            }
        }
    }
    epilogue_tail(epi, acc, b, blockIdx.x, gridDim.x);
}

template <typename T, class Epi, bool UNIT, bool DIAG, bool WRAP = false>
int launch_star_spmv_tma_t(const CUtensorMap& tm, const StarDesc& st, T* Y, int64_t ldy, int batch, const Epi& epi, cudaStream_t stream, const StarDiag& dg,
                           const CUtensorMap* tm_row = nullptr, const CUtensorMap* tm_col = nullptr) {
    auto kernel = k_star_spmv_tma<T, Epi, UNIT, DIAG, WRAP>;
    constexpr int smem = star_spmv_smem<T, WRAP>();
    static int occ_dev[kMaxDevices] = {};
    int& occ = occ_dev[device_slot()];
    if (occ == 0) {
        PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kBlock, smem) != cudaSuccess || v < 1) { cudaGetLastError(); v = 1; }
        occ = v;
    }
    dim3 grid;
    int xchunk;
    star_grid_balanced(st, batch, occ, grid, xchunk);
    PHB_LAUNCH(kernel, grid, kBlock, smem, stream, tm, st, Y, ldy, xchunk, epi, dg, tm_row ? *tm_row : tm, tm_col ? *tm_col : tm);
    return PHB200_OK;
}

// coded star with a per-row diagonal (dg != nullptr): the same kernel in its DIAG form
template <typename T, class Epi>
int launch_star_spmv_tma_diag(const StarDesc& st, StarDiag dg, const T* X, int64_t ldx, T* Y, int64_t ldy, int batch, int dtype, const Epi& epi, cudaStream_t stream, bool* used) {
    *used = false;
    dg.x = X;
    dg.ldx = ldx;
    CUtensorMap tm;
    if (!make_star_tmap(&tm, X, dtype, st, batch, ldx)) return PHB200_OK;
    const bool unit = st.cxm == -1.0 && st.cxp == -1.0 && st.cym == -1.0 && st.cyp == -1.0 && st.czm == -1.0 && st.czp == -1.0;
    *used = true;
    if (st.wrap) {
        CUtensorMap tm_row, tm_col;      // boxes of one row / of a 4-column strip: the neighbours across a PERIODIC boundary
        if (!make_star_tmap(&tm_row, X, dtype, st, batch, ldx, SROW, 1) || !make_star_tmap(&tm_col, X, dtype, st, batch, ldx, 4, SROWS)) { *used = false; return PHB200_OK; }
        return unit ? launch_star_spmv_tma_t<T, Epi, true, true, true>(tm, st, Y, ldy, batch, epi, stream, dg, &tm_row, &tm_col)
                    : launch_star_spmv_tma_t<T, Epi, false, true, true>(tm, st, Y, ldy, batch, epi, stream, dg, &tm_row, &tm_col);
    }
    return unit ? launch_star_spmv_tma_t<T, Epi, true, true>(tm, st, Y, ldy, batch, epi, stream, dg) : launch_star_spmv_tma_t<T, Epi, false, true>(tm, st, Y, ldy, batch, epi, stream, dg);
}

template <typename T, class Epi>
int launch_star_spmv_tma(const StarDesc& st, const T* X, int64_t ldx, T* Y, int64_t ldy, int batch, int dtype, const Epi& epi, cudaStream_t stream, bool* used) {
    *used = false;
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("PHB200_SPMV_TMA"); disabled = (e && e[0] == '0') ? 1 : 0; }
    if (disabled) return PHB200_OK;
    CUtensorMap tm;
    if (!make_star_tmap(&tm, X, dtype, st, batch, ldx)) return PHB200_OK;
    const bool unit = st.cxm == -1.0 && st.cxp == -1.0 && st.cym == -1.0 && st.cyp == -1.0 && st.czm == -1.0 && st.czp == -1.0;
    *used = true;
    return unit ? launch_star_spmv_tma_t<T, Epi, true, false>(tm, st, Y, ldy, batch, epi, stream, StarDiag{}) : launch_star_spmv_tma_t<T, Epi, false, false>(tm, st, Y, ldy, batch, epi, stream, StarDiag{});
}

}  // namespace phb
