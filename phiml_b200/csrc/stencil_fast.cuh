// Stencil-specialised kernels for the "star" stencils matrix_from_function emits for Laplace / central differences:
// centre + nearest neighbours along each axis (5-point in 2-D, 7-point in 3-D), constant coefficients, ZERO boundaries.
//
// One CTA owns a (SY x SZ) tile of the y-z plane and marches along x over a chunk of planes, keeping three planes of the
// input vector (tile + one halo row/column) in shared memory, so every input element is read from global memory once
// per CTA (plus the halo) with 128-bit loads; x-neighbours come from the rotating plane buffers, y/z-neighbours from the
// current plane.  Row sums are accumulated in ascending column order with separately rounded products, i.e. exactly like
// the canonical CSR row (out-of-domain neighbours contribute a signed zero, which leaves the sum untouched).
//
// Two input policies:
//   PlainLoad  — v = x (plain product, any epilogue: dots, scalar steps, convergence test)
//   CgDirLoad  — v = d_new = M^-1 r + beta d_old computed on the fly (and stored once), fused with x += alpha_prev d_old:
//                the "direction" and "product" phases of a CG iteration in one pass over memory.
#pragma once
#include "phase.cuh"

namespace phb {

constexpr int SY = 8;            // tile rows (one per warp)
constexpr int SZ = 128;          // tile columns (4 per lane)
constexpr int SROW = SZ + 8;     // smem row: [3 pad][left halo][SZ][right halo][3 pad] -> interior 16-byte aligned
constexpr int SROWS = SY + 2;

// host: is this stencil a star with nz % 4 == 0 (128-bit rows)?
inline bool star_from_stencil(const StencilDesc& st, StarDesc& out) {
    out.nx = st.dims[0]; out.ny = st.dims[1]; out.nz = st.dims[2];
    out.xb = 0; out.xe = out.nx; out.store_halo = 0;
    out.c0 = out.cxm = out.cxp = out.cym = out.cyp = out.czm = out.czp = 0.0;
    if (out.nz % 4 != 0) return false;
    for (int p = 0; p < st.npts; ++p) {
        const int ox = st.off[p][0], oy = st.off[p][1], oz = st.off[p][2];
        const int nzr = (ox != 0) + (oy != 0) + (oz != 0);
        if (nzr > 1) return false;
        if (nzr == 0) out.c0 = st.coef[p];
        else if (ox == -1) out.cxm = st.coef[p];
        else if (ox == 1) out.cxp = st.coef[p];
        else if (oy == -1) out.cym = st.coef[p];
        else if (oy == 1) out.cyp = st.coef[p];
        else if (oz == -1) out.czm = st.coef[p];
        else if (oz == 1) out.czp = st.coef[p];
        else return false;
    }
    return true;
}

template <typename T> struct Quad { T v[4]; };

template <typename T> __device__ __forceinline__ Quad<T> ld4(const T* p);
template <> __device__ __forceinline__ Quad<float> ld4<float>(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    return Quad<float>{{t.x, t.y, t.z, t.w}};
}
template <> __device__ __forceinline__ Quad<double> ld4<double>(const double* p) {
    double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    return Quad<double>{{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ void st4(float* p, const Quad<float>& q) { *reinterpret_cast<float4*>(p) = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]); }
__device__ __forceinline__ void st4(double* p, const Quad<double>& q) {
    *reinterpret_cast<double2*>(p) = make_double2(q.v[0], q.v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(q.v[2], q.v[3]);
}

// ------------------------------------------------------------------------------------------------ input policies
template <typename T>
struct PlainLoad {
    const T* x;
    int64_t ld;
    struct Ctx { const T* x; };
    __device__ Ctx begin(int b, const Sys*) const { return Ctx{x + (int64_t)b * ld}; }
    __device__ Quad<T> load4(const Ctx& c, int64_t i, bool) const { return ld4<T>(c.x + i); }
    __device__ T load1(const Ctx& c, int64_t i) const { return c.x[i]; }
};

// PRE: 0 none, 1 Jacobi with inv_diag vector, 3 Jacobi with constant inverse diagonal
template <typename T, int PRE>
struct CgDirLoad {
    const T* r;
    T* buf_a;          // direction buffers: the current direction lives in buf_b after an odd number of executed passes
    T* buf_b;
    T* xvec;
    const T* w;
    T cinv;
    int64_t ld;
    const Glob* glob;
    struct Ctx { const T* r; const T* d_old; T* d_new; T* x; T beta, alpha; };
    __device__ Ctx begin(int b, const Sys* sys) const {
        const int64_t o = (int64_t)b * ld;
        const bool odd = (glob->pass & 1) != 0;
        return Ctx{r + o, (odd ? buf_b : buf_a) + o, (odd ? buf_a : buf_b) + o, xvec + o, (T)sys[b].s[1 /*S_BETA*/], (T)sys[b].s[0 /*S_ALPHA*/]};
    }
    __device__ __forceinline__ T dir(T rv, T dv, T wv, T beta) const {
        const T s = PRE == 1 ? mul_rn(rv, wv) : (PRE == 3 ? mul_rn(rv, cinv) : rv);
        return add_rn(s, mul_rn(beta, dv));
    }
    __device__ Quad<T> load4(const Ctx& c, int64_t i, bool owned) const {
        const Quad<T> rv = ld4<T>(c.r + i), dv = ld4<T>(c.d_old + i);
        Quad<T> wv;
        if (PRE == 1) wv = ld4<T>(w + i);
        Quad<T> out;
#pragma unroll
        for (int k = 0; k < 4; ++k) out.v[k] = dir(rv.v[k], dv.v[k], PRE == 1 ? wv.v[k] : T(0), c.beta);
        if (owned) {
            st4(c.d_new + i, out);
            Quad<T> xv = ld4<T>(c.x + i);
#pragma unroll
            for (int k = 0; k < 4; ++k) xv.v[k] = add_rn(xv.v[k], mul_rn(c.alpha, dv.v[k]));
            st4(c.x + i, xv);
        }
        return out;
    }
    __device__ T load1(const Ctx& c, int64_t i) const { return dir(c.r[i], c.d_old[i], PRE == 1 ? w[i] : T(0), c.beta); }
};

// ------------------------------------------------------------------------------------------------ the kernel
// SELF_DOT: acc[0] += v_i * y_i with v the loaded input (d.q of CG) instead of calling epi.row per element.
template <typename T, class Loader, class Epi, bool SELF_DOT>
__global__ void __launch_bounds__(kBlock) k_star(StarDesc st, Loader loader, T* __restrict__ Y, int64_t ldy, int xchunk, Epi epi) {
    if (epi.idle()) return;
    __shared__ __align__(16) T s[3][SROWS][SROW];
    const int b = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int ND = AccSize<Epi>::N;
    double acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[d] = 0.0;
    if (!epi.skip(b)) {
        const typename Loader::Ctx lc = loader.begin(b, epi.sys);
        const int64_t nx = st.nx, ny = st.ny, nz = st.nz;
        const int tiles_z = (int)((nz + SZ - 1) / SZ), tiles_y = (int)((ny + SY - 1) / SY);
        const int chunks = (int)((st.xe - st.xb + xchunk - 1) / xchunk);
        const int64_t items = (int64_t)chunks * tiles_y * tiles_z;
        const T c0 = (T)st.c0, cxm = (T)st.cxm, cxp = (T)st.cxp, cym = (T)st.cym, cyp = (T)st.cyp, czm = (T)st.czm, czp = (T)st.czp;
        T* __restrict__ yb = Y + (int64_t)b * ldy;

        for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
            const int tz = (int)(item % tiles_z), ty = (int)((item / tiles_z) % tiles_y), ch = (int)(item / ((int64_t)tiles_z * tiles_y));
            const int64_t x0 = st.xb + (int64_t)ch * xchunk, x1 = min(st.xe, x0 + xchunk);
            const int64_t y0 = (int64_t)ty * SY, z0 = (int64_t)tz * SZ;
            const int64_t jz = z0 + 4 * lane;

            // fills buffer `slot` with plane px of the input (tile + halo); out-of-domain -> 0
            auto fill = [&](int slot, int64_t px, bool owned_plane) {
                const bool in_x = px >= 0 && px < nx;
                // interior rows: warp w -> smem row w+1
                {
                    const int64_t jy = y0 + warp;
                    Quad<T> q{{T(0), T(0), T(0), T(0)}};
                    if (in_x && jy < ny && jz < nz) q = loader.load4(lc, (px * ny + jy) * nz + jz, owned_plane);
                    st4(&s[slot][warp + 1][4 + 4 * lane], q);
                }
                // halo rows: warps 0/1 -> smem rows 0 / SY+1
                if (warp < 2) {
                    const int64_t jy = warp == 0 ? y0 - 1 : y0 + SY;
                    Quad<T> q{{T(0), T(0), T(0), T(0)}};
                    if (in_x && jy >= 0 && jy < ny && jz < nz) q = loader.load4(lc, (px * ny + jy) * nz + jz, false);
                    st4(&s[slot][warp == 0 ? 0 : SY + 1][4 + 4 * lane], q);
                }
                // halo columns: 2 per smem row
                if (t >= 64 && t < 64 + 2 * SROWS) {
                    const int row = (t - 64) >> 1, side = (t - 64) & 1;
                    const int64_t jy = y0 - 1 + row, zz = side ? z0 + SZ : z0 - 1;
                    T v = T(0);
                    if (in_x && jy >= 0 && jy < ny && zz >= 0 && zz < nz) v = loader.load1(lc, (px * ny + jy) * nz + zz);
                    s[slot][row][side ? 4 + SZ : 3] = v;
                }
            };

            __syncthreads();   // buffers of the previous item are no longer read
            fill(0, x0 - 1, false);
            fill(1, x0, true);
            int prev = 0, cur = 1, next = 2;
            for (int64_t x = x0; x < x1; ++x) {
                fill(next, x + 1, x + 1 < x1);
                __syncthreads();
                const int64_t jy = y0 + warp;
                if (jy < ny && jz < nz) {
                    const int sr = warp + 1, c = 4 + 4 * lane;
                    const Quad<T> ce = ld4<T>(&s[cur][sr][c]);
                    const Quad<T> ym = ld4<T>(&s[cur][sr - 1][c]), yp = ld4<T>(&s[cur][sr + 1][c]);
                    const Quad<T> xm = ld4<T>(&s[prev][sr][c]), xp = ld4<T>(&s[next][sr][c]);
                    const T left = s[cur][sr][c - 1], right = s[cur][sr][c + 4];
                    Quad<T> out;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const T zm = k == 0 ? left : ce.v[k - 1];
                        const T zp = k == 3 ? right : ce.v[k + 1];
                        T sum = mul_rn(cxm, xm.v[k]);
                        sum = add_rn(sum, mul_rn(cym, ym.v[k]));
                        sum = add_rn(sum, mul_rn(czm, zm));
                        sum = add_rn(sum, mul_rn(c0, ce.v[k]));
                        sum = add_rn(sum, mul_rn(czp, zp));
                        sum = add_rn(sum, mul_rn(cyp, yp.v[k]));
                        sum = add_rn(sum, mul_rn(cxp, xp.v[k]));
                        out.v[k] = sum;
                    }
                    const int64_t i = (x * ny + jy) * nz + jz;
                    st4(yb + i, out);
                    if constexpr (SELF_DOT) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[0] += (double)mul_rn(ce.v[k], out.v[k]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) epi.row(b, i + k, out.v[k], acc);
                    }
                }
                __syncthreads();
                const int tmp = prev; prev = cur; cur = next; next = tmp;
            }
        }
    }
    epilogue_tail(epi, acc, b, blockIdx.x, gridDim.x);
}

// grid for k_star: one item per CTA when everything fits in one resident wave, else grid-stride over items
inline void star_grid(const StarDesc& st, int batch, int occ, dim3& grid, int& xchunk) {
    const int64_t tiles = ((st.ny + SY - 1) / SY) * ((st.nz + SZ - 1) / SZ);
    int64_t resident = ((int64_t)kSMs * (occ < 1 ? 1 : (occ > 8 ? 8 : occ))) / (batch < 1 ? 1 : batch);
    if (resident < 1) resident = 1;
    const int64_t planes = st.xe - st.xb;
    int64_t chunks = resident / tiles;
    static int max_chunks = -1;       // PHB200_STAR_CHUNKS: upper bound on the number of x-chunks (fewer chunks = fewer DRAM fronts)
    if (max_chunks < 0) { const char* e = getenv("PHB200_STAR_CHUNKS"); max_chunks = e ? atoi(e) : 0; }
    if (max_chunks > 0 && chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    if (chunks > planes) chunks = planes;
    xchunk = (int)((planes + chunks - 1) / chunks);
    if (xchunk < 1) xchunk = 1;
    chunks = (planes + xchunk - 1) / xchunk;
    int64_t items = chunks * tiles;
    int64_t gx = items < resident ? items : resident;
    if (gx < 1) gx = 1;
    grid = dim3((unsigned)gx, (unsigned)batch);
}

// Balanced decomposition for the TMA kernels (SegIter in star_tma.cuh): every resident CTA slot gets an equal share of the
// tiles x planes space; xchunk = 0 tells the kernel.  Off by default (PHB200_STAR_BALANCED=1 enables it): measured SLOWER than
// equal x-chunks marched in lockstep (256^3: 81 vs 74 us, 512^3: 693 vs 608 us) although it fills all 444 CTA slots instead of 384 / 256 —
// neighbouring tiles no longer sit on the same x-plane at the same time, so halo rows miss L2 and DRAM sees many more open fronts.
inline void star_grid_balanced(const StarDesc& st, int batch, int occ, dim3& grid, int& xchunk) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("PHB200_STAR_BALANCED"); enabled = (e && e[0] == '1') ? 1 : 0; }
    if (!enabled) { star_grid(st, batch, occ, grid, xchunk); return; }
    const int64_t tiles = ((st.ny + SY - 1) / SY) * ((st.nz + SZ - 1) / SZ);
    int64_t resident = ((int64_t)kSMs * (occ < 1 ? 1 : (occ > 8 ? 8 : occ))) / (batch < 1 ? 1 : batch);
    if (resident < 1) resident = 1;
    const int64_t work = tiles * (st.xe - st.xb);
    // a range shorter than ~4 planes is mostly halo planes: use fewer CTAs then
    int64_t gx = resident;
    if (gx > work / 4) gx = work / 4;
    if (gx < 1) gx = 1;
    xchunk = 0;
    grid = dim3((unsigned)gx, (unsigned)batch);
}

template <typename T, class Loader, class Epi, bool SELF_DOT>
int launch_star(const StarDesc& st, const Loader& loader, T* Y, int64_t ldy, int batch, const Epi& epi, cudaStream_t stream) {
    dim3 grid;
    int xchunk;
    star_grid(st, batch, occupancy(k_star<T, Loader, Epi, SELF_DOT>), grid, xchunk);
    PHB_LAUNCH((k_star<T, Loader, Epi, SELF_DOT>), grid, kBlock, 0, stream, st, loader, Y, ldy, xchunk, epi);
    return PHB200_OK;
}

}  // namespace phb
