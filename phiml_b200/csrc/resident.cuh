// On-chip ("resident") CG for systems that fit into the shared memory + registers of one CTA or one thread-block cluster:
// the whole Krylov loop of one system — direction, stencil product, both dot products, x/r update, convergence test —
// runs inside one kernel with r in registers and x, the direction and q = A d in shared memory.  HBM is touched once per
// solve (load x0, r0, d0; store x, r, d), not once per iteration: the batched small-system case (BASELINE config 3,
// 4096 x 128^2) and the single small system (config 1) are bound by on-chip throughput / latency instead of HBM.
//
// Layout: the grid's lines (nz contiguous elements; nx*ny of them) are dealt out contiguously to the CTAs of a cluster
// (`lpc` lines per CTA); y/x neighbours that live in another CTA are read through distributed shared memory
// (mapa + ld.shared::cluster, 32-bit addresses precomputed per thread).  A thread owns kResQpt packs of 16 bytes
// (4 floats / 2 doubles); pack k of thread t is pack number k*1024+t of the CTA, so a warp always touches 512
// contiguous bytes of shared memory (conflict-free 128-bit accesses).  Every reduction is ONE barrier: warp butterflies
// -> 32 partials in shared memory -> barrier -> every warp adds the partials of all CTAs in rank order (deterministic,
// identical in every thread, no broadcast needed).
//
// Arithmetic follows the streaming kernels operation for operation (separately rounded products, same stencil summation
// order); only the summation tree of the dot products differs (fp32: the products of one pack are added in fp32 before
// they join the double accumulator).
// Algorithms: `cg` phiml/backend/_linalg.py:52-90 with stop_on_l2 :23-40 (rules 0) and torch_sparse_cg
// phiml/backend/torch/_torch_backend.py:1353-1382 (rules 1: refresh every 20th pass, finished systems keep going until
// all are finished — see run_resident in solver.cu for how that coupling is resolved).
#pragma once
#include <cooperative_groups.h>
#include "stencil_fast.cuh"

namespace phb {

namespace cg = cooperative_groups;

constexpr int kResThreads = 1024;
constexpr int kResMaxCluster = 8;
constexpr int kResQpt = 4;         // 16-byte packs per thread

template <typename T>
struct ResCoef {                   // already rounded to T: used straight from the constant bank
    T c0, cxm, cxp, cym, cyp, czm, czp, cinv;
};

struct ResArgs {
    int nx, ny, nz;
    void* X;
    void* R;
    void* D;
    void* Q;       // q = A d (adaptive variant: carried across launches)
    const void* Y;
    int64_t ld;
    int batch;
    Sys* sys;
    const Glob* glob;
    int torch_rules;
    int lb;        // a system stops at the first pass >= lb after which it does not continue ...
    int cap;       // ... or when it has made `cap` passes
    int lpc;       // lines per CTA
    int cs;        // CTAs per cluster
};

template <typename T> constexpr int res_pack() { return 16 / (int)sizeof(T); }
template <typename T> inline size_t res_smem_bytes(int lpc, int nz) { return (size_t)3 * lpc * nz * sizeof(T) + 6 * 32 * sizeof(double) + 16; }

template <typename T> struct Pack;
template <> struct Pack<float> { float v[4]; };
template <> struct Pack<double> { double v[2]; };

// 32-bit shared-window addresses: plain (own CTA) and cluster-wide
__device__ __forceinline__ Pack<float> lds_pack(uint32_t a, float) {
    Pack<float> p;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]) : "r"(a) : "memory");
    return p;
}
__device__ __forceinline__ Pack<double> lds_pack(uint32_t a, double) {
    Pack<double> p;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(p.v[0]), "=d"(p.v[1]) : "r"(a) : "memory");
    return p;
}
__device__ __forceinline__ Pack<float> ldc_pack(uint32_t a, float) {
    Pack<float> p;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]) : "r"(a) : "memory");
    return p;
}
__device__ __forceinline__ Pack<double> ldc_pack(uint32_t a, double) {
    Pack<double> p;
    asm volatile("ld.shared::cluster.v2.f64 {%0,%1}, [%2];" : "=d"(p.v[0]), "=d"(p.v[1]) : "r"(a) : "memory");
    return p;
}
__device__ __forceinline__ void sts_pack(uint32_t a, const Pack<float>& p) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(p.v[0]), "f"(p.v[1]), "f"(p.v[2]), "f"(p.v[3]) : "memory");
}
__device__ __forceinline__ void sts_pack(uint32_t a, const Pack<double>& p) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(p.v[0]), "d"(p.v[1]) : "memory");
}
__device__ __forceinline__ float lds_one(uint32_t a, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds_one(uint32_t a, double) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double ldc_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
template <typename T> __device__ __forceinline__ Pack<T> ldg_pack(const T* p);
template <> __device__ __forceinline__ Pack<float> ldg_pack<float>(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    return Pack<float>{{t.x, t.y, t.z, t.w}};
}
template <> __device__ __forceinline__ Pack<double> ldg_pack<double>(const double* p) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    return Pack<double>{{t.x, t.y}};
}
__device__ __forceinline__ void stg_pack(float* p, const Pack<float>& q) { *reinterpret_cast<float4*>(p) = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]); }
__device__ __forceinline__ void stg_pack(double* p, const Pack<double>& q) { *reinterpret_cast<double2*>(p) = make_double2(q.v[0], q.v[1]); }

__device__ __forceinline__ double butterfly_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// flags of a pack
constexpr unsigned RF_VALID = 1u, RF_YM = 2u, RF_YP = 4u, RF_XM = 8u, RF_XP = 16u, RF_Z0 = 32u, RF_ZN = 64u, RF_LO = 128u /* first */, RF_HI = 256u /* last line of the CTA */;

// ALG 0: cg / torch_sparse_cg;  ALG 1: cg_adaptive / torch_sparse_cg_adaptive (phiml/backend/_linalg.py:93-128,
// phiml/backend/torch/_torch_backend.py:1385-1413; no preconditioner)
template <typename T, int PRE, bool RANK3, bool CLUSTER, int ALG>
__global__ void __launch_bounds__(kResThreads, 1) k_cg_resident(const ResArgs a, const ResCoef<T> cf) {
    constexpr int V = res_pack<T>();
    constexpr int QPT = kResQpt;
    constexpr uint32_t W = sizeof(T);
    extern __shared__ __align__(16) unsigned char res_smem[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nz = a.nz, ny = a.ny, nx = a.nx;
    const int lines = nx * ny;
    int rank = 0, cs = 1;
    if constexpr (CLUSTER) {
        rank = (int)cg::this_cluster().block_rank();
        cs = a.cs;
    }
    const int lpc = a.lpc;
    const int l0 = rank * lpc;
    const int nl = max(0, min(lpc, lines - l0));
    const uint32_t vec_bytes = (uint32_t)lpc * (uint32_t)nz * W;
    const uint32_t d_a = (uint32_t)__cvta_generic_to_shared(res_smem);   // direction
    const uint32_t q_off = vec_bytes, x_off = 2 * vec_bytes;              // q and x relative to the direction
    const uint32_t red_a = d_a + 3 * vec_bytes;
    T* d_s = reinterpret_cast<T*>(res_smem);
    for (int i = t; i < lpc * nz; i += kResThreads) d_s[i] = T(0);
    auto barrier = [&]() {
        if constexpr (CLUSTER) cg::this_cluster().sync();
        else __syncthreads();
    };
    barrier();

    // ---- static element map: pack k of this thread -> own address, flags (+ z), x-neighbour addresses ----
    // y neighbours are one line away: own -/+ nz*W unless that line belongs to the previous / next CTA of the cluster
    const int npacks = nl * (nz / V);
    const uint32_t line_bytes = (uint32_t)nz * W;
    uint32_t own[QPT], a_xm[RANK3 ? QPT : 1], a_xp[RANK3 ? QPT : 1];
    unsigned fl[QPT];       // RF_* bits, z * W in bits 9..31
    uint32_t rem_lo = d_a, rem_hi = d_a;   // last line of the previous CTA / first line of the next one (cluster addresses)
    if constexpr (CLUSTER) {
        if (rank > 0) rem_lo = mapa_u32(d_a + (uint32_t)(lpc - 1) * line_bytes, (uint32_t)(rank - 1));
        if (rank + 1 < cs) rem_hi = mapa_u32(d_a, (uint32_t)(rank + 1));
    }
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        const int g = k * kResThreads + t;
        const bool valid = g < npacks;
        const int e = valid ? V * g : 0;
        const int l = e / nz, z = e - l * nz;
        own[k] = d_a + (uint32_t)e * W;
        const int L = l0 + l;
        const int ix = RANK3 ? L / ny : 0, iy = RANK3 ? L - ix * ny : L;
        unsigned f = ((unsigned)z * W) << 9;
        uint32_t xm = own[k], xp = own[k];
        if (valid) {
            f |= RF_VALID;
            if (z == 0) f |= RF_Z0;
            if (z + V >= nz) f |= RF_ZN;
            if (iy > 0) f |= RF_YM;
            if (iy + 1 < ny) f |= RF_YP;
            if (l == 0) f |= RF_LO;
            if (l == nl - 1) f |= RF_HI;
            if constexpr (RANK3) {
                // address of the same z in global line Lq (local or in a peer CTA)
                auto addr_of = [&](int Lq) -> uint32_t {
                    const int c = Lq / lpc;
                    const uint32_t local = d_a + ((uint32_t)(Lq - c * lpc) * (uint32_t)nz + (uint32_t)z) * W;
                    if constexpr (CLUSTER) return (c == rank) ? local : mapa_u32(local, (uint32_t)c);
                    else return local;
                };
                if (ix > 0) { f |= RF_XM; xm = addr_of(L - ny); }
                if (ix + 1 < nx) { f |= RF_XP; xp = addr_of(L + ny); }
            }
        }
        fl[k] = f;
        if constexpr (RANK3) { a_xm[k] = xm; a_xp[k] = xp; }
    }
    uint32_t red_peer[CLUSTER ? kResMaxCluster : 1];
    if constexpr (CLUSTER) {
#pragma unroll
        for (int c = 0; c < kResMaxCluster; ++c) red_peer[c] = mapa_u32(red_a, (uint32_t)(c < cs ? c : rank));
    }
    Pack<T> zero;
#pragma unroll
    for (int e = 0; e < V; ++e) zero.v[e] = T(0);
    // neighbour loads: own CTA through ld.shared, a peer's line through ld.shared::cluster
    auto yload = [&](int k, bool minus, uint32_t boff) -> Pack<T> {
        const unsigned f = fl[k];
        if constexpr (CLUSTER) {
            if (f & (minus ? RF_LO : RF_HI)) return ldc_pack((minus ? rem_lo : rem_hi) + (f >> 9) + boff, T(0));
        }
        return lds_pack((minus ? own[k] - line_bytes : own[k] + line_bytes) + boff, T(0));
    };
    auto xload = [&](uint32_t addr, uint32_t boff) -> Pack<T> {
        if constexpr (CLUSTER) {
            if (addr - d_a >= vec_bytes) return ldc_pack(addr + boff, T(0));   // mapa'd addresses lie outside the own window range
        }
        return lds_pack(addr + boff, T(0));
    };

    // y = A v for pack k, v in the buffer at byte offset `boff` from the direction buffer (own + peers), centre pack `ce`;
    // same summation order as k_star
    auto stencil_at = [&](int k, const Pack<T>& ce, uint32_t boff) -> Pack<T> {
        const unsigned f = fl[k];
        const T up = __shfl_up_sync(0xffffffffu, ce.v[V - 1], 1), dn = __shfl_down_sync(0xffffffffu, ce.v[0], 1);
        Pack<T> out = zero;
        if (f & RF_VALID) {
            const T left = (f & RF_Z0) ? T(0) : (lane == 0 ? lds_one(own[k] + boff - W, T(0)) : up);
            const T right = (f & RF_ZN) ? T(0) : (lane == 31 ? lds_one(own[k] + boff + V * W, T(0)) : dn);
            const Pack<T> ym = (f & RF_YM) ? yload(k, true, boff) : zero;
            const Pack<T> yp = (f & RF_YP) ? yload(k, false, boff) : zero;
            Pack<T> xm = zero, xp = zero;
            if constexpr (RANK3) {
                if (f & RF_XM) xm = xload(a_xm[k], boff);
                if (f & RF_XP) xp = xload(a_xp[k], boff);
            }
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const T zm = e == 0 ? left : ce.v[e > 0 ? e - 1 : 0];
                const T zp = e == V - 1 ? right : ce.v[e < V - 1 ? e + 1 : 0];
                T sum;
                if constexpr (RANK3) {
                    sum = mul_rn(cf.cxm, xm.v[e]);
                    sum = add_rn(sum, mul_rn(cf.cym, ym.v[e]));
                } else {
                    sum = mul_rn(cf.cym, ym.v[e]);
                }
                sum = add_rn(sum, mul_rn(cf.czm, zm));
                sum = add_rn(sum, mul_rn(cf.c0, ce.v[e]));
                sum = add_rn(sum, mul_rn(cf.czp, zp));
                sum = add_rn(sum, mul_rn(cf.cyp, yp.v[e]));
                if constexpr (RANK3) sum = add_rn(sum, mul_rn(cf.cxp, xp.v[e]));
                out.v[e] = sum;
            }
        }
        return out;
    };
    auto stencil = [&](int k, const Pack<T>& ce) -> Pack<T> { return stencil_at(k, ce, 0u); };
    // sum of the V products of a pack, added to the double accumulator
    auto dot_pack = [&](const Pack<T>& u, const Pack<T>& v, double& acc) {
        T p = mul_rn(u.v[0], v.v[0]);
#pragma unroll
        for (int e = 1; e < V; ++e) p = add_rn(p, mul_rn(u.v[e], v.v[e]));
        acc += (double)p;
    };

    uint32_t par = 0;
    auto reduce = [&](double v) -> double {
        v = butterfly_sum(v);
        if (lane == 0) asm volatile("st.shared.f64 [%0], %1;" ::"r"(red_a + (par * 32 + warp) * 8), "d"(v) : "memory");
        barrier();
        double tot = 0.0;
        if constexpr (CLUSTER) {
#pragma unroll
            for (int c = 0; c < kResMaxCluster; ++c)
                if (c < cs) tot += butterfly_sum(ldc_f64(red_peer[c] + (par * 32 + lane) * 8));
        } else {
            tot = butterfly_sum(lds_one(red_a + (par * 32 + lane) * 8, 0.0));
        }
        par ^= 1u;
        return tot;
    };

    // two sums with one barrier (adaptive variant); slots 2*32.. of the reduction scratch
    auto reduce2 = [&](double& v0, double& v1) {
        v0 = butterfly_sum(v0);
        v1 = butterfly_sum(v1);
        if (lane == 0) {
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(red_a + ((par * 2 + 0) * 32 + warp) * 8 + 512), "d"(v0) : "memory");
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(red_a + ((par * 2 + 1) * 32 + warp) * 8 + 512), "d"(v1) : "memory");
        }
        barrier();
        double t0 = 0.0, t1 = 0.0;
        if constexpr (CLUSTER) {
#pragma unroll
            for (int c = 0; c < kResMaxCluster; ++c)
                if (c < cs) {
                    t0 += butterfly_sum(ldc_f64(red_peer[c] + ((par * 2 + 0) * 32 + lane) * 8 + 512));
                    t1 += butterfly_sum(ldc_f64(red_peer[c] + ((par * 2 + 1) * 32 + lane) * 8 + 512));
                }
        } else {
            t0 = butterfly_sum(lds_one(red_a + ((par * 2 + 0) * 32 + lane) * 8 + 512, 0.0));
            t1 = butterfly_sum(lds_one(red_a + ((par * 2 + 1) * 32 + lane) * 8 + 512, 0.0));
        }
        par ^= 1u;
        v0 = t0;
        v1 = t1;
    };

    const int n_clusters = (int)gridDim.x / cs;
    const T tol = (T)a.glob->tol_min;
    for (int b = (int)blockIdx.x / cs; b < a.batch; b += n_clusters) {
        Sys& S = a.sys[b];
        int go = S.go, its = S.its, fev = S.fev, pass = S.pad0, conv = S.conv, div = S.div;
        const int max_iter = S.max_iter;
        auto stop_now = [&]() { return pass >= a.cap || (!go && (!a.torch_rules || pass >= a.lb)); };
        if (stop_now()) continue;   // uniform over the cluster
        T delta = (T)S.s[2 /*S_DELTA*/], beta = (T)S.s[1 /*S_BETA*/], dq = (T)S.s[3 /*S_DQ*/];
        const T rsq0 = (T)S.rsq0;
        double rsq = S.rsq;
        const int64_t base = (int64_t)b * a.ld + (int64_t)l0 * nz;
        T* __restrict__ Xg = (T*)a.X + base;
        T* __restrict__ Rg = (T*)a.R + base;
        T* __restrict__ Dg = (T*)a.D + base;
        const T* __restrict__ Yg = (const T*)a.Y + base;
        if constexpr (ALG == 1) {
            // ---------------- cg_adaptive: x += alpha d ; r -= alpha q ; d = r - (r.q / d.q) d ; q = A d ; alpha = d.r / d.q
            T* __restrict__ Qg = (T*)a.Q + base;
            T alpha = (T)S.s[0 /*S_ALPHA*/];
            T dr = (T)S.s[4 /*S_DR*/], rqv = (T)S.s[5 /*S_RQ*/];
            Pack<T> r[QPT];
#pragma unroll
            for (int k = 0; k < QPT; ++k) {
                r[k] = zero;
                if (fl[k] & RF_VALID) {
                    const int o = (int)((own[k] - d_a) / W);
                    sts_pack(own[k] + x_off, ldg_pack<T>(Xg + o));
                    r[k] = ldg_pack<T>(Rg + o);
                    sts_pack(own[k], ldg_pack<T>(Dg + o));
                    sts_pack(own[k] + q_off, ldg_pack<T>(Qg + o));
                }
            }
            while (!stop_now()) {
                pass += 1;
                const bool refresh = a.torch_rules && (pass % 20 == 0);
                double p0 = 0.0, p1 = 0.0;
                if (refresh) {   // x += alpha d first, then r = y - A x needs every CTA's x
#pragma unroll
                    for (int k = 0; k < QPT; ++k) {
                        if (fl[k] & RF_VALID) {
                            const Pack<T> dv = lds_pack(own[k], T(0));
                            Pack<T> xv = lds_pack(own[k] + x_off, T(0));
#pragma unroll
                            for (int e = 0; e < V; ++e) xv.v[e] = add_rn(xv.v[e], mul_rn(alpha, dv.v[e]));
                            sts_pack(own[k] + x_off, xv);
                        }
                    }
                    barrier();
                }
#pragma unroll
                for (int k = 0; k < QPT; ++k) {
                    Pack<T> tv = zero;
                    if (refresh) tv = stencil_at(k, (fl[k] & RF_VALID) ? lds_pack(own[k] + x_off, T(0)) : zero, x_off);
                    if (fl[k] & RF_VALID) {
                        const Pack<T> qv = lds_pack(own[k] + q_off, T(0));
                        if (refresh) {
                            const Pack<T> yv = ldg_pack<T>(Yg + (int)((own[k] - d_a) / W));
#pragma unroll
                            for (int e = 0; e < V; ++e) r[k].v[e] = add_rn(yv.v[e], -tv.v[e]);
                        } else {
                            const Pack<T> dv = lds_pack(own[k], T(0));
                            Pack<T> xv = lds_pack(own[k] + x_off, T(0));
#pragma unroll
                            for (int e = 0; e < V; ++e) {
                                xv.v[e] = add_rn(xv.v[e], mul_rn(alpha, dv.v[e]));
                                r[k].v[e] = add_rn(r[k].v[e], -mul_rn(alpha, qv.v[e]));
                            }
                            sts_pack(own[k] + x_off, xv);
                        }
                        dot_pack(r[k], r[k], p0);
                        dot_pack(r[k], qv, p1);
                    }
                }
                reduce2(p0, p1);
                if (refresh) fev += 1;
                its += go;
                delta = (T)p0;
                rqv = (T)p1;
                rsq = (double)delta;
                // ---- direction d = r - (rq d) / dq   (the barrier of the reduction ordered every read of d before this)
#pragma unroll
                for (int k = 0; k < QPT; ++k) {
                    if (fl[k] & RF_VALID) {
                        Pack<T> dv = lds_pack(own[k], T(0));
#pragma unroll
                        for (int e = 0; e < V; ++e) {
                            const T tt = mul_rn(rqv, dv.v[e]);
                            const T u = dq == T(0) ? T(0) : tt / dq;
                            dv.v[e] = add_rn(r[k].v[e], -u);
                        }
                        sts_pack(own[k], dv);
                    }
                }
                barrier();
                // ---- q = A d ; d.q, d.r
                p0 = 0.0;
                p1 = 0.0;
#pragma unroll
                for (int k = 0; k < QPT; ++k) {
                    const Pack<T> ce = (fl[k] & RF_VALID) ? lds_pack(own[k], T(0)) : zero;
                    const Pack<T> qv = stencil(k, ce);
                    if (fl[k] & RF_VALID) {
                        sts_pack(own[k] + q_off, qv);
                        dot_pack(ce, qv, p0);
                        dot_pack(ce, r[k], p1);
                    }
                }
                reduce2(p0, p1);
                fev += go;
                dq = (T)p0;
                dr = (T)p1;
                const T rq_ = (T)fabs(rsq);
                const bool cv = rq_ <= tol;
                const T ratio = rq_ / rsq0;
                const bool dvg = a.torch_rules ? ((ratio > T(100)) && its >= 8) : (((ratio > T(1e5)) && its >= 8) || !isfinite(rq_));
                conv = cv;
                div = dvg;
                go = (!cv && !dvg && its < max_iter) ? 1 : 0;
                alpha = go ? safe_div(dr, dq) : T(0);
            }
#pragma unroll
            for (int k = 0; k < QPT; ++k) {
                if (fl[k] & RF_VALID) {
                    const int o = (int)((own[k] - d_a) / W);
                    stg_pack(Xg + o, lds_pack(own[k] + x_off, T(0)));
                    stg_pack(Rg + o, r[k]);
                    stg_pack(Dg + o, lds_pack(own[k], T(0)));
                    stg_pack(Qg + o, lds_pack(own[k] + q_off, T(0)));
                }
            }
            if (rank == 0 && t == 0) {
                S.s[0 /*S_ALPHA*/] = (double)alpha;
                S.s[2 /*S_DELTA*/] = (double)delta;
                S.s[3 /*S_DQ*/] = (double)dq;
                S.s[4 /*S_DR*/] = (double)dr;
                S.s[5 /*S_RQ*/] = (double)rqv;
                S.rsq = rsq;
                S.its = its;
                S.fev = fev;
                S.go = go;
                S.conv = conv;
                S.div = div;
                S.pad0 = pass;
            }
            barrier();
        } else {
        Pack<T> r[QPT];
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
            r[k] = zero;
            if (fl[k] & RF_VALID) {
                sts_pack(own[k] + x_off, ldg_pack<T>(Xg + (int)((own[k] - d_a) / W)));
                r[k] = ldg_pack<T>(Rg + (int)((own[k] - d_a) / W));
                sts_pack(own[k], ldg_pack<T>(Dg + (int)((own[k] - d_a) / W)));
            }
        }
        while (!stop_now()) {
            pass += 1;
            const bool refresh = a.torch_rules && (pass % 20 == 0);
            // ---- direction: d = M^-1 r + beta d
#pragma unroll
            for (int k = 0; k < QPT; ++k) {
                if (fl[k] & RF_VALID) {
                    Pack<T> dv = lds_pack(own[k], T(0));
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const T s = PRE == 3 ? mul_rn(r[k].v[e], cf.cinv) : r[k].v[e];
                        dv.v[e] = add_rn(s, mul_rn(beta, dv.v[e]));
                    }
                    sts_pack(own[k], dv);
                }
            }
            barrier();
            // ---- q = A d, d.q
            double part = 0.0;
#pragma unroll
            for (int k = 0; k < QPT; ++k) {
                const Pack<T> ce = (fl[k] & RF_VALID) ? lds_pack(own[k], T(0)) : zero;
                const Pack<T> qv = stencil(k, ce);
                if (fl[k] & RF_VALID) {
                    sts_pack(own[k] + q_off, qv);
                    dot_pack(ce, qv, part);
                }
            }
            dq = (T)reduce(part);
            its += go;
            fev += go;
            const T alpha = go ? safe_div(delta, dq) : T(0);
            // ---- x += alpha d ; r -= alpha q (or r = y - A x on a refresh pass) ; r.M^-1 r
            part = 0.0;
            if (!refresh) {
#pragma unroll
                for (int k = 0; k < QPT; ++k) {
                    if (fl[k] & RF_VALID) {
                        const Pack<T> dv = lds_pack(own[k], T(0)), qv = lds_pack(own[k] + q_off, T(0));
                        Pack<T> xv = lds_pack(own[k] + x_off, T(0)), sv;
#pragma unroll
                        for (int e = 0; e < V; ++e) {
                            xv.v[e] = add_rn(xv.v[e], mul_rn(alpha, dv.v[e]));
                            r[k].v[e] = add_rn(r[k].v[e], -mul_rn(alpha, qv.v[e]));
                            sv.v[e] = PRE == 3 ? mul_rn(r[k].v[e], cf.cinv) : r[k].v[e];
                        }
                        sts_pack(own[k] + x_off, xv);
                        dot_pack(r[k], sv, part);
                    }
                }
            } else {
                // every CTA has finished reading the direction (the reduction above was a barrier): park d in q, put x in d
#pragma unroll
                for (int k = 0; k < QPT; ++k) {
                    if (fl[k] & RF_VALID) {
                        const Pack<T> dv = lds_pack(own[k], T(0));
                        Pack<T> xv = lds_pack(own[k] + x_off, T(0));
#pragma unroll
                        for (int e = 0; e < V; ++e) xv.v[e] = add_rn(xv.v[e], mul_rn(alpha, dv.v[e]));
                        sts_pack(own[k] + x_off, xv);
                        sts_pack(own[k] + q_off, dv);
                        sts_pack(own[k], xv);
                    }
                }
                barrier();
#pragma unroll
                for (int k = 0; k < QPT; ++k) {
                    const Pack<T> tv = stencil(k, (fl[k] & RF_VALID) ? lds_pack(own[k], T(0)) : zero);
                    if (fl[k] & RF_VALID) {
                        const Pack<T> yv = ldg_pack<T>(Yg + (int)((own[k] - d_a) / W));
#pragma unroll
                        for (int e = 0; e < V; ++e) r[k].v[e] = add_rn(yv.v[e], -tv.v[e]);
                        dot_pack(r[k], r[k], part);
                    }
                }
            }
            const T dnew = (T)reduce(part);
            if (refresh) {
                fev += 1;
#pragma unroll
                for (int k = 0; k < QPT; ++k)
                    if (fl[k] & RF_VALID) sts_pack(own[k], lds_pack(own[k] + q_off, T(0)));
            }
            beta = safe_div(dnew, delta);
            delta = dnew;
            rsq = (double)dnew;
            // ---- progress check of stop_on_l2 / torch_sparse_cg for this system
            const T rq = (T)fabs(rsq);
            const bool cv = rq <= tol;
            const T ratio = rq / rsq0;
            const bool dv = a.torch_rules ? ((ratio > T(100)) && its >= 8) : (((ratio > T(1e5)) && its >= 8) || !isfinite(rq));
            conv = cv;
            div = dv;
            go = (!cv && !dv && its < max_iter) ? 1 : 0;
        }
        // ---- write the state back (d too: a later launch — next round of the torch rules — continues from it)
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
            if (fl[k] & RF_VALID) {
                stg_pack(Xg + (int)((own[k] - d_a) / W), lds_pack(own[k] + x_off, T(0)));
                stg_pack(Rg + (int)((own[k] - d_a) / W), r[k]);
                stg_pack(Dg + (int)((own[k] - d_a) / W), lds_pack(own[k], T(0)));
            }
        }
        if (rank == 0 && t == 0) {
            S.s[0 /*S_ALPHA*/] = 0.0;
            S.s[1 /*S_BETA*/] = (double)beta;
            S.s[2 /*S_DELTA*/] = (double)delta;
            S.s[3 /*S_DQ*/] = (double)dq;
            S.rsq = rsq;
            S.its = its;
            S.fev = fev;
            S.go = go;
            S.conv = conv;
            S.div = div;
            S.pad0 = pass;
        }
        barrier();   // nobody overwrites its direction buffer for the next system while a peer still reads it
        }
    }
}

// max pass over the batch and whether every system sits at that pass without wanting to continue
__global__ void __launch_bounds__(kBlock) k_res_round(const Sys* __restrict__ sys, int batch, int cap, Glob* glob, int* out /* [2]: max pass, all settled */) {
    __shared__ int s_max, s_unsettled, s_any;
    if (threadIdx.x == 0) { s_max = 0; s_unsettled = 0; s_any = 0; }
    __syncthreads();
    int m = 0, any = 0;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        m = max(m, sys[b].pad0);
        any |= sys[b].go;
    }
    atomicMax(&s_max, m);
    if (any) atomicOr(&s_any, 1);
    __syncthreads();
    const int P = s_max;
    int uns = 0;
    for (int b = threadIdx.x; b < batch; b += blockDim.x)
        if (sys[b].pad0 != P || (sys[b].go && P < cap)) uns = 1;
    if (uns) atomicOr(&s_unsettled, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        out[0] = P;
        out[1] = s_unsettled ? 0 : 1;
        glob->any_go = s_any;
        glob->pass = P;
        glob->checks = glob->checks + 1;
    }
}

}  // namespace phb
