// Fourth formulation of the structured triangular solves (ILU factors of a recognised star stencil): X-MARCHING ROW PIPELINES.
//
// The earlier forms tile the x-y plane (8 planes x 8 rows per CTA) and pay one hand-over through L2 per tile edge in BOTH x and y
// (profiles/r01b_ncu_resident_trsv.txt: 70 tile-chunk steps of ~9 us at 256^3 = the whole run time).  Here the x direction never
// leaves a thread:
//   * a warp owns ONE row chunk — row y, 32 x V consecutive z' (V = 4 floats / 2 doubles per lane, one 16-byte vector) — and marches over
//     ALL x planes; the x neighbours of a step are the warp's own previous results (registers).  The z recurrence w_k = a_k + b_k w_(k-1)
//     is composed over the lane's V elements first, then across lanes with the 5-step shuffle scan, then expanded locally again;
//   * the R warps of a CTA own R consecutive rows and form a systolic pipeline: warp w is one plane behind warp w-1 and receives its
//     y neighbours through a ring of self-validating slots in shared memory ({value, plane tag} in one 64-bit word: the reader spins on
//     the slot itself, no separate flag, no fence);
//   * only two kinds of edges cross CTAs — the row above the CTA's first row (tile j-1, same chunk) and the carry of the z recurrence
//     (same rows, chunk c-1).  They travel through two small global hand-over arrays of the SAME self-validating words (tag = launch
//     epoch + plane): the producer just stores as it goes, the consumer loads the words of the next G planes one group ahead and
//     re-reads only those whose tag has not arrived — no flags, no release / acquire fences anywhere.
// What the measurements said on the way (256^3 fp32, L + U, row-scan kernel 1.20 ms): one element per lane with per-warp progress
// counters (fence.acq_rel + st.release per group) 1.29 ms; the same with self-validating words 0.93 ms — and identical run time with
// the scans, the stores or the hand-over compiled out: ~250 instructions per 32-element step made the kernel ISSUE bound (3.5 warps per
// scheduler, all busy).  Hence V elements per lane: a quarter of the warps, a third of the instructions per element.
// Critical path: (nx + ny + nz/(32 V)) steps of one warp scan plus (ny/R + nz/(32 V)) L2 round trips, instead of
// (nx/8 + ny/8 + nz/32) x 15 in-tile levels with CTA barriers and an L2 hand-over each.
#pragma once
#include "star_trsv.cuh"

namespace phb {

constexpr int kPpRing = 4;         // in-CTA y hand-over ring (planes) = planes per group: slot of plane g0 + p is p
constexpr int kPpDepth = 16;       // planes of rhs / pivots staged ahead per warp (cp.async)
constexpr int kPpGroup = 4;        // planes per group (boundary words are requested one group ahead)

struct PipeTrsvArgs {
    int nx, ny, nz;
    int nj, nch;            // row tiles (R rows each), chunks of 32 V elements along z
    int batch;
    int64_t ld;
    const void* rd;
    const void* B;
    void* X;
    unsigned long long* ybuf;       // [batch][nj][nx][nch*32*V][W] last row of tile j, self-validating words {value word, tag}
    unsigned long long* zbuf;       // [batch][nch][nx][ny][W]      last z' of chunk c
    unsigned tag_base;              // tags of this launch: tag_base + plane + 1
    unsigned long long* dbg;        // optional: per-CTA per-warp cycle counters [cta][R][4] = {total, boundary wait, 0, 0}
    unsigned* ticket;
    unsigned n_ctas;
    const int32_t* order;   // [nj*nch] CTA ids (j * nch + c) sorted by j + c: dependencies first
    double cx, cy, cz;
};

template <typename T> constexpr int pipe_vec() { return 16 / (int)sizeof(T); }
template <typename T, int R> constexpr size_t pipe_ring_bytes() { return (size_t)R * kPpRing * 32 * pipe_vec<T>() * ((int)sizeof(T) / 4) * sizeof(unsigned long long); }
template <typename T, int R> constexpr size_t pipe_smem_bytes() { return pipe_ring_bytes<T, R>() + (size_t)R * kPpDepth * 2 * 32 * 16; }

// shared-memory accessors on 32-bit shared-window addresses (computed once; a generic pointer costs a window-base computation per access)
__device__ __forceinline__ unsigned long long lds_volatile_u64(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_u64(uint32_t addr, unsigned long long v) {
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void stg_relaxed_u64(unsigned long long* p, unsigned long long v) {       // 8-byte stores are single-copy atomic
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <typename T> __device__ __forceinline__ Vec16<T> lds16(uint32_t addr);
template <> __device__ __forceinline__ Vec16<float> lds16<float>(uint32_t addr) {
    Vec16<float> r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(addr) : "memory");
    return r;
}
template <> __device__ __forceinline__ Vec16<double> lds16<double>(uint32_t addr) {
    Vec16<double> r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void stg16(float* p, const Vec16<float>& v) { *reinterpret_cast<float4*>(p) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]); }
__device__ __forceinline__ void stg16(double* p, const Vec16<double>& v) { *reinterpret_cast<double2*>(p) = make_double2(v.v[0], v.v[1]); }

__device__ __forceinline__ unsigned long long mk64(unsigned lo, unsigned hi) {
    unsigned long long v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
    return v;
}
__device__ __forceinline__ unsigned lo32(unsigned long long v) {
    unsigned lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
    return lo;
}
__device__ __forceinline__ unsigned hi32(unsigned long long v) {
    unsigned lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
    return hi;
}

// value <-> tagged 64-bit words (W = 1 word for float, 2 for double: low / high half, each with the tag)
template <typename T> struct Tagged;
template <> struct Tagged<float> {
    static constexpr int W = 1;
    __device__ static void pack(float v, unsigned tag, unsigned long long (&w)[1]) { w[0] = mk64(__float_as_uint(v), tag); }
    __device__ static float unpack(const unsigned long long (&w)[1]) { return __uint_as_float(lo32(w[0])); }
};
template <> struct Tagged<double> {
    static constexpr int W = 2;
    __device__ static void pack(double v, unsigned tag, unsigned long long (&w)[2]) {
        w[0] = mk64((unsigned)__double2loint(v), tag);
        w[1] = mk64((unsigned)__double2hiint(v), tag);
    }
    __device__ static double unpack(const unsigned long long (&w)[2]) { return __hiloint2double((int)lo32(w[1]), (int)lo32(w[0])); }
};

template <typename T, bool LOWER, int R>
__global__ void __launch_bounds__(R * 32) k_star_trsv_pipe(const PipeTrsvArgs a) {
    constexpr int RING = kPpRing, D = kPpDepth, G = kPpGroup, NG = D / G, V = pipe_vec<T>(), W = Tagged<T>::W, ZC = 32 * V;
    static_assert(D % G == 0 && RING == G && (D & (D - 1)) == 0, "ring sizes");
    extern __shared__ __align__(16) unsigned char pipe_smem[];
    unsigned long long* ring = reinterpret_cast<unsigned long long*>(pipe_smem);      // [R][RING][V][W][32]
    __shared__ unsigned s_ticket;
    __shared__ volatile int s_prog[R];             // planes whose y neighbours warp w has consumed
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u);
    if (tid < R) s_prog[tid] = 0;
    for (int k = tid; k < R * RING * V * W * 32; k += R * 32) ring[k] = 0ull;         // tag 0 = empty (tags start at 1)
    __syncthreads();
    const unsigned cta = s_ticket;
    if (tid == 0 && cta == a.n_ctas - 1) *a.ticket = 0u;
    const int n_tiles = a.nj * a.nch;
    const int b = (int)(cta / (unsigned)n_tiles);
    const int id = a.order[cta % (unsigned)n_tiles];
    const int tj = id / a.nch, c = id % a.nch;
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const int y0 = tj * R, gy = y0 + warp;
    if (gy >= ny) return;                          // rows beyond the domain: nobody depends on them
    const bool has_down = warp < R - 1 && gy + 1 < ny;
    const bool has_up = warp > 0;
    const T cx = (T)a.cx, cy = (T)a.cy, cz = (T)a.cz;
    const T* __restrict__ rd = (const T*)a.rd;
    const T* Bb = (const T*)a.B + (int64_t)b * a.ld;
    T* Xb = (T*)a.X + (int64_t)b * a.ld;
    // logical coordinates run in solve order; the upper solve mirrors all three axes.  Lane l holds logical z' = zl0 .. zl0 + V - 1
    // (nz % V == 0, so a vector is either fully inside or fully outside); physically that is ONE aligned vector, reversed for the upper solve.
    const int zl0 = c * ZC + lane * V;
    const bool z_ok = zl0 < nz;
    const int zph0 = LOWER ? zl0 : nz - V - zl0;   // physical index of the vector's first element
    const int pyr = LOWER ? gy : ny - 1 - gy;
    const int64_t ps = (int64_t)ny * nz;
    const int64_t xstep = LOWER ? ps : -ps;        // physical offset from one logical plane to the next
    const int64_t off0 = (LOWER ? (int64_t)0 : (int64_t)(nx - 1) * ps) + (int64_t)pyr * nz + (z_ok ? zph0 : 0);   // element offset of logical plane 0
    const bool feeds_y = (warp == R - 1) && (y0 + R < ny);                 // the tile below reads this row
    const bool feeds_z = c + 1 < a.nch;                                    // chunk c+1 reads this row's last z'
    const bool needs_y = tj > 0 && warp == 0, needs_z = c > 0;
    const int64_t zrow = (int64_t)a.nch * ZC;
    unsigned long long* y_out = a.ybuf + ((((int64_t)b * a.nj + tj) * nx) * zrow + zl0) * W;
    const unsigned long long* y_in = a.ybuf + ((((int64_t)b * a.nj + (tj > 0 ? tj - 1 : 0)) * nx) * zrow + zl0) * W;
    unsigned long long* z_out = a.zbuf + ((((int64_t)b * a.nch + c) * nx) * ny + gy) * W;
    const unsigned long long* z_in = a.zbuf + ((((int64_t)b * a.nch + (c > 0 ? c - 1 : 0)) * nx) * ny + gy) * W;
    const int64_t y_stride = zrow * W, z_stride = (int64_t)ny * W;
    const int n_valid = min(ZC, nz - c * ZC);                              // valid z' of this chunk
    const int last_lane = (n_valid - 1) / V, e_last = (n_valid - 1) % V;   // where the chunk's last valid z' lives
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t my_ring = ring_u32 + (uint32_t)((warp * RING * V * W * 32 + lane) * 8);
    const uint32_t up_ring = ring_u32 + (uint32_t)(((has_up ? warp - 1 : 0) * RING * V * W * 32 + lane) * 8);
    const unsigned tag_base = a.tag_base;

    // ---- rhs / pivot vectors of this row chunk, staged D planes ahead (each lane copies and later reads its own 16 bytes)
    const uint32_t in_ring_u32 = (uint32_t)__cvta_generic_to_shared(pipe_smem) + (uint32_t)(pipe_ring_bytes<T, R>() + ((size_t)warp * D * 2 * 32 + lane) * 16);
    int x_issue = 0;
    int64_t off_issue = off0;
    auto issue_group = [&]() {
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const bool ok = x_issue < nx && z_ok;
            const uint32_t dst = in_ring_u32 + (uint32_t)((x_issue & (D - 1)) * 2 * 32 * 16);
            cp_async16(dst, Bb + (ok ? off_issue : 0), ok);
            cp_async16(dst + 32 * 16, rd + (ok ? off_issue : 0), ok);
            x_issue += 1;
            off_issue += xstep;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < NG; ++k) issue_group();

    // ---- boundary words of G planes: y words (V per plane, this lane's own z') and carry words (lane p <-> plane g + p)
    unsigned long long yw[G][V][W], zw[W];
    auto load_boundary = [&](int g) {
        if (needs_y) {
#pragma unroll
            for (int p = 0; p < G; ++p)
#pragma unroll
                for (int e = 0; e < V; ++e)
#pragma unroll
                    for (int k = 0; k < W; ++k) yw[p][e][k] = (g + p < nx) ? ld_poll_u64(y_in + ((int64_t)g * y_stride + (int64_t)p * y_stride) + e * W + k) : 0ull;
        }
        if (needs_z) {
#pragma unroll
            for (int k = 0; k < W; ++k) zw[k] = (lane < G && g + lane < nx) ? ld_poll_u64(z_in + (int64_t)(g + lane) * z_stride + k) : 0ull;
        }
    };
    auto validate_boundary = [&](int g, int gn) {                                     // re-read what has not arrived yet
        if (!needs_y && !needs_z) return;
        unsigned spins = 0;
        while (true) {
            bool ok = true;
            if (needs_y) {
#pragma unroll
                for (int p = 0; p < G; ++p)
#pragma unroll
                    for (int e = 0; e < V; ++e)
#pragma unroll
                        for (int k = 0; k < W; ++k)
                            if (p < gn && hi32(yw[p][e][k]) != tag_base + (unsigned)(g + p + 1)) {
                                yw[p][e][k] = ld_poll_u64(y_in + ((int64_t)g * y_stride + (int64_t)p * y_stride) + e * W + k);
                                ok = ok && hi32(yw[p][e][k]) == tag_base + (unsigned)(g + p + 1);
                            }
            }
            if (needs_z && lane < gn) {
#pragma unroll
                for (int k = 0; k < W; ++k)
                    if (hi32(zw[k]) != tag_base + (unsigned)(g + lane + 1)) {
                        zw[k] = ld_poll_u64(z_in + (int64_t)(g + lane) * z_stride + k);
                        ok = ok && hi32(zw[k]) == tag_base + (unsigned)(g + lane + 1);
                    }
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (++spins > (1u << 26)) __trap();
        }
    };

    T wx[V];
#pragma unroll
    for (int e = 0; e < V; ++e) wx[e] = T(0);
    long long t_bound = 0;
    const long long t_begin = a.dbg ? clock64() : 0;
    T* xp = Xb + off0;                             // current plane of this row chunk in X
    unsigned long long* yo = y_out;                // current plane in the hand-over arrays
    unsigned long long* zo = z_out;
    load_boundary(0);
    for (int g0 = 0; g0 < nx; g0 += G) {
        const int gn = min(G, nx - g0);
        // ---- this group's inputs have landed (NG-1 younger groups may still be in flight); their slots take the group D planes ahead
        cp_async_wait<NG - 1>();
        const uint32_t in_base = in_ring_u32 + (uint32_t)((g0 & (D - 1)) * 2 * 32 * 16);
        // ---- the B halves of the group's recurrences depend on the pivots only: lane-local products, then the G warp scans side by side.
        //      (vectors beyond nz were zero-filled by the copies: their pivots are 0, so a = b = 0 and they drop out without any select)
        T Bs[G][5], Bf[G];
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const Vec16<T> vr = lds16<T>(in_base + (uint32_t)((p * 2 + 1) * 32 * 16));
            Bf[p] = -cz * vr.v[0];
#pragma unroll
            for (int e = 1; e < V; ++e) Bf[p] = Bf[p] * (-cz * vr.v[e]);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
#pragma unroll
            for (int p = 0; p < G; ++p) {
                Bs[p][k] = Bf[p];
                const T Bp = __shfl_up_sync(0xffffffffu, Bf[p], 1 << k);
                if (lane >= (1 << k)) Bf[p] = Bf[p] * Bp;
            }
        }
        // ---- cross-CTA neighbours of this group: their words were requested one group ago; re-read the ones that had not arrived
        if (a.dbg) { const long long t1 = clock64(); validate_boundary(g0, gn); t_bound += clock64() - t1; }
        else validate_boundary(g0, gn);
        const T czv = needs_z ? Tagged<T>::unpack(zw) : T(0);
        T by[G][V];                  // (only warp 0 of a tile with a tile above carries values here)
#pragma unroll
        for (int p = 0; p < G; ++p)
#pragma unroll
            for (int e = 0; e < V; ++e) by[p][e] = needs_y ? Tagged<T>::unpack(yw[p][e]) : T(0);
        load_boundary(g0 + G);       // words of the next group travel while this one is computed
#pragma unroll
        for (int p = 0; p < G; ++p) {
            if (p < gn) {
                const int x = g0 + p;
                const unsigned tag = (unsigned)(x + 1);
                constexpr int kSlotWords = V * W * 32;
                const int slot = p * kSlotWords;               // RING == G and g0 % G == 0
                if (has_down && x >= RING) {                      // the slot of this plane is free once warp+1 has consumed plane x - RING
                    unsigned spins = 0;
                    while (s_prog[warp + 1] < x - RING + 1)
                        if (++spins > (1u << 26)) __trap();
                }
                const T carry = __shfl_sync(0xffffffffu, czv, p);
                const Vec16<T> vb = lds16<T>(in_base + (uint32_t)(p * 2 * 32 * 16)), vr = lds16<T>(in_base + (uint32_t)((p * 2 + 1) * 32 * 16));
                T g[V], cr[V], bl[V];
#pragma unroll
                for (int e = 0; e < V; ++e) {      // logical order: the upper solve walks the vector backwards
                    cr[e] = vr.v[LOWER ? e : V - 1 - e];
                    bl[e] = -cz * cr[e];
                    g[e] = fma_t<T>(-cx, wx[e], vb.v[LOWER ? e : V - 1 - e]);
                }
                if (has_up) {
                    // y neighbours from warp-1: spin on the slots themselves until they carry this plane's tag (all loads first, then the checks)
                    unsigned long long w[V][W];
#pragma unroll
                    for (int e = 0; e < V; ++e)
#pragma unroll
                        for (int k = 0; k < W; ++k) w[e][k] = lds_volatile_u64(up_ring + (uint32_t)((slot + (e * W + k) * 32) * 8));
                    unsigned spins = 0;
#pragma unroll
                    for (int e = 0; e < V; ++e) {
#pragma unroll
                        for (int k = 0; k < W; ++k) {
                            while (hi32(w[e][k]) != tag) {
                                w[e][k] = lds_volatile_u64(up_ring + (uint32_t)((slot + (e * W + k) * 32) * 8));
                                if (++spins > (1u << 26)) __trap();
                            }
                        }
                        g[e] = fma_t<T>(-cy, Tagged<T>::unpack(w[e]), g[e]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < V; ++e) g[e] = fma_t<T>(-cy, by[p][e], g[e]);
                }
                // lane-local composition: w_last = A + Bloc * w_in
                T al[V];
#pragma unroll
                for (int e = 0; e < V; ++e) al[e] = g[e] * cr[e];
                T A = al[0];
#pragma unroll
                for (int e = 1; e < V; ++e) A = fma_t<T>(bl[e], A, al[e]);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const T Ap = __shfl_up_sync(0xffffffffu, A, 1 << k);
                    if (lane >= (1 << k)) A = fma_t<T>(Bs[p][k], Ap, A);
                }
                const T w_last = fma_t<T>(Bf[p], carry, A);       // w at this lane's last element
                T w_in = __shfl_up_sync(0xffffffffu, w_last, 1);  // ... and at the element before this lane's first
                if (lane == 0) w_in = carry;
                T w[V];
                w[0] = fma_t<T>(bl[0], w_in, al[0]);
#pragma unroll
                for (int e = 1; e < V; ++e) w[e] = fma_t<T>(bl[e], w[e - 1], al[e]);
                if (has_down) {                                   // first thing once w is known: the row below is waiting for it
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        unsigned long long t[W];
                        Tagged<T>::pack(w[e], tag, t);
#pragma unroll
                        for (int k = 0; k < W; ++k) sts_volatile_u64(my_ring + (uint32_t)((slot + (e * W + k) * 32) * 8), t[k]);
                    }
                }
                if (feeds_y || (feeds_z && lane == last_lane)) {
                    const unsigned gtag = tag_base + tag;
                    if (feeds_y) {
#pragma unroll
                        for (int e = 0; e < V; ++e) {
                            unsigned long long t[W];
                            Tagged<T>::pack(w[e], gtag, t);
#pragma unroll
                            for (int k = 0; k < W; ++k) stg_relaxed_u64(yo + e * W + k, t[k]);
                        }
                    }
                    if (feeds_z && lane == last_lane) {
                        T wl = w[0];
#pragma unroll
                        for (int e = 1; e < V; ++e) wl = e == e_last ? w[e] : wl;
                        unsigned long long t[W];
                        Tagged<T>::pack(wl, gtag, t);
#pragma unroll
                        for (int k = 0; k < W; ++k) stg_relaxed_u64(zo + k, t[k]);
                    }
                }
                if (has_up && lane == 0) s_prog[warp] = x + 1;    // every lane has read its slots (the shuffles above synchronised the warp)
                Vec16<T> out;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    // the lower solve returns t = g - c_z w_(z-1), not w = t rd
                    const T v = LOWER ? fma_t<T>(-cz, e == 0 ? w_in : w[e - 1], g[e]) : w[e];
                    out.v[LOWER ? e : V - 1 - e] = v;
                }
                if (z_ok) stg16(xp, out);
                xp += xstep;
                yo += y_stride;
                zo += z_stride;
#pragma unroll
                for (int e = 0; e < V; ++e) wx[e] = w[e];
            }
        }
        issue_group();               // the slots just read take the group D planes ahead
    }
    cp_async_wait<0>();
    if (a.dbg && lane == 0) {
        unsigned long long* d = a.dbg + ((size_t)cta * R + warp) * 4;
        d[0] = (unsigned long long)(clock64() - t_begin); d[1] = (unsigned long long)t_bound; d[2] = 0; d[3] = 0;
    }
}

}  // namespace phb
