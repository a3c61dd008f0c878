// Krylov loops on the device: CG (generic / adaptive / torch fast-path rules) and BiCGstab(L).
// Algorithms follow phiml/backend/_linalg.py:23-274 and phiml/backend/torch/_torch_backend.py:1353-1413 (see
// include/phb200.h); the loop of Backend.while_loop (phiml/backend/_backend.py:697-735) runs on the device: every
// kernel starts with a test of Glob::any_go, per-system `go` flags freeze finished systems, the host only polls a
// pinned flag every few passes.
#include "dispatch.cuh"
#include "star_tma.cuh"
#include "precond.cuh"
#include "resident.cuh"
#include <vector>
#include <mutex>
#include <stdlib.h>
#include <stddef.h>

namespace phb {

enum Slot {
    S_ALPHA = 0, S_BETA, S_DELTA, S_DQ, S_DR, S_RQ, S_RHO0, S_RHO1, S_OMEGA, S_RHOP, S_GAMMA, S_YY,
    S_TAU = 12, S_SIGMA = 17, S_GAM = 22, S_GAMP = 27, S_GAMPP = 32   // 5 entries each: poly_order <= 4
};
static_assert(S_YY == kSlotYY, "progress_check reads |y|^2 from this slot");

enum Rules { RULES_GENERIC = 0, RULES_TORCH = 1 };

// ------------------------------------------------------------------------------------------------ functor base
struct FBase : Ctx {
    int always;   // 1: initialisation kernel — never idle, never skip
    int freeze;   // 1: finished systems are skipped (their x, r stay untouched)
    __device__ bool idle() const { return !always && glob->any_go == 0; }
    __device__ bool skip(int b) const { return !always && freeze && sys[b].go == 0; }
};

template <typename T>
__device__ __forceinline__ void check_all(const FBase& f, int rules) {
    progress_check<T>(f.sys, f.glob, f.batch, rules == RULES_TORCH, f.freeze != 0);
}

// tolerances of the batch-sum rule: tol_b = max(rtol_b^2 * sum_over_batch |y|^2, atol_b^2); all threads of one CTA
template <typename T>
__device__ __forceinline__ void apply_batch_sum(Sys* sys, int batch, const T* rtol, const T* atol, double yy_sum) {
    const T norm = (T)yy_sum;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        const T rt = rtol[b], at = atol[b];
        const T t1 = rt * rt * norm, t2 = at * at;
        ((volatile Sys*)sys)[b].tol = (double)(t1 > t2 ? t1 : (t2 >= t1 ? t2 : (T)NAN));
    }
}

// ------------------------------------------------------------------------------------------------ initial residual
// r = y - q (q = A x0), s = M r, d = s.  Dots: r.s, y.y, r.r.  Sets up tolerances (minimum over the batch) and runs
// the first progress check.  PRE: 0 none, 1 Jacobi (inv_diag), 2 external (r and s already computed).
template <typename T, int PRE>
struct InitResidual : FBase {
    static constexpr int NA = 5;   // r, y, q|s, d, w
    static constexpr unsigned RMASK = PRE == 2 ? 0b00111u : (PRE == 1 ? 0b10110u : 0b00110u);
    static constexpr unsigned WMASK = PRE == 2 ? 0b01000u : 0b01001u;
    static constexpr unsigned SHARED = 0b10000u;
    static constexpr int ND = 3;
    static constexpr bool GLOBAL = true;
    const void* v[NA];
    const T* rtol;
    const T* atol;
    const int32_t* max_iter;
    int tol_from_y;    // 0: tolerance relative to |r0.M^-1 r0| and monitor r.M^-1 r (generic cg); 1: |y|^2 and r.r
    int rules;
    int bicg;
    // generic cg_adaptive / bicg / bicg_stab_first_order hand stop_on_l2 the (batch,) vector sum(y**2, -1), which stop_on_l2 sums
    // AGAIN over its last axis (phiml/backend/_linalg.py:26): the tolerance of every system is relative to the sum of |y_b|^2
    // over the WHOLE batch.  (cg passes a (batch,1) array and the torch fast paths do their own per-system test.)
    __device__ bool batch_sum_rule() const { return tol_from_y && rules == RULES_GENERIC; }
    struct Scal {};
    __device__ Scal load(int) const { return Scal{}; }
    __device__ void elem(T (&e)[NA], double (&acc)[ND], const Scal&) const {
        T s;
        if (PRE == 2) {
            s = e[2];
        } else {
            e[0] = add_rn(e[1], -e[2]);
            s = PRE == 1 ? mul_rn(e[0], e[4]) : e[0];
        }
        e[3] = s;
        acc[0] += (double)mul_rn(e[0], s);
        acc[1] += (double)mul_rn(e[1], e[1]);
        acc[2] += (double)mul_rn(e[0], e[0]);
    }
    __device__ void finalize(int b, const double (&tot)[ND]) const {
        Sys& s = sys[b];
        const T delta0 = (T)tot[0], yy = (T)tot[1], rr = (T)tot[2];
        const T norm = tol_from_y ? yy : (T)fabs((double)delta0);
        const T rt = rtol[b], at = atol[b];
        const T t1 = rt * rt * norm, t2 = at * at;
        s.tol = (double)(t1 > t2 ? t1 : (t2 >= t1 ? t2 : (T)NAN));   // np.maximum propagates NaN
        s.rsq = (double)(tol_from_y ? rr : delta0);
        s.rsq0 = 0.0;
        s.its = 0;
        s.fev = 1;
        s.go = 1;
        s.conv = s.div = 0;
        s.max_iter = max_iter[b];
        s.pad0 = 0;   // passes made by this system (resident kernels)
        for (int k = 0; k < 40; ++k) s.s[k] = 0.0;
        s.s[S_DELTA] = (double)(tol_from_y ? rr : delta0);
        s.s[S_YY] = (double)yy;
        if (bicg) {
            s.s[S_RHO0] = s.s[S_RHO1] = s.s[S_OMEGA] = s.s[S_RHOP] = 1.0;
        }
    }
    __device__ void finalize_all() const {
        __shared__ double s_min[kBlock];
        {   // sum of |y_b|^2 over the batch (kept for batch sharding, where the sum runs over all ranks' systems)
            double part = 0.0;
            for (int b = threadIdx.x; b < batch; b += blockDim.x) part += ((volatile Sys*)sys)[b].s[S_YY];
            s_min[threadIdx.x] = part;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                for (int k = 0; k < (int)blockDim.x; ++k) tot += s_min[k];
                glob->yy_sum = tot;
            }
            __syncthreads();
            if (batch_sum_rule()) apply_batch_sum(sys, batch, rtol, atol, ((volatile Glob*)glob)->yy_sum);
            __syncthreads();
        }
        double m = INFINITY;
        bool nan = false;
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            double t = ((volatile Sys*)sys)[b].tol;
            if (t != t) nan = true;
            else m = fmin(m, t);
        }
        s_min[threadIdx.x] = nan ? NAN : m;
        __syncthreads();
        if (threadIdx.x == 0) {
            double mm = INFINITY;
            bool anynan = false;
            for (int k = 0; k < (int)blockDim.x; ++k) {
                if (s_min[k] != s_min[k]) anynan = true;
                else mm = fmin(mm, s_min[k]);
            }
            glob->tol_min = anynan ? NAN : mm;
            glob->checks = 0;
            glob->pass = 0;
            glob->refresh = 0;
            __threadfence();
        }
        __syncthreads();
        check_all<T>(*this, rules);
    }
};

// r = y - q only (ILU path: s = M r is then computed by the triangular solves before InitResidual<T,2>)
template <typename T>
struct Residual : FBase {
    static constexpr int NA = 3;   // r, y, q
    static constexpr unsigned RMASK = 0b110u, WMASK = 0b001u, SHARED = 0u;
    static constexpr int ND = 0;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    struct Scal {};
    __device__ Scal load(int) const { return Scal{}; }
    __device__ void elem(T (&e)[NA], double (&)[1], const Scal&) const { e[0] = add_rn(e[1], -e[2]); }
    __device__ void finalize(int, const double (&)[1]) const {}
    __device__ void finalize_all() const {}
};

// ------------------------------------------------------------------------------------------------ CG: q = A d, d.q
template <typename T>
struct CgDq : FBase {   // SpMV epilogue
    static constexpr int ND = 1;
    static constexpr bool GLOBAL = false;
    const T* d;
    int64_t ld;
    __device__ void row(int b, int64_t i, T y, double (&acc)[1]) const { acc[0] += (double)mul_rn(d[(int64_t)b * ld + i], y); }
    __device__ void finalize(int b, const double (&tot)[1]) const {
        Sys& s = sys[b];
        if (freeze && !s.go) return;
        s.its += s.go;
        s.fev += s.go;
        const T dq = (T)tot[0];
        s.s[S_DQ] = (double)dq;
        s.s[S_ALPHA] = s.go ? (double)safe_div((T)s.s[S_DELTA], dq) : 0.0;
    }
    __device__ void finalize_all() const {}
};

// x += alpha d ; r -= alpha q ; s = M r ; delta' = r.s     (PRE 0/1 fused; PRE 2: no dot, s comes from the ILU solves)
template <typename T, int PRE>
struct CgUpdate : FBase {
    static constexpr int NA = 5;   // x, r, d, q, w
    static constexpr unsigned RMASK = PRE == 1 ? 0b11111u : 0b01111u, WMASK = 0b00011u, SHARED = 0b10000u;
    static constexpr int ND = PRE == 2 ? 0 : 1;
    static constexpr bool GLOBAL = PRE != 2;
    const void* v[NA];
    int rules;
    struct Scal { T alpha; };
    __device__ Scal load(int b) const { return Scal{(T)sys[b].s[S_ALPHA]}; }
    __device__ void elem(T (&e)[NA], double (&acc)[ND == 0 ? 1 : ND], const Scal& sc) const {
        e[0] = add_rn(e[0], mul_rn(sc.alpha, e[2]));
        e[1] = add_rn(e[1], -mul_rn(sc.alpha, e[3]));
        if (PRE != 2) {
            const T s = PRE == 1 ? mul_rn(e[1], e[4]) : e[1];
            acc[0] += (double)mul_rn(e[1], s);
        }
    }
    __device__ void finalize(int b, const double (&tot)[ND == 0 ? 1 : ND]) const {
        Sys& s = sys[b];
        if (freeze && !s.go) return;
        const T dold = (T)s.s[S_DELTA], dnew = (T)tot[0];
        s.s[S_DELTA] = (double)dnew;
        s.s[S_BETA] = (double)safe_div(dnew, dold);
        s.rsq = (double)dnew;
    }
    __device__ void finalize_all() const {
        check_all<T>(*this, rules);
        if (threadIdx.x == 0) glob->pass = glob->pass + 1;
    }
};

// delta' = r.s with s given (ILU), same scalar update as CgUpdate
template <typename T>
struct CgDeltaDot : FBase {
    static constexpr int NA = 2;   // r, s
    static constexpr unsigned RMASK = 0b11u, WMASK = 0u, SHARED = 0u;
    static constexpr int ND = 1;
    static constexpr bool GLOBAL = true;
    const void* v[NA];
    int rules;
    struct Scal {};
    __device__ Scal load(int) const { return Scal{}; }
    __device__ void elem(T (&e)[NA], double (&acc)[1], const Scal&) const { acc[0] += (double)mul_rn(e[0], e[1]); }
    __device__ void finalize(int b, const double (&tot)[1]) const {
        Sys& s = sys[b];
        if (freeze && !s.go) return;
        const T dold = (T)s.s[S_DELTA], dnew = (T)tot[0];
        s.s[S_DELTA] = (double)dnew;
        s.s[S_BETA] = (double)safe_div(dnew, dold);
        s.rsq = (double)dnew;
    }
    __device__ void finalize_all() const {
        check_all<T>(*this, rules);
        if (threadIdx.x == 0) glob->pass = glob->pass + 1;
    }
};

// d = s + beta d
template <typename T, int PRE>
struct CgDirection : FBase {
    static constexpr int NA = 3;   // d, r|s, w
    static constexpr unsigned RMASK = PRE == 1 ? 0b111u : 0b011u, WMASK = 0b001u, SHARED = 0b100u;
    static constexpr int ND = 0;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    struct Scal { T beta; };
    __device__ Scal load(int b) const { return Scal{(T)sys[b].s[S_BETA]}; }
    __device__ void elem(T (&e)[NA], double (&)[1], const Scal& sc) const {
        const T s = PRE == 1 ? mul_rn(e[1], e[2]) : e[1];
        e[0] = add_rn(s, mul_rn(sc.beta, e[0]));
    }
    __device__ void finalize(int, const double (&)[1]) const {}
    __device__ void finalize_all() const {}
};

// r -= alpha q ; s = M r ; delta' = r.s — second half of the fused stencil iteration (x is updated by the next k_star
// pass together with the direction).  PRE: 0 none, 1 Jacobi vector, 3 Jacobi constant.
template <typename T, int PRE>
struct CgResid : FBase {
    static constexpr int NA = 3;   // r, q, w
    static constexpr unsigned RMASK = PRE == 1 ? 0b111u : 0b011u, WMASK = 0b001u, SHARED = 0b100u;
    static constexpr int ND = 1;
    static constexpr bool GLOBAL = true;
    const void* v[NA];
    T cinv;
    int rules;
    PeerHalo peer;     // slab mode: the new residual's boundary planes also go to the neighbours' halo planes
    struct Scal { T alpha; };
    __device__ Scal load(int b) const { return Scal{(T)sys[b].s[S_ALPHA]}; }
    __device__ void elem(T (&e)[NA], double (&acc)[1], const Scal& sc) const {
        e[0] = add_rn(e[0], -mul_rn(sc.alpha, e[1]));
        const T s = PRE == 1 ? mul_rn(e[0], e[2]) : (PRE == 3 ? mul_rn(e[0], cinv) : e[0]);
        acc[0] += (double)mul_rn(e[0], s);
    }
    __device__ void finalize(int b, const double (&tot)[1]) const {
        Sys& s = sys[b];
        if (freeze && !s.go) return;
        const T dold = (T)s.s[S_DELTA], dnew = (T)tot[0];
        s.s[S_DELTA] = (double)dnew;
        s.s[S_BETA] = (double)safe_div(dnew, dold);
        s.rsq = (double)dnew;
    }
    __device__ void finalize_all() const {
        check_all<T>(*this, rules);
        if (threadIdx.x == 0) glob->pass = glob->pass + 1;
    }
};

// The same second half for a pattern-coded star (coded.cuh, star_diag): Jacobi's inverse diagonal comes from the pattern table through the row's
// code byte (N bytes of traffic instead of the N w of an inverse-diagonal vector).  The table values are the 1 / d the vector would hold.
template <typename T>
struct CgResidCoded : FBase {
    static constexpr int NA = 2;   // r, q
    static constexpr unsigned RMASK = 0b11u, WMASK = 0b01u, SHARED = 0u;
    static constexpr int ND = 1;
    static constexpr bool GLOBAL = true;
    static constexpr bool per_index = true;
    const void* v[NA];
    const uint8_t* code;
    const T* inv_tab;
    int rules;
    struct Scal { T alpha; T inv0; };
    __device__ Scal load(int b) const { return Scal{(T)sys[b].s[S_ALPHA], inv_tab[0]}; }
    __device__ void elem(T (&e)[NA], double (&acc)[1], const Scal& sc, int64_t i) const {
        e[0] = add_rn(e[0], -mul_rn(sc.alpha, e[1]));
        const unsigned c = code[i];
        const T s = mul_rn(e[0], c == 0u ? sc.inv0 : __ldg(inv_tab + c));
        acc[0] += (double)mul_rn(e[0], s);
    }
    __device__ void finalize(int b, const double (&tot)[1]) const {
        Sys& s = sys[b];
        if (freeze && !s.go) return;
        const T dold = (T)s.s[S_DELTA], dnew = (T)tot[0];
        s.s[S_DELTA] = (double)dnew;
        s.s[S_BETA] = (double)safe_div(dnew, dold);
        s.rsq = (double)dnew;
    }
    __device__ void finalize_all() const {
        check_all<T>(*this, rules);
        if (threadIdx.x == 0) glob->pass = glob->pass + 1;
    }
};

// x += alpha d_cur for every system (pending update of the fused iteration), then alpha = 0
template <typename T>
__global__ void __launch_bounds__(kBlock) k_flush_x(T* __restrict__ X, const T* __restrict__ buf_a, const T* __restrict__ buf_b,
                                                    const Sys* __restrict__ sys, const Glob* __restrict__ glob, int64_t n, int64_t ld, int flip) {
    const int b = blockIdx.y;
    if (flip && glob->any_go == 0) return;     // mid-pass flush of a pass that is skipped (everything finished): stays pending
    const T alpha = (T)sys[b].s[S_ALPHA];
    if (alpha == T(0)) return;
    // flip: called in the middle of a pass (torch refresh), the pass counter does not yet include the direction just written
    const T* __restrict__ d = ((((glob->pass & 1) != 0) != (flip != 0)) ? buf_b : buf_a) + (int64_t)b * ld;
    T* __restrict__ x = X + (int64_t)b * ld;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) x[i] = add_rn(x[i], mul_rn(alpha, d[i]));
}

__global__ void k_zero_alpha(Sys* sys, int batch, const Glob* glob, int only_if_active) {
    if (only_if_active && glob->any_go == 0) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) sys[b].s[S_ALPHA] = 0.0;
}

// ------------------------------------------------------------------------------------------------ torch refresh pass
// x += alpha d only
template <typename T>
struct AxpyX : FBase {
    static constexpr int NA = 2;   // x, d
    static constexpr unsigned RMASK = 0b11u, WMASK = 0b01u, SHARED = 0u;
    static constexpr int ND = 0;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    struct Scal { T alpha; };
    __device__ Scal load(int b) const { return Scal{(T)sys[b].s[S_ALPHA]}; }
    __device__ void elem(T (&e)[NA], double (&)[1], const Scal& sc) const { e[0] = add_rn(e[0], mul_rn(sc.alpha, e[1])); }
    __device__ void finalize(int, const double (&)[1]) const {}
    __device__ void finalize_all() const {}
};

// SpMV epilogue that only carries the idle test (t = A x in the refresh pass)
template <typename T>
struct PlainEpi : FBase {
    static constexpr int ND = 0;
    static constexpr bool GLOBAL = false;
    __device__ void row(int, int64_t, T, double (&)[1]) const {}
    __device__ void finalize(int, const double (&)[1]) const {}
    __device__ void finalize_all() const {}
};

// r = y - t ; dots r.r, r.q ; function_evaluations += 1 for ALL systems (torch_sparse_cg, :1372-1373)
template <typename T, bool ADAPTIVE>
struct RefreshResidual : FBase {
    static constexpr int NA = 4;   // r, y, t, q
    static constexpr unsigned RMASK = ADAPTIVE ? 0b1110u : 0b0110u, WMASK = 0b0001u, SHARED = 0u;
    static constexpr int ND = 2;
    static constexpr bool GLOBAL = !ADAPTIVE;
    const void* v[NA];
    int rules;
    struct Scal {};
    __device__ Scal load(int) const { return Scal{}; }
    __device__ void elem(T (&e)[NA], double (&acc)[2], const Scal&) const {
        e[0] = add_rn(e[1], -e[2]);
        acc[0] += (double)mul_rn(e[0], e[0]);
        if (ADAPTIVE) acc[1] += (double)mul_rn(e[0], e[3]);
    }
    __device__ void finalize(int b, const double (&tot)[2]) const {
        Sys& s = sys[b];
        s.fev += 1;
        const T dold = (T)s.s[S_DELTA], dnew = (T)tot[0];
        s.s[S_DELTA] = (double)dnew;
        s.s[S_BETA] = (double)safe_div(dnew, dold);
        s.s[S_RQ] = (double)(T)tot[1];
        s.rsq = (double)dnew;
        if (ADAPTIVE) s.its += s.go;
    }
    __device__ void finalize_all() const {
        check_all<T>(*this, rules);
        if (threadIdx.x == 0) glob->pass = glob->pass + 1;
    }
};

// ------------------------------------------------------------------------------------------------ CG-adaptive
// x += alpha d ; r -= alpha q ; dots r.r, r.q
template <typename T>
struct AdUpdate : FBase {
    static constexpr int NA = 4;   // x, r, d, q
    static constexpr unsigned RMASK = 0b1111u, WMASK = 0b0011u, SHARED = 0u;
    static constexpr int ND = 2;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    struct Scal { T alpha; };
    __device__ Scal load(int b) const { return Scal{(T)sys[b].s[S_ALPHA]}; }
    __device__ void elem(T (&e)[NA], double (&acc)[2], const Scal& sc) const {
        e[0] = add_rn(e[0], mul_rn(sc.alpha, e[2]));
        e[1] = add_rn(e[1], -mul_rn(sc.alpha, e[3]));
        acc[0] += (double)mul_rn(e[1], e[1]);
        acc[1] += (double)mul_rn(e[1], e[3]);
    }
    __device__ void finalize(int b, const double (&tot)[2]) const {
        Sys& s = sys[b];
        if (freeze && !s.go) return;
        s.its += s.go;
        s.s[S_DELTA] = (double)(T)tot[0];
        s.s[S_RQ] = (double)(T)tot[1];
        s.rsq = (double)(T)tot[0];
    }
    __device__ void finalize_all() const {}
};

// r -= alpha q ; dots r.r, r.q — residual half of the fused adaptive iteration (x is updated by the stencil kernel that follows)
template <typename T>
struct AdResid : FBase {
    static constexpr int NA = 2;   // r, q
    static constexpr unsigned RMASK = 0b11u, WMASK = 0b01u, SHARED = 0u;
    static constexpr int ND = 2;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    struct Scal { T alpha; };
    __device__ Scal load(int b) const { return Scal{(T)sys[b].s[S_ALPHA]}; }
    __device__ void elem(T (&e)[NA], double (&acc)[2], const Scal& sc) const {
        e[0] = add_rn(e[0], -mul_rn(sc.alpha, e[1]));
        acc[0] += (double)mul_rn(e[0], e[0]);
        acc[1] += (double)mul_rn(e[0], e[1]);
    }
    __device__ void finalize(int b, const double (&tot)[2]) const {
        Sys& s = sys[b];
        if (freeze && !s.go) return;
        s.its += s.go;
        s.s[S_DELTA] = (double)(T)tot[0];
        s.s[S_RQ] = (double)(T)tot[1];
        s.rsq = (double)(T)tot[0];
    }
    __device__ void finalize_all() const {}
};

// d = r - safe_div(rq * d, dq)
template <typename T>
struct AdDirection : FBase {
    static constexpr int NA = 2;   // d, r
    static constexpr unsigned RMASK = 0b11u, WMASK = 0b01u, SHARED = 0u;
    static constexpr int ND = 0;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    struct Scal { T rq, dq; };
    __device__ Scal load(int b) const { return Scal{(T)sys[b].s[S_RQ], (T)sys[b].s[S_DQ]}; }
    __device__ void elem(T (&e)[NA], double (&)[1], const Scal& sc) const {
        const T t = mul_rn(sc.rq, e[0]);
        const T u = sc.dq == T(0) ? T(0) : t / sc.dq;
        e[0] = add_rn(e[1], -u);
    }
    __device__ void finalize(int, const double (&)[1]) const {}
    __device__ void finalize_all() const {}
};

// q = A d with d.q, d.r ; then progress check and the next pass's alpha = safe_div(d.r, d.q) * go
template <typename T>
struct AdDots : FBase {   // SpMV epilogue
    static constexpr int ND = 2;
    static constexpr bool GLOBAL = true;
    const T* d;
    const T* r;
    int64_t ld;
    int rules;
    int do_check;    // 0 for the product inside start(): the first check was already made
    int count_fev;
    __device__ void row(int b, int64_t i, T y, double (&acc)[2]) const {
        const T dv = d[(int64_t)b * ld + i];
        acc[0] += (double)mul_rn(dv, y);
        acc[1] += (double)mul_rn(dv, r[(int64_t)b * ld + i]);
    }
    __device__ void finalize(int b, const double (&tot)[2]) const {
        Sys& s = sys[b];
        if (freeze && !always && !s.go) return;
        if (count_fev) s.fev += s.go;
        s.s[S_DQ] = (double)(T)tot[0];
        s.s[S_DR] = (double)(T)tot[1];
    }
    __device__ void finalize_all() const {
        if (do_check) check_all<T>(*this, rules);
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            volatile Sys* s = sys + b;
            s->s[S_ALPHA] = s->go ? (double)safe_div((T)s->s[S_DR], (T)s->s[S_DQ]) : 0.0;
        }
        if (threadIdx.x == 0 && do_check) glob->pass = glob->pass + 1;
    }
};

// ------------------------------------------------------------------------------------------------ BiCGstab primitives
// out = a + sign * s[slot] * b
template <typename T>
struct Lin2 : FBase {
    static constexpr int NA = 3;   // out, a, b
    static constexpr unsigned RMASK = 0b110u, WMASK = 0b001u, SHARED = 0u;
    static constexpr int ND = 0;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    int slot;
    int sign;
    struct Scal { T c; };
    __device__ Scal load(int b) const { return Scal{(T)(sign * sys[b].s[slot])}; }
    __device__ void elem(T (&e)[NA], double (&)[1], const Scal& sc) const { e[0] = add_rn(e[1], mul_rn(sc.c, e[2])); }
    __device__ void finalize(int, const double (&)[1]) const {}
    __device__ void finalize_all() const {}
};

// out = a .* w (w shared along the batch): Jacobi apply
template <typename T>
struct Scale : FBase {
    static constexpr int NA = 3;   // out, a, w
    static constexpr unsigned RMASK = 0b110u, WMASK = 0b001u, SHARED = 0b100u;
    static constexpr int ND = 0;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    struct Scal {};
    __device__ Scal load(int) const { return Scal{}; }
    __device__ void elem(T (&e)[NA], double (&)[1], const Scal&) const { e[0] = mul_rn(e[1], e[2]); }
    __device__ void finalize(int, const double (&)[1]) const {}
    __device__ void finalize_all() const {}
};

enum BiStep {
    BI_RHO = 0,      // rho1 = tot0; (j==0: its += go, rho0 = -omega rho0); beta = alpha rho1 / rho0; rho0 = rho1
    BI_GAMMA,        // fev += go; alpha = rho0 / tot0
    BI_NEXT,         // fev += go; if (j+1 < L) like BI_RHO with j+1
    BI_TAU,          // tau[j] = tot0 / sigma[i]
    BI_SIGMA,        // sigma[j] = tot0; gam_p[j] = tot1 / sigma[j]; j == L: MR coefficients
    BI_FINAL,        // rsq = tot0 -> progress check
    B1_RHO,          // its += go; rho = tot0; beta = rho / rho_prev * alpha / omega; rho_prev = rho
    B1_ALPHA,        // fev += go; alpha = rho / tot0
    B1_OMEGA,        // omega = tot0 / tot1
    B1_FEV           // fev += go
};

template <typename T>
__device__ __forceinline__ void bi_scalar_step(Sys& s, int step, int j, int i, int L, const double* tot) {
    double* v = s.s;
    switch (step) {
        case BI_RHO: {
            if (j == 0) {
                s.its += s.go;
                v[S_RHO0] = (double)(-(T)v[S_OMEGA] * (T)v[S_RHO0]);
            }
            const T rho1 = (T)tot[0];
            v[S_RHO1] = (double)rho1;
            v[S_BETA] = (double)((T)v[S_ALPHA] * rho1 / (T)v[S_RHO0]);
            v[S_RHO0] = (double)rho1;
            break;
        }
        case BI_GAMMA: {
            s.fev += s.go;
            v[S_GAMMA] = (double)(T)tot[0];
            v[S_ALPHA] = (double)((T)v[S_RHO0] / (T)tot[0]);
            break;
        }
        case BI_NEXT: {
            s.fev += s.go;
            if (j + 1 < L) {
                const T rho1 = (T)tot[0];
                v[S_RHO1] = (double)rho1;
                v[S_BETA] = (double)((T)v[S_ALPHA] * rho1 / (T)v[S_RHO0]);
                v[S_RHO0] = (double)rho1;
            }
            break;
        }
        case BI_TAU: {
            v[S_TAU + j] = (double)((T)tot[0] / (T)v[S_SIGMA + i]);
            break;
        }
        case BI_SIGMA: {
            const T sig = (T)tot[0];
            v[S_SIGMA + j] = (double)sig;
            v[S_GAMP + j] = (double)((T)tot[1] / sig);
            if (j == L) {   // MR part (tau rows are aliased in the reference: tau[i][j] == tau[j])
                v[S_OMEGA] = v[S_GAM + L] = v[S_GAMP + L];
                for (int jj = L - 1; jj >= 1; --jj) {
                    T acc = T(0);
                    for (int ii = jj + 1; ii <= L; ++ii) acc = acc + (T)v[S_TAU + ii] * (T)v[S_GAM + ii];
                    v[S_GAM + jj] = (double)((T)v[S_GAMP + jj] - acc);
                }
                for (int jj = 1; jj < L; ++jj) {
                    T acc = T(0);
                    for (int ii = jj + 1; ii < L; ++ii) acc = acc + (T)v[S_TAU + ii] * (T)v[S_GAM + ii + 1];
                    v[S_GAMPP + jj] = (double)((T)v[S_GAM + jj + 1] + acc);
                }
            }
            break;
        }
        case BI_FINAL: {
            s.rsq = (double)(T)tot[0];
            break;
        }
        case B1_RHO: {
            s.its += s.go;
            const T rho = (T)tot[0];
            v[S_RHO1] = (double)rho;
            v[S_BETA] = (double)(rho / (T)v[S_RHOP] * (T)v[S_ALPHA] / (T)v[S_OMEGA]);
            v[S_RHOP] = (double)rho;
            break;
        }
        case B1_ALPHA: {
            s.fev += s.go;
            v[S_ALPHA] = (double)((T)v[S_RHO1] / (T)tot[0]);
            break;
        }
        case B1_OMEGA: {
            v[S_OMEGA] = (double)((T)tot[0] / (T)tot[1]);
            break;
        }
        case B1_FEV: {
            s.fev += s.go;
            break;
        }
    }
}

// two dot products a0.b0, a1.b1 followed by a scalar step
template <typename T>
struct Dot2 : FBase {
    static constexpr int NA = 4;   // a0, b0, a1, b1
    static constexpr unsigned RMASK = 0b1111u, WMASK = 0u, SHARED = 0u;
    static constexpr int ND = 2;
    static constexpr bool GLOBAL = true;
    const void* v[NA];
    int step, j, i, L, rules;
    struct Scal {};
    __device__ Scal load(int) const { return Scal{}; }
    __device__ void elem(T (&e)[NA], double (&acc)[2], const Scal&) const {
        acc[0] += (double)mul_rn(e[0], e[1]);
        acc[1] += (double)mul_rn(e[2], e[3]);
    }
    __device__ void finalize(int b, const double (&tot)[2]) const { bi_scalar_step<T>(sys[b], step, j, i, L, tot); }
    __device__ void finalize_all() const {
        if (step == BI_FINAL) {
            check_all<T>(*this, rules);
            if (threadIdx.x == 0) glob->pass = glob->pass + 1;
        }
    }
};

// SpMV epilogue: dots y.w0 and (optional) y.w1, then a scalar step
template <typename T>
struct BiEpi : FBase {
    static constexpr int ND = 2;
    static constexpr bool GLOBAL = false;
    const T* w0;
    const T* w1;   // may be null
    int64_t ld;
    int step, j, i, L;
    __device__ void row(int b, int64_t k, T y, double (&acc)[2]) const {
        acc[0] += (double)mul_rn(y, w0[(int64_t)b * ld + k]);
        if (w1) acc[1] += (double)mul_rn(y, w1[(int64_t)b * ld + k]);
    }
    __device__ void finalize(int b, const double (&tot)[2]) const { bi_scalar_step<T>(sys[b], step, j, i, L, tot); }
    __device__ void finalize_all() const {}
};

// K independent updates in one pass (BiCGstab(L) inner loops), same arithmetic as K separate Lin2 launches:
//   MODE 0: t_i = s_i + c t_i      (uh[i] = rh[i] - beta uh[i])
//   MODE 1: t_i = t_i + c s_i      (rh[i] = rh[i] - alpha uh[i+1]);  WITH_X adds x = x + cx u  (x += alpha uh[0])
// vectors: t_0, s_0, t_1, s_1, ... [, x, u]
template <typename T, int K, int MODE, bool WITH_X>
struct LinMulti : FBase {
    static constexpr int NA = 2 * K + (WITH_X ? 2 : 0);
    static constexpr unsigned RMASK = (1u << NA) - 1u;
    static constexpr unsigned WMASK = (K >= 1 ? 1u : 0u) | (K >= 2 ? 4u : 0u) | (K >= 3 ? 16u : 0u) | (WITH_X ? (1u << (2 * K)) : 0u);
    static constexpr unsigned SHARED = 0u;
    static constexpr int ND = 0;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    int slot, sign, xslot, xsign;
    struct Scal { T c, cx; };
    __device__ Scal load(int b) const { return Scal{(T)(sign * sys[b].s[slot]), WITH_X ? (T)(xsign * sys[b].s[xslot]) : T(0)}; }
    __device__ void elem(T (&e)[NA], double (&)[1], const Scal& sc) const {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            if (MODE == 0) e[2 * i] = add_rn(e[2 * i + 1], mul_rn(sc.c, e[2 * i]));
            else e[2 * i] = add_rn(e[2 * i], mul_rn(sc.c, e[2 * i + 1]));
        }
        if (WITH_X) e[2 * K] = add_rn(e[2 * K], mul_rn(sc.cx, e[2 * K + 1]));
    }
    __device__ void finalize(int, const double (&)[1]) const {}
    __device__ void finalize_all() const {}
};

// t = t + c s, then the two dot products of the BI_SIGMA step on the updated t:  t.t and p.t   (vectors t, s, p)
template <typename T>
struct LinSigma : FBase {
    static constexpr int NA = 3;
    static constexpr unsigned RMASK = 0b111u, WMASK = 0b001u, SHARED = 0u;
    static constexpr int ND = 2;
    static constexpr bool GLOBAL = false;
    const void* v[NA];
    int slot, sign, j, L;
    struct Scal { T c; };
    __device__ Scal load(int b) const { return Scal{(T)(sign * sys[b].s[slot])}; }
    __device__ void elem(T (&e)[NA], double (&acc)[2], const Scal& sc) const {
        e[0] = add_rn(e[0], mul_rn(sc.c, e[1]));
        acc[0] += (double)mul_rn(e[0], e[0]);
        acc[1] += (double)mul_rn(e[2], e[0]);
    }
    __device__ void finalize(int b, const double (&tot)[2]) const { bi_scalar_step<T>(sys[b], BI_SIGMA, j, 0, L, tot); }
    __device__ void finalize_all() const {}
};

// Last block of a BiCGstab(2) pass in one sweep (six Lin2 launches + the final dot of the unfused form, same order of
// operations per element): vectors x, rh0, rh1, rh2, uh0, uh1, uh2
template <typename T>
struct BiFinal2 : FBase {
    static constexpr int NA = 7;
    static constexpr unsigned RMASK = 0b1111111u, WMASK = 0b0010011u, SHARED = 0u;
    static constexpr int ND = 1;
    static constexpr bool GLOBAL = true;
    const void* v[NA];
    int rules;
    struct Scal { T g1, gp2, g2, gpp1, gp1; };
    __device__ Scal load(int b) const {
        const double* q = sys[b].s;
        return Scal{(T)q[S_GAM + 1], (T)(-q[S_GAMP + 2]), (T)(-q[S_GAM + 2]), (T)q[S_GAMPP + 1], (T)(-q[S_GAMP + 1])};
    }
    __device__ void elem(T (&e)[NA], double (&acc)[1], const Scal& sc) const {
        e[0] = add_rn(e[0], mul_rn(sc.g1, e[1]));        // x   += gam[1]    rh0
        e[1] = add_rn(e[1], mul_rn(sc.gp2, e[3]));       // rh0 -= gam_p[2]  rh2
        e[4] = add_rn(e[4], mul_rn(sc.g2, e[6]));        // uh0 -= gam[2]    uh2
        e[4] = add_rn(e[4], mul_rn(-sc.g1, e[5]));       // uh0 -= gam[1]    uh1
        e[0] = add_rn(e[0], mul_rn(sc.gpp1, e[2]));      // x   += gam_pp[1] rh1
        e[1] = add_rn(e[1], mul_rn(sc.gp1, e[2]));       // rh0 -= gam_p[1]  rh1
        acc[0] += (double)mul_rn(e[1], e[1]);
    }
    __device__ void finalize(int b, const double (&tot)[1]) const {
        double t2[2] = {tot[0], tot[0]};
        bi_scalar_step<T>(sys[b], BI_FINAL, 0, 0, 2, t2);
    }
    __device__ void finalize_all() const {
        check_all<T>(*this, rules);
        if (threadIdx.x == 0) glob->pass = glob->pass + 1;
    }
};

// multi-GPU batch sharding (or a batch split over several solver objects): mode 2 installs the sum of |y|^2 over ALL systems
// (vals[1]), recomputes this shard's tolerances under the batch-sum rule and its local minimum; mode 1 installs the batch-wide
// minimum tolerance (vals[0]) and repeats the first progress check with it.
template <typename T>
__global__ void __launch_bounds__(kBlock) k_recheck(FBase f, int rules, const double* __restrict__ vals, int adaptive, int mode, const T* rtol, const T* atol, int sum_rule) {
    if (mode == 2) {
        if (!sum_rule) return;
        __shared__ double s_min[kBlock];
        apply_batch_sum<T>(f.sys, f.batch, rtol, atol, vals[1]);
        __syncthreads();
        double m = INFINITY;
        bool nan = false;
        for (int b = threadIdx.x; b < f.batch; b += blockDim.x) {
            const double t = ((volatile Sys*)f.sys)[b].tol;
            if (t != t) nan = true;
            else m = fmin(m, t);
        }
        s_min[threadIdx.x] = nan ? NAN : m;
        __syncthreads();
        if (threadIdx.x == 0) {
            double mm = INFINITY;
            bool anynan = false;
            for (int k = 0; k < (int)blockDim.x; ++k) {
                if (s_min[k] != s_min[k]) anynan = true;
                else mm = fmin(mm, s_min[k]);
            }
            f.glob->tol_min = anynan ? NAN : mm;
            f.glob->yy_sum = vals[1];
        }
        return;
    }
    if (threadIdx.x == 0) {
        f.glob->tol_min = vals[0];
        f.glob->checks = 0;
        __threadfence();
    }
    __syncthreads();
    check_all<T>(f, rules);
    __syncthreads();
    if (adaptive) {   // the first step size was computed from the flags of the local check
        for (int b = threadIdx.x; b < f.batch; b += blockDim.x) {
            volatile Sys* s = f.sys + b;
            s->s[S_ALPHA] = s->go ? (double)safe_div((T)s->s[S_DR], (T)s->s[S_DQ]) : 0.0;
        }
    }
}

__global__ void k_collect(const Sys* __restrict__ sys, int batch, int32_t* its, int32_t* fev, uint8_t* conv, uint8_t* div) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (its) its[b] = sys[b].its;
    if (fev) fev[b] = sys[b].fev;
    if (conv) conv[b] = (uint8_t)sys[b].conv;
    if (div) div[b] = (uint8_t)sys[b].div;
}

// Generic cg / cg_adaptive (phiml/backend/_linalg.py:52-128): a system that stopped on a NON-FINITE monitored value (NaN / Inf in y or x0, overflow)
// is not left alone by the reference while other systems of the batch keep the loop alive — its step size is divide_no_nan(delta, dx.dy) * continue_
// = (non-finite) * 0 = NaN (:74-77, :112-114), so x += step_size * dx and residual -= step_size * dy turn the whole system NaN on the very next pass.
// The device loop skips stopped systems; this kernel writes the NaNs the reference would have produced (rows [off, off + n) of system blockIdx.y).
template <typename T>
__global__ void __launch_bounds__(256) k_poison_nonfinite(const Sys* __restrict__ sys, int batch, T* __restrict__ X, T* __restrict__ R, int64_t ld, int64_t off, int64_t n) {
    const int b = blockIdx.y;
    const Sys& me = sys[b];
    if (!me.div || isfinite(me.rsq)) return;
    __shared__ int s_more;
    if (threadIdx.x == 0) s_more = 0;
    __syncthreads();
    int more = 0;
    for (int k = threadIdx.x; k < batch; k += blockDim.x) more |= sys[k].its > me.its ? 1 : 0;      // did the loop make another pass after this system stopped?
    if (more) atomicOr(&s_more, 1);
    __syncthreads();
    if (!s_more) return;
    const T nan = (T)NAN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        X[(int64_t)b * ld + off + i] = nan;
        R[(int64_t)b * ld + off + i] = nan;
    }
}

template <typename T>
__global__ void k_fill(T* p, int64_t n, T value) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = value;
}

}  // namespace phb

using namespace phb;

// ================================================================================================== solver object
struct phb200_solver {
    const phb200_matrix* A;
    const phb200_precond* M;
    int method, L, dtype, pre;
    int64_t batch, n;
    // device bookkeeping (owned)
    Sys* sys;
    Glob* glob;
    double* partials;
    unsigned* tickets;
    int* h_flag;               // pinned, 2 slots
    void* scratch;             // one device block holding sys / glob / partials / tickets (recycled through g_scratch_pool)
    size_t scratch_bytes;
    int device;
    cudaEvent_t ev[2];
    // per-solve bindings
    const void* Y;
    void* X;
    void* R;
    char* ws;
    size_t ws_bytes;
    int started;
    int64_t passes;            // loop passes enqueued so far
    int64_t launches0;
    long long launches;
    // optional per-kernel timing (bench.py): events around the k-th kernel of every pass
    int prof;
    int prof_slot;             // index of the next kernel within the current pass
    std::vector<cudaEvent_t>* prof_events;   // pairs (start, stop)
    std::vector<int>* prof_slots;
    // TMA descriptors of r and the two direction buffers (fused stencil CG); valid for the vectors bound by start()
    CUtensorMap tm_r, tm_a, tm_b;
    int tma_ok;
    int coded_star_off;     // coded star matrix, but no tensor maps could be built: generic phases + coded product
    int tma_occ;               // resident CTAs per SM of the TMA stencil kernel (decides its x-chunking)
    // slab / multi-GPU mode: kernels store local sums in `totals`, the host all-reduces, dist_step() finalises
    double* totals;            // device, caller-owned (batch x kMaxDots doubles); non-null = distributed mode
    const void* rtol;
    const void* atol;
    const int32_t* max_iter;
    // callback-driven distributed mode (any method): the library calls these between its kernels
    phb200_allreduce_fn ar_cb;
    phb200_halo_fn halo_cb;
    void* cb_user;
    PeerReduce* peer_reduce;   // device table for the in-kernel all-reduce (phb200_solver_set_peer_reduce); owned
    unsigned long long* peer_seq;   // device, owned
    void* peer_lo;             // slab mode: the neighbours' halo planes of R (CUDA IPC mappings), see phb200_solver_set_peer_halos
    void* peer_hi;
    int64_t peer_ld_lo, peer_ld_hi;
    // native distributed mode (phb200_solver_set_peer_native): any method on x-slabs with NO host in the loop — reductions all-reduce
    // inside the kernels (peer_reduce), boundary planes of every product's input are pushed into the neighbours' arrays by k_halo_push
    int native_peer;
    phb200_peer_slab nb_lo, nb_hi;          // the slab neighbours' X / R / workspace (CUDA IPC mappings), leading dimension, halo-plane offset, flag word
    unsigned long long* halo_flags;         // device, caller-owned, zero-filled: [0] written by the lower neighbour, [1] by the upper one
    unsigned long long halo_seq;            // exchanges made so far (host counter; identical on all ranks)
    unsigned* halo_ticket;                  // device, owned
    int resident;              // PHB200_OPT_RESIDENT: -1 cost model decides, 0 never, 1 whenever the system fits on chip
    int resident_used;         // 1 if the last run() went through the resident kernel
    int64_t off, n_act;        // element range [off, off + n_act) of every vector that this rank owns (whole vector if not a slab)
};

namespace phb {

// Upper bound of gridDim.x over every kernel of a solve: grid_x() hands out at most kTargetCtas CTAs per launch,
// spread over the gridDim.y rows (systems, or tiles of kCsrBt systems in the streaming CSR kernel).
static int max_blocks_for(int64_t batch) {
    const int64_t rows = batch >= 8 * kCsrBt ? (batch + kCsrBt - 1) / kCsrBt : batch;
    int64_t mb = kTargetCtas / (rows < 1 ? 1 : rows);
    return (int)(mb < 1 ? 1 : mb);
}

static Ctx make_ctx(const phb200_solver* s) {
    Ctx c;
    c.sys = s->sys;
    c.glob = s->glob;
    c.red.partials = s->partials;
    c.red.tickets = s->tickets;
    c.red.max_blocks = max_blocks_for(s->batch);
    c.red.peer = s->peer_reduce;
    c.red.peer_data = 0;
    c.defer = nullptr;
    c.batch = (int)s->batch;
    return c;
}

template <class F>
static void fill_base(F& f, const phb200_solver* s, bool always, bool freeze) {
    Ctx c = make_ctx(s);
    f.sys = c.sys;
    f.glob = c.glob;
    f.red = c.red;
    f.batch = c.batch;
    f.defer = nullptr;
    f.always = always ? 1 : 0;
    f.freeze = freeze ? 1 : 0;
}

// preconditioners applied by kernels of their own between the phases of a pass (not fusable into the stencil / phase kernels)
static inline bool pre_applied(int pre) { return pre == PHB200_PRE_ILU || pre == PHB200_PRE_CLUSTER; }

static int vectors_needed(int method, int L, int pre) {
    switch (method) {
        case PHB200_CG: return pre_applied(pre) ? 4 : 3;              // d, q, d' (fused ping-pong) | d, q, s, tmp
        case PHB200_CG_ADAPTIVE: return 3;                                   // d, q, d' (fused ping-pong)
        case PHB200_CG_TORCH: return 4;                                      // d, q, d' (fused ping-pong) | t, t
        case PHB200_CG_ADAPTIVE_TORCH: return 4;                             // d, q, d', t
        case PHB200_BICGSTAB:
            if (L == 1) return 10;                                           // ones, u, p, ph, z, t, sl, tl, tmp, tmp2
            return 2 * L + 4;                                                // shadow, uh[0..L], rh[1..L], pv, tmp
    }
    return 0;
}

// Native halo exchange of the distributed mode: every rank stores the first / last owned plane of `v` into the slab neighbours' halo planes
// of THEIR copy of the same vector (peer memory, plain coalesced stores over NVLink), fences at system scope, and the last CTA tells the
// neighbours "exchange number seq has landed" with one 8-byte word each, then waits for theirs.  The product that follows on the stream
// therefore reads current halo planes without any host involvement.  Skipped (like every other kernel of a pass) once all systems finished.
template <typename T>
__global__ void __launch_bounds__(kBlock) k_halo_push(const T* __restrict__ v, int64_t ld, int64_t first_off, int64_t last_off, int64_t plane, int batch,
                                                      T* lo_dst, int64_t lo_ld, T* hi_dst, int64_t hi_ld,
                                                      unsigned long long* lo_flag, unsigned long long* hi_flag, const unsigned long long* my_flags,
                                                      unsigned long long seq, unsigned* ticket, const Glob* glob, int always) {
    if (!always && glob->any_go == 0) return;
    const int64_t total = plane * batch;
    bool wrote = false;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < total; i += (int64_t)gridDim.x * kBlock) {
        const int64_t b = i / plane, o = i - b * plane;
        if (lo_dst) { lo_dst[b * lo_ld + o] = v[b * ld + first_off + o]; wrote = true; }
        if (hi_dst) { hi_dst[b * hi_ld + o] = v[b * ld + last_off + o]; wrote = true; }
    }
    if (wrote) __threadfence_system();
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    *ticket = 0u;
    __threadfence_system();
    if (lo_flag) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(lo_flag), "l"(seq) : "memory");
    if (hi_flag) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(hi_flag), "l"(seq) : "memory");
    const long long t0 = clock64();
    for (int side = 0; side < 2; ++side) {
        if (!(side == 0 ? lo_flag : hi_flag)) continue;          // no neighbour on that side
        unsigned long long got = 0;
        while (true) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(my_flags + side) : "memory");
            if (got >= seq) break;
            if (clock64() - t0 > 8000000000ll) __trap();
        }
    }
    __threadfence_system();
}

template <typename T>
struct Run {
    phb200_solver* s;
    cudaStream_t st;
    int64_t n, ld;
    int batch;
    T* vec(int k) const { return (T*)s->ws + (int64_t)k * batch * ld; }
    double* defer() const { return s->peer_reduce ? nullptr : s->totals; }   // in-kernel all-reduce: nothing is deferred
    bool frozen() const { return s->method == PHB200_CG || s->method == PHB200_CG_ADAPTIVE; }
    int rules() const { return (s->method == PHB200_CG_TORCH || s->method == PHB200_CG_ADAPTIVE_TORCH) ? RULES_TORCH : RULES_GENERIC; }
    const T* invd() const { return s->pre == PHB200_PRE_JACOBI ? (const T*)s->M->inv_diag : nullptr; }

    void prof_begin() {
        if (!s->prof) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        s->prof_events->push_back(a);
        s->prof_events->push_back(b);
        s->prof_slots->push_back(s->prof_slot++);
        cudaEventRecord(a, st);
    }
    void prof_end() {
        if (s->prof) cudaEventRecord(s->prof_events->back(), st);
    }
    bool cb_mode() const { return s->ar_cb != nullptr || s->native_peer; }
    bool native() const { return s->native_peer != 0; }
    // all-reduce of the deferred local sums over the ranks, then the scalar step on every rank identically
    template <class F> int reduce_and_finalize(const F& f) {
        const int rc = s->ar_cb(s->totals, (int64_t)batch * kMaxDots, (void*)st, s->cb_user);
        if (rc != 0) { set_error("all-reduce callback failed (%d)", rc); return PHB200_ERR_CUDA; }
        return finalize_only(f);
    }
    template <class P> int phase(const P& p) {
        if constexpr (P::ND > 0) {
            if (cb_mode() && !native()) {        // (native mode: the kernel all-reduces and finalises by itself, red.peer is set)
                P q = p;
                q.defer = s->totals;
                prof_begin();
                int r = launch_phase<T, P>(q, s->n_act, ld, batch, st, s->off);
                prof_end();
                if (r != PHB200_OK) return r;
                return reduce_and_finalize(q);
            }
        }
        prof_begin();
        int r = launch_phase<T, P>(p, s->n_act, ld, batch, st, s->off);
        prof_end();
        return r;
    }
    template <class F> int finalize_only(const F& f) {
        PHB_LAUNCH((k_finalize<F>), 1, kBlock, 0, st, f, (const double*)s->totals);
        return PHB200_OK;
    }
    template <class E> int spmv(const T* x, T* y, const E& e) {
        if (native()) {
            PHB_TRY(native_halo(x, e.always != 0));
        } else if (cb_mode()) {
            const int rc = s->halo_cb((void*)x, (void*)st, s->cb_user);   // the neighbours' boundary planes of the input vector
            if (rc != 0) { set_error("halo callback failed (%d)", rc); return PHB200_ERR_CUDA; }
            if constexpr (E::ND > 0) {
                E q = e;
                q.defer = s->totals;
                prof_begin();
                int r = launch_spmv<T, E>(s->A, x, ld, y, ld, (const T*)nullptr, batch, q, st);
                prof_end();
                if (r != PHB200_OK) return r;
                return reduce_and_finalize(q);
            }
        }
        prof_begin();
        int r = launch_spmv<T, E>(s->A, x, ld, y, ld, (const T*)nullptr, batch, e, st);
        prof_end();
        return r;
    }

    // the same vector in a neighbour's memory: X, R or work vector k of ITS workspace (ranks may own different numbers of planes)
    T* peer_vector(const phb200_peer_slab& nb, const T* x) const {
        if (!nb.X) return nullptr;
        if ((const void*)x == s->X) return (T*)nb.X;
        if ((const void*)x == s->R) return (T*)nb.R;
        const int64_t k = (x - (const T*)s->ws) / ((int64_t)batch * ld);
        return (T*)nb.ws + k * batch * nb.ld;
    }
    int native_halo(const T* x, bool always) {
        const StarDesc& sd = s->A->star;
        const int64_t plane = sd.ny * sd.nz;
        T* lo = peer_vector(s->nb_lo, x);
        T* hi = peer_vector(s->nb_hi, x);
        if (lo) lo += s->nb_lo.halo_off;
        if (hi) hi += s->nb_hi.halo_off;
        s->halo_seq += 1;
        const int64_t total = plane * batch;
        const unsigned grid = (unsigned)std::min<int64_t>((total + kBlock - 1) / kBlock, kSMs * 2);
        PHB_LAUNCH(k_halo_push<T>, grid, kBlock, 0, st, x, ld, sd.xb * plane, (sd.xe - 1) * plane, plane, batch, lo, s->nb_lo.ld, hi, s->nb_hi.ld,
                   (unsigned long long*)s->nb_lo.flag, (unsigned long long*)s->nb_hi.flag, (const unsigned long long*)s->halo_flags, s->halo_seq, s->halo_ticket,
                   (const Glob*)s->glob, always ? 1 : 0);
        return PHB200_OK;
    }

    int lin2(T* out, const T* a, const T* b, int slot, int sign) {
        Lin2<T> f;
        fill_base(f, s, false, false);
        f.v[0] = out; f.v[1] = a; f.v[2] = b;
        f.slot = slot; f.sign = sign;
        return phase(f);
    }
    int dot2(const T* a0, const T* b0, const T* a1, const T* b1, int step, int j, int i) {
        Dot2<T> f;
        fill_base(f, s, false, false);
        f.v[0] = a0; f.v[1] = b0; f.v[2] = a1; f.v[3] = b1;
        f.step = step; f.j = j; f.i = i; f.L = s->L; f.rules = RULES_GENERIC;
        return phase(f);
    }
    int spmv_dot(const T* x, T* y, const T* w0, const T* w1, int step, int j) {
        BiEpi<T> e;
        fill_base(e, s, false, false);
        e.w0 = w0; e.w1 = w1; e.ld = ld; e.step = step; e.j = j; e.i = 0; e.L = s->L;
        return spmv(x, y, e);
    }
    // out = M^-1 in; returns the vector that holds the result (in itself without preconditioner)
    int apply_pre(const T* in, T* out, T* tmp, const T** result) {
        if (s->pre == PHB200_PRE_NONE) { *result = in; return PHB200_OK; }
        if (s->pre == PHB200_PRE_JACOBI) {
            Scale<T> f;
            fill_base(f, s, false, false);
            f.v[0] = out; f.v[1] = in; f.v[2] = s->M->inv_diag;
            PHB_TRY(phase(f));
            *result = out;
            return PHB200_OK;
        }
        PHB_TRY(precond_apply(s->M, in, out, tmp, batch, ld, st));
        *result = out;
        return PHB200_OK;
    }
    int apply_inv_l(const T* in, T* out, const T** result) {
        if (s->pre == PHB200_PRE_NONE) { *result = in; return PHB200_OK; }
        if (s->pre == PHB200_PRE_JACOBI) {
            Scale<T> f;
            fill_base(f, s, false, false);
            f.v[0] = out; f.v[1] = in; f.v[2] = s->M->inv_diag;
            PHB_TRY(phase(f));
            *result = out;
            return PHB200_OK;
        }
        PHB_TRY(precond_apply_inv_l(s->M, in, out, batch, ld, st));
        *result = out;
        return PHB200_OK;
    }

    // ------------------------------------------------------------------------------------------ start
    int start(const T* Y, const T* X0, T* X, T* R, const T* rtol, const T* atol, const int32_t* max_iter) {
        const int method = s->method;
        s->rtol = rtol; s->atol = atol; s->max_iter = max_iter;
        s->tma_ok = 0;
        s->coded_star_off = 0;
        if ((fused() || fused_adaptive()) && !getenv("PHB200_NO_TMA")) {
            const StarDesc& sd = s->A->star;
            s->tma_ok = make_star_tmap(&s->tm_r, R, s->dtype, sd, batch, ld) && make_star_tmap(&s->tm_a, vec(0), s->dtype, sd, batch, ld) &&
                        make_star_tmap(&s->tm_b, vec(2), s->dtype, sd, batch, ld);
        }
        if (s->A->kind == PHB200_MAT_CODED && !s->tma_ok) s->coded_star_off = 1;      // no tensor maps (alignment): the generic phases + coded product
        if (X != X0) PHB_CUDA(cudaMemcpyAsync(X, X0, sizeof(T) * (size_t)batch * ld, cudaMemcpyDeviceToDevice, st));
        PHB_CUDA(cudaMemsetAsync(s->tickets, 0, sizeof(unsigned) * (size_t)batch, st));
        PHB_CUDA(cudaMemsetAsync(s->glob, 0, sizeof(Glob), st));
        T* d = vec(0);
        T* q = vec(1);
        {   // q = A x0
            PlainEpi<T> e;
            fill_base(e, s, true, false);
            PHB_TRY(spmv(X, q, e));
        }
        const bool bicg = method == PHB200_BICGSTAB;
        if (pre_applied(s->pre) && method == PHB200_CG) {
            Residual<T> r0;
            fill_base(r0, s, true, false);
            r0.v[0] = R; r0.v[1] = Y; r0.v[2] = q;
            PHB_TRY(phase(r0));
            T* sv = vec(2);
            PHB_TRY(precond_apply(s->M, R, sv, vec(3), batch, ld, st));
            PHB_TRY(phase(make_init<2>(R, Y, sv, d, false)));
        } else if (s->pre == PHB200_PRE_JACOBI && method == PHB200_CG) {
            PHB_TRY(phase(make_init<1>(R, Y, q, d, false)));
        } else {
            // no preconditioner inside the initial direction (cg without M, cg_adaptive, torch paths, bicg: d = r)
            PHB_TRY(phase(make_init<0>(R, Y, q, bicg ? vec(0) : d, bicg)));
        }
        if (method == PHB200_CG_ADAPTIVE || method == PHB200_CG_ADAPTIVE_TORCH) {
            AdDots<T> e;   // q = A d, d.q, d.r -> alpha of the first pass
            fill_base(e, s, true, frozen());
            e.d = d; e.r = R; e.ld = ld; e.rules = rules(); e.do_check = 0; e.count_fev = 0;
            PHB_TRY(spmv(d, q, e));
        }
        if (bicg) {
            // vec(0) = shadow residual (r0 for L >= 2, ones for L == 1); all other work vectors start at zero
            const int nv = vectors_needed(method, s->L, s->pre);
            PHB_CUDA(cudaMemsetAsync(vec(1), 0, sizeof(T) * (size_t)(nv - 1) * batch * ld, st));
            if (s->L == 1) PHB_LAUNCH(k_fill<T>, kSMs * 4, kBlock, 0, st, vec(0), (int64_t)batch * ld, T(1));
        }
        return PHB200_OK;
    }

    // ------------------------------------------------------------------------------------------ fused stencil CG
    // Jacobi / un-preconditioned 'CG' on a star stencil: two kernels per iteration (SURVEY §8d counts 13 N w of traffic
    // for the three-phase form; this form moves 9 N w, +2 N w with a non-constant diagonal).
    bool fused() const {
        static int torch_on = -1;
        if (torch_on < 0) { const char* e = getenv("PHB200_CG_TORCH_FUSED"); torch_on = (e && e[0] == '0') ? 0 : 1; }
        const bool cg = s->method == PHB200_CG || (torch_on && s->method == PHB200_CG_TORCH && s->pre == PHB200_PRE_NONE && !s->totals);
        return cg && (star_stencil() || coded_star()) && !pre_applied(s->pre) && !cb_mode();
    }
    static bool coded_resid() {      // PHB200_CODED_RESID=0: residual half of the coded fused CG with the inverse-diagonal vector (A/B)
        static int v = -1;
        if (v < 0) { const char* e = getenv("PHB200_CODED_RESID"); v = (e && e[0] == '0') ? 0 : 1; }
        return v == 1;
    }
    bool star_stencil() const { return s->A->kind == PHB200_MAT_STENCIL && s->A->is_star; }
    // pattern-coded star with a per-row diagonal (ZERO_GRADIENT-type boundaries): the fused kernels in their DIAG form, TMA path only
    bool coded_star() const { return s->A->kind == PHB200_MAT_CODED && s->A->coded.star_diag && s->A->is_star && s->A->star.wrap == 0 && !s->coded_star_off && !s->totals; }
    StarDiag star_diag() const { return StarDiag{s->A->coded.code, s->A->coded.diag_tab, s->A->coded.inv_tab, s->A->coded.npat, nullptr, 0}; }
    // 'CG-adaptive' / 'auto' on a star stencil: residual kernel + TMA stencil kernel (direction, x update, product, d.q, d.r)
    bool fused_adaptive() const {
        static int on = -1;
        if (on < 0) { const char* e = getenv("PHB200_ADAPTIVE_FUSED"); on = (e && e[0] == '0') ? 0 : 1; }
        return on && (s->method == PHB200_CG_ADAPTIVE || s->method == PHB200_CG_ADAPTIVE_TORCH) && (star_stencil() || coded_star()) &&
               s->pre == PHB200_PRE_NONE && !cb_mode() && !s->totals;
    }
    CgDq<T> make_dq() {
        CgDq<T> e;
        fill_base(e, s, false, frozen());
        e.d = nullptr; e.ld = ld;
        e.defer = defer();
        return e;
    }
    template <int PRE>
    CgResid<T, PRE> make_resid(T* R) {
        CgResid<T, PRE> u;
        fill_base(u, s, false, frozen());
        u.v[0] = R; u.v[1] = vec(1); u.v[2] = PRE == 1 ? s->M->inv_diag : nullptr;
        u.cinv = PRE == 3 ? (T)s->M->const_inv : T(1); u.rules = rules();
        u.defer = defer();
        u.peer = PeerHalo{s->peer_lo, s->peer_hi, s->A->kind == PHB200_MAT_STENCIL ? s->A->star.ny * s->A->star.nz : 0, s->peer_ld_lo, s->peer_ld_hi};
        u.red.peer_data = (s->peer_lo || s->peer_hi) ? 1 : 0;      // halo planes go to peer memory: system-scope fences around the all-reduce
        return u;
    }
    // first half: d = M^-1 r + beta d, x += alpha d_old, q = A d, d.q
    template <int PRE>
    int fused_first(T* X, T* R) {
        T* q = vec(1);
        CgDirLoad<T, PRE> loader;
        loader.r = R; loader.buf_a = vec(0); loader.buf_b = vec(2); loader.xvec = X;
        loader.w = PRE == 1 ? (const T*)s->M->inv_diag : nullptr;
        loader.cinv = PRE == 3 ? (T)s->M->const_inv : T(1);
        loader.ld = ld; loader.glob = s->glob;
        CgDq<T> e = make_dq();
        prof_begin();
        int r;
        if (coded_star()) {
            // per-row diagonal: Jacobi (PRE == 1 here: the inverse diagonal is a vector for the residual kernel) takes the inverse from the pattern table
            if constexpr (PRE == 0) r = launch_tma_coded<0>(X, q, e);
            else r = launch_tma_coded<3>(X, q, e);
        } else if constexpr (PRE != 1) {
            if (s->tma_ok) r = launch_tma<PRE>(X, q, loader.cinv, e);
            else r = launch_star<T, CgDirLoad<T, PRE>, CgDq<T>, true>(s->A->star, loader, q, ld, batch, e, st);
        } else {
            r = launch_star<T, CgDirLoad<T, PRE>, CgDq<T>, true>(s->A->star, loader, q, ld, batch, e, st);
        }
        prof_end();
        return r;
    }
    // the residual half walks the chunks of the stencil kernel backwards (PhaseOrder): L2 reuse of q and r
    PhaseOrder resid_order() {
        PhaseOrder o{};
        static int mode = -1;
        if (mode < 0) { const char* e = getenv("PHB200_PHASE_ORDER"); mode = (e && e[0] == '1') ? 1 : 0; }   // off by default: measured neutral without L2 eviction hints
        o.keep_mask = l2_hints(1);
        o.stream_mask = l2_hints(2);
        if (!mode || !s->tma_ok) return o;
        dim3 grid;
        int xchunk;
        star_grid_balanced(s->A->star, batch, s->tma_occ > 0 ? s->tma_occ : 1, grid, xchunk);
        if (xchunk <= 0) return o;          // balanced decomposition: no common chunk structure to walk backwards
        const StarDesc& sd = s->A->star;
        o.plane = sd.ny * sd.nz;
        o.xchunk = xchunk;
        o.planes = sd.xe - sd.xb;
        o.chunks = (int)((o.planes + xchunk - 1) / xchunk);
        return o;
    }
    // PHB200_L2_HINTS=<stencil kernel bits, hex>,<residual kernel keep mask>,<residual kernel stream mask>  (see star_tma.cuh, phase.cuh)
    static unsigned l2_hints(int which) {
        // default: the residual kernel keeps r (evict_last: the next stencil pass reads it) and streams q (evict_first: dead after this read) —
        // measured -2.7 % per iteration at 256^3, neutral at 512^3 / 128^3; evict_first on d_old / x / d' in the stencil kernel gains 4 % at 256^3
        // but costs 6 % at 512^3 (halo rows of the neighbouring tiles leave L2 too early) — profiles/r02_l2_residency_experiments.jsonl
        static unsigned v[3] = {0, 1, 2};
        static bool parsed = false;
        if (!parsed) {
            parsed = true;
            if (const char* e = getenv("PHB200_L2_HINTS")) sscanf(e, "%x,%x,%x", &v[0], &v[1], &v[2]);
        }
        return v[which];
    }
    template <class P> int phase_ordered(const P& p, const PhaseOrder& ord) {
        prof_begin();
        int r = launch_phase<T, P>(p, s->n_act, ld, batch, st, s->off, ord, true);
        prof_end();
        return r;
    }
    template <int PRE>
    int pass_cg_fused_t(T* X, T* R) {
        // experimental (PHB200_L2_PERSIST=q|r): keep q (written by the first half, read by the second) or r (read by both halves,
        // rewritten by the second) in the persisting part of L2
        const int l2m = l2_persist_mode();
        static int which = -1;      // PHB200_L2_WHICH: 1 stencil kernel only, 2 residual kernel only, 3 both (default)
        if (which < 0) { const char* e = getenv("PHB200_L2_WHICH"); which = e ? atoi(e) : 3; }
        const L2Window win = l2m ? l2_persist_window(l2m == 2 ? (void*)R : (void*)vec(1), sizeof(T) * (size_t)batch * (size_t)ld) : L2Window{nullptr, 0, 0.f};
        if (which & 1) l2_window() = win;
        int r = fused_first<PRE>(X, R);
        l2_window() = (which & 2) ? win : L2Window{nullptr, 0, 0.f};
        if (r == PHB200_OK) {
            if (PRE == 1 && coded_star() && coded_resid()) {      // inverse diagonal from the pattern table (code byte per row) instead of the vector
                CgResidCoded<T> u;
                fill_base(u, s, false, frozen());
                u.v[0] = R; u.v[1] = vec(1); u.code = s->A->coded.code; u.inv_tab = (const T*)s->A->coded.inv_tab; u.rules = rules();
                u.defer = defer();
                PhaseOrder o = resid_order();
                o.keep_mask &= 0b11u; o.stream_mask &= 0b11u;
                r = phase_ordered(u, o);
            } else {
                r = phase_ordered(make_resid<PRE>(R), resid_order());
            }
        }
        l2_window() = L2Window{nullptr, 0, 0.f};
        return r;
    }
    int fused_pre() const { return s->pre == PHB200_PRE_NONE ? 0 : (s->M->is_const ? 3 : 1); }
    // host-driven halves for the multi-GPU slab mode: all-reduce of s->totals between the steps (multi_gpu.py)
    int dist_step(int step) {
        T* X = (T*)s->X;
        T* R = (T*)s->R;
        const int pre = fused_pre();
        switch (step) {
            case 0:   // first half, local sums -> totals
                s->prof_slot = 0;
                return pre == 0 ? fused_first<0>(X, R) : fused_first<3>(X, R);
            case 1:   // alpha from the reduced d.q, then the residual half (local r.s -> totals)
                if (!s->peer_reduce) PHB_TRY(finalize_only(make_dq()));
                return pre == 0 ? phase_ordered(make_resid<0>(R), resid_order()) : phase_ordered(make_resid<3>(R), resid_order());
            case 2:   // beta, delta, progress check from the reduced r.s
                s->passes += 1;
                if (s->peer_reduce) return PHB200_OK;     // done inside the residual kernel
                return pre == 0 ? finalize_only(make_resid<0>(R)) : finalize_only(make_resid<3>(R));
            case 10:  // scalar set-up after the all-reduce of the initial sums
                if (s->peer_reduce) return PHB200_OK;     // done inside the initial-residual kernel
                return pre == 0 ? finalize_only(make_init<0>(R, (const T*)s->Y, vec(1), vec(0), false)) : finalize_only(make_init<3>(R, (const T*)s->Y, vec(1), vec(0), false));
            case 20:
                return flush_x(X);
        }
        set_error("unknown distributed step %d", step);
        return PHB200_ERR_ARG;
    }
    template <int PRE>
    InitResidual<T, PRE == 3 ? 1 : PRE> make_init(T* R, const T* Y, const T* third, T* d, bool bicg) {
        InitResidual<T, PRE == 3 ? 1 : PRE> f;
        fill_base(f, s, true, frozen());
        f.v[0] = R; f.v[1] = Y; f.v[2] = third; f.v[3] = d; f.v[4] = (PRE == 1 || PRE == 3) ? s->M->inv_diag : nullptr;
        f.rtol = (const T*)s->rtol; f.atol = (const T*)s->atol; f.max_iter = s->max_iter;
        f.tol_from_y = s->method == PHB200_CG ? 0 : 1; f.rules = rules(); f.bicg = bicg ? 1 : 0;
        f.defer = defer();
        return f;
    }
    template <int PRE>
    int launch_tma(T* X, T* q, T cinv, const CgDq<T>& e) {
        const StarDesc& sd = s->A->star;
        const bool unit = sd.cxm == -1.0 && sd.cxp == -1.0 && sd.cym == -1.0 && sd.cyp == -1.0 && sd.czm == -1.0 && sd.czp == -1.0;
        return unit ? launch_tma_t<PRE, true>(X, q, cinv, e) : launch_tma_t<PRE, false>(X, q, cinv, e);
    }
    template <int PRE, bool UNIT>
    int launch_tma_t(T* X, T* q, T cinv, const CgDq<T>& e) {
        auto kernel = k_star_cg_tma<T, PRE, CgDq<T>, UNIT>;
        constexpr int smem = star_tma_smem<T>();
        static int occ_dev[kMaxDevices] = {};
        int& occ = occ_dev[device_slot()];
        if (occ == 0) {
            PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            int v = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kBlock, smem) != cudaSuccess || v < 1) { cudaGetLastError(); v = 1; }
            occ = v;
        }
        s->tma_occ = occ;
        dim3 grid;
        int xchunk;
        star_grid_balanced(s->A->star, batch, occ, grid, xchunk);
        PHB_LAUNCH_PDL(kernel, grid, dim3(kBlock), (size_t)smem, st, s->tm_r, s->tm_a, s->tm_b, s->A->star, X, vec(0), vec(2), q, cinv, ld, xchunk, 1, e, l2_hints(0), StarDiag{});
        return PHB200_OK;
    }
    template <int PRE>
    int launch_tma_coded(T* X, T* q, const CgDq<T>& e) {
        const StarDesc& sd = s->A->star;
        const bool unit = sd.cxm == -1.0 && sd.cxp == -1.0 && sd.cym == -1.0 && sd.cyp == -1.0 && sd.czm == -1.0 && sd.czp == -1.0;
        return unit ? launch_tma_coded_t<PRE, true>(X, q, e) : launch_tma_coded_t<PRE, false>(X, q, e);
    }
    template <int PRE, bool UNIT>
    int launch_tma_coded_t(T* X, T* q, const CgDq<T>& e) {
        auto kernel = k_star_cg_tma<T, PRE, CgDq<T>, UNIT, 0, true>;
        constexpr int smem = star_tma_smem<T>();
        static int occ_dev[kMaxDevices] = {};
        int& occ = occ_dev[device_slot()];
        if (occ == 0) {
            PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            int v = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kBlock, smem) != cudaSuccess || v < 1) { cudaGetLastError(); v = 1; }
            occ = v;
        }
        s->tma_occ = occ;
        dim3 grid;
        int xchunk;
        star_grid_balanced(s->A->star, batch, occ, grid, xchunk);
        PHB_LAUNCH_PDL(kernel, grid, dim3(kBlock), (size_t)smem, st, s->tm_r, s->tm_a, s->tm_b, s->A->star, X, vec(0), vec(2), q, T(1), ld, xchunk, 1, e, l2_hints(0), star_diag());
        return PHB200_OK;
    }
    int pass_cg_fused(T* X, T* R, const T* Y, bool refresh) {
        if (s->method == PHB200_CG_TORCH && refresh) {
            // torch_sparse_cg every 20th pass: x += alpha d, r = y - A x instead of r -= alpha q (_torch_backend.py:1371-1373)
            PHB_TRY(fused_first<0>(X, R));                  // d, pending x update, q = A d, alpha
            PHB_TRY(flush_x(X, 1));                         // x += alpha d NOW (and alpha = 0: nothing stays pending)
            T* t = vec(3);
            PlainEpi<T> pe;
            fill_base(pe, s, false, false);
            PHB_TRY(spmv(X, t, pe));
            RefreshResidual<T, false> rr;
            fill_base(rr, s, false, false);
            rr.v[0] = R; rr.v[1] = Y; rr.v[2] = t; rr.v[3] = vec(1); rr.rules = rules();
            return phase(rr);
        }
        if (s->pre == PHB200_PRE_NONE) return pass_cg_fused_t<0>(X, R);
        if (s->M->is_const) return pass_cg_fused_t<3>(X, R);
        return pass_cg_fused_t<1>(X, R);
    }
    // applies the pending x += alpha d after a batch of fused passes (x is then the true iterate)
    int flush_x(T* X, int flip = 0) {
        dim3 grid((unsigned)grid_x((s->n_act + kBlock * 4 - 1) / (kBlock * 4), batch, 8), (unsigned)batch);
        PHB_LAUNCH(k_flush_x<T>, grid, kBlock, 0, st, X + s->off, (const T*)vec(0) + s->off, (const T*)vec(2) + s->off, (const Sys*)s->sys, (const Glob*)s->glob, s->n_act, ld, flip);
        PHB_LAUNCH(k_zero_alpha, (unsigned)((batch + 255) / 256), 256, 0, st, s->sys, batch, (const Glob*)s->glob, flip);
        return PHB200_OK;
    }

    // ------------------------------------------------------------------------------------------ one pass
    int pass_cg(T* X, T* R, const T* Y, bool refresh) {
        if (fused()) return pass_cg_fused(X, R, Y, refresh);
        T* d = vec(0);
        T* q = vec(1);
        const bool fr = frozen();
        const bool torch = s->method == PHB200_CG_TORCH;
        {
            CgDq<T> e;
            fill_base(e, s, false, fr);
            e.d = d; e.ld = ld;
            PHB_TRY(spmv(d, q, e));
        }
        if (torch && refresh) {
            T* t = vec(2);
            AxpyX<T> ax;
            fill_base(ax, s, false, fr);
            ax.v[0] = X; ax.v[1] = d;
            PHB_TRY(phase(ax));
            PlainEpi<T> pe;
            fill_base(pe, s, false, fr);
            PHB_TRY(spmv(X, t, pe));
            RefreshResidual<T, false> rr;
            fill_base(rr, s, false, fr);
            rr.v[0] = R; rr.v[1] = Y; rr.v[2] = t; rr.v[3] = q; rr.rules = rules();
            PHB_TRY(phase(rr));
            CgDirection<T, 0> c;
            fill_base(c, s, false, fr);
            c.v[0] = d; c.v[1] = R; c.v[2] = nullptr;
            return phase(c);
        }
        if (s->pre == PHB200_PRE_JACOBI) {
            CgUpdate<T, 1> u;
            fill_base(u, s, false, fr);
            u.v[0] = X; u.v[1] = R; u.v[2] = d; u.v[3] = q; u.v[4] = s->M->inv_diag; u.rules = rules();
            PHB_TRY(phase(u));
            CgDirection<T, 1> c;
            fill_base(c, s, false, fr);
            c.v[0] = d; c.v[1] = R; c.v[2] = s->M->inv_diag;
            return phase(c);
        }
        if (pre_applied(s->pre)) {
            T* sv = vec(2);
            CgUpdate<T, 2> u;
            fill_base(u, s, false, fr);
            u.v[0] = X; u.v[1] = R; u.v[2] = d; u.v[3] = q; u.v[4] = nullptr; u.rules = rules();
            PHB_TRY(phase(u));
            // the triangular solves run unconditionally (they do not know about any_go); harmless after the end
            PHB_TRY(precond_apply(s->M, R, sv, vec(3), batch, ld, st));
            CgDeltaDot<T> dd;
            fill_base(dd, s, false, fr);
            dd.v[0] = R; dd.v[1] = sv; dd.rules = rules();
            PHB_TRY(phase(dd));
            CgDirection<T, 0> c;
            fill_base(c, s, false, fr);
            c.v[0] = d; c.v[1] = sv; c.v[2] = nullptr;
            return phase(c);
        }
        CgUpdate<T, 0> u;
        fill_base(u, s, false, fr);
        u.v[0] = X; u.v[1] = R; u.v[2] = d; u.v[3] = q; u.v[4] = nullptr; u.rules = rules();
        PHB_TRY(phase(u));
        CgDirection<T, 0> c;
        fill_base(c, s, false, fr);
        c.v[0] = d; c.v[1] = R; c.v[2] = nullptr;
        return phase(c);
    }

    template <bool UNIT, bool DIAG = false>
    int launch_tma_adaptive(T* X, T* q, int apply_x, const AdDots<T>& e) {
        auto kernel = k_star_cg_tma<T, 0, AdDots<T>, UNIT, 1, DIAG>;
        constexpr int smem = star_tma_smem<T>();
        static int occ_dev[kMaxDevices] = {};
        int& occ = occ_dev[device_slot()];
        if (occ == 0) {
            PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            int v = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kBlock, smem) != cudaSuccess || v < 1) { cudaGetLastError(); v = 1; }
            occ = v;
        }
        s->tma_occ = occ;
        dim3 grid;
        int xchunk;
        star_grid_balanced(s->A->star, batch, occ, grid, xchunk);
        PHB_LAUNCH_PDL(kernel, grid, dim3(kBlock), (size_t)smem, st, s->tm_r, s->tm_a, s->tm_b, s->A->star, X, vec(0), vec(2), q, T(1), ld, xchunk, apply_x, e, 0u,
                       DIAG ? star_diag() : StarDiag{});
        return PHB200_OK;
    }
    int pass_cg_adaptive_fused(T* X, T* R, const T* Y, bool refresh) {
        T* q = vec(1);
        const bool fr = frozen();
        T* dcur = (s->passes & 1) ? vec(2) : vec(0);      // direction of the previous pass (ping-pong parity = passes made)
        int apply_x = 1;
        if (s->method == PHB200_CG_ADAPTIVE_TORCH && refresh) {
            T* t = vec(3);
            AxpyX<T> ax;
            fill_base(ax, s, false, fr);
            ax.v[0] = X; ax.v[1] = dcur;
            PHB_TRY(phase(ax));
            PlainEpi<T> pe;
            fill_base(pe, s, false, fr);
            PHB_TRY(spmv(X, t, pe));
            RefreshResidual<T, true> rr;
            fill_base(rr, s, false, fr);
            rr.v[0] = R; rr.v[1] = Y; rr.v[2] = t; rr.v[3] = q; rr.rules = rules();
            PHB_TRY(phase(rr));
            apply_x = 0;                                  // x already carries alpha d
        } else {
            AdResid<T> u;
            fill_base(u, s, false, fr);
            u.v[0] = R; u.v[1] = q;
            PHB_TRY(phase(u));
        }
        AdDots<T> e;
        fill_base(e, s, false, fr);
        e.d = nullptr; e.r = nullptr; e.ld = ld; e.rules = rules(); e.do_check = 1; e.count_fev = 1;
        const StarDesc& sd = s->A->star;
        const bool unit = sd.cxm == -1.0 && sd.cxp == -1.0 && sd.cym == -1.0 && sd.cyp == -1.0 && sd.czm == -1.0 && sd.czp == -1.0;
        prof_begin();
        int r;
        if (coded_star()) r = unit ? launch_tma_adaptive<true, true>(X, q, apply_x, e) : launch_tma_adaptive<false, true>(X, q, apply_x, e);
        else r = unit ? launch_tma_adaptive<true>(X, q, apply_x, e) : launch_tma_adaptive<false>(X, q, apply_x, e);
        prof_end();
        return r;
    }

    int pass_cg_adaptive(T* X, T* R, const T* Y, bool refresh) {
        if (fused_adaptive() && s->tma_ok) return pass_cg_adaptive_fused(X, R, Y, refresh);
        T* d = vec(0);
        T* q = vec(1);
        const bool fr = frozen();
        if (s->method == PHB200_CG_ADAPTIVE_TORCH && refresh) {
            T* t = vec(3);
            AxpyX<T> ax;
            fill_base(ax, s, false, fr);
            ax.v[0] = X; ax.v[1] = d;
            PHB_TRY(phase(ax));
            PlainEpi<T> pe;
            fill_base(pe, s, false, fr);
            PHB_TRY(spmv(X, t, pe));
            RefreshResidual<T, true> rr;
            fill_base(rr, s, false, fr);
            rr.v[0] = R; rr.v[1] = Y; rr.v[2] = t; rr.v[3] = q; rr.rules = rules();
            PHB_TRY(phase(rr));
        } else {
            AdUpdate<T> u;
            fill_base(u, s, false, fr);
            u.v[0] = X; u.v[1] = R; u.v[2] = d; u.v[3] = q;
            PHB_TRY(phase(u));
        }
        AdDirection<T> c;
        fill_base(c, s, false, fr);
        c.v[0] = d; c.v[1] = R;
        PHB_TRY(phase(c));
        AdDots<T> e;
        fill_base(e, s, false, fr);
        e.d = d; e.r = R; e.ld = ld; e.rules = rules(); e.do_check = 1; e.count_fev = 1;
        return spmv(d, q, e);
    }

    // fused helpers of the BiCGstab(L) pass (fall back to single Lin2 launches beyond 3 pairs)
    template <int K, int MODE, bool WITH_X>
    int lin_multi(T* const* t, T* const* sv, int slot, int sign, T* X, const T* u, int xslot, int xsign) {
        LinMulti<T, K, MODE, WITH_X> f;
        fill_base(f, s, false, false);
        for (int i = 0; i < K; ++i) { f.v[2 * i] = t[i]; f.v[2 * i + 1] = sv[i]; }
        if (WITH_X) { f.v[2 * K] = X; f.v[2 * K + 1] = u; }
        f.slot = slot; f.sign = sign; f.xslot = xslot; f.xsign = xsign;
        return phase(f);
    }
    static bool bicg_fused_enabled() {
        static int v = -1;
        if (v < 0) { const char* e = getenv("PHB200_BICG_FUSED"); v = (e && e[0] == '0') ? 0 : 1; }
        return v == 1;
    }

    // BiCGstab(L), L >= 2 (phiml/backend/_linalg.py:158-221).  vec: 0 shadow, 1..L+1 uh[0..L], L+2..2L+1 rh[1..L], 2L+2 pv, 2L+3 tmp
    int pass_bicg_l(T* X, T* R) {
        const int L = s->L;
        const bool fuse = bicg_fused_enabled();
        const T* shadow = vec(0);
        std::vector<T*> uh(L + 1), rh(L + 1);
        for (int k = 0; k <= L; ++k) uh[k] = vec(1 + k);
        rh[0] = R;
        for (int k = 1; k <= L; ++k) rh[k] = vec(L + 1 + k);
        T* pv = vec(2 * L + 2);
        T* tmp = vec(2 * L + 3);
        const T* pin = nullptr;
        PHB_TRY(dot2(rh[0], shadow, rh[0], shadow, BI_RHO, 0, 0));
        for (int j = 0; j < L; ++j) {
            // uh[i] = rh[i] - beta uh[i], i <= j
            if (fuse && j == 0) PHB_TRY((lin_multi<1, 0, false>(&uh[0], &rh[0], S_BETA, -1, nullptr, nullptr, 0, 0)));
            else if (fuse && j == 1) PHB_TRY((lin_multi<2, 0, false>(&uh[0], &rh[0], S_BETA, -1, nullptr, nullptr, 0, 0)));
            else if (fuse && j == 2) PHB_TRY((lin_multi<3, 0, false>(&uh[0], &rh[0], S_BETA, -1, nullptr, nullptr, 0, 0)));
            else for (int i = 0; i <= j; ++i) PHB_TRY(lin2(uh[i], rh[i], uh[i], S_BETA, -1));
            PHB_TRY(apply_pre(uh[j], pv, tmp, &pin));
            PHB_TRY(spmv_dot(pin, uh[j + 1], shadow, nullptr, BI_GAMMA, j));
            // rh[i] = rh[i] - alpha uh[i+1], i <= j ; x += alpha uh[0] (independent of the product below: done in the same sweep)
            if (fuse && j == 0) PHB_TRY((lin_multi<1, 1, true>(&rh[0], &uh[1], S_ALPHA, -1, X, uh[0], S_ALPHA, +1)));
            else if (fuse && j == 1) PHB_TRY((lin_multi<2, 1, true>(&rh[0], &uh[1], S_ALPHA, -1, X, uh[0], S_ALPHA, +1)));
            else if (fuse && j == 2) PHB_TRY((lin_multi<3, 1, true>(&rh[0], &uh[1], S_ALPHA, -1, X, uh[0], S_ALPHA, +1)));
            else for (int i = 0; i <= j; ++i) PHB_TRY(lin2(rh[i], rh[i], uh[i + 1], S_ALPHA, -1));
            PHB_TRY(apply_pre(rh[j], pv, tmp, &pin));
            PHB_TRY(spmv_dot(pin, rh[j + 1], shadow, nullptr, BI_NEXT, j));
            if (!(fuse && j <= 2)) PHB_TRY(lin2(X, X, uh[0], S_ALPHA, +1));
        }
        for (int j = 1; j <= L; ++j) {
            for (int i = 1; i < j; ++i) {
                PHB_TRY(dot2(rh[j], rh[i], rh[j], rh[i], BI_TAU, j, i));
                if (fuse && i == j - 1) {   // last orthogonalisation step: fused with the sigma / gamma' dots
                    LinSigma<T> f;
                    fill_base(f, s, false, false);
                    f.v[0] = rh[j]; f.v[1] = rh[i]; f.v[2] = rh[0];
                    f.slot = S_TAU + j; f.sign = -1; f.j = j; f.L = L;
                    PHB_TRY(phase(f));
                } else {
                    PHB_TRY(lin2(rh[j], rh[j], rh[i], S_TAU + j, -1));
                }
            }
            if (!(fuse && j >= 2)) PHB_TRY(dot2(rh[j], rh[j], rh[0], rh[j], BI_SIGMA, j, 0));
        }
        if (fuse && L == 2) {
            BiFinal2<T> f;
            fill_base(f, s, false, false);
            f.v[0] = X; f.v[1] = rh[0]; f.v[2] = rh[1]; f.v[3] = rh[2]; f.v[4] = uh[0]; f.v[5] = uh[1]; f.v[6] = uh[2];
            f.rules = RULES_GENERIC;
            return phase(f);
        }
        PHB_TRY(lin2(X, X, rh[0], S_GAM + 1, +1));
        PHB_TRY(lin2(rh[0], rh[0], rh[L], S_GAMP + L, -1));
        PHB_TRY(lin2(uh[0], uh[0], uh[L], S_GAM + L, -1));
        for (int j = 1; j < L; ++j) {
            PHB_TRY(lin2(uh[0], uh[0], uh[j], S_GAM + j, -1));
            PHB_TRY(lin2(X, X, rh[j], S_GAMPP + j, +1));
            PHB_TRY(lin2(rh[0], rh[0], rh[j], S_GAMP + j, -1));
        }
        return dot2(R, R, R, R, BI_FINAL, 0, 0);
    }

    // BiCGstab(1) (phiml/backend/_linalg.py:250-271).  vec: 0 ones, 1 u, 2 p, 3 ph, 4 z, 5 t, 6 sl, 7 tl, 8 tmp, 9 tmp2
    int pass_bicg_1(T* X, T* R) {
        const T* ones = vec(0);
        T* u = vec(1);
        T* p = vec(2);
        T* ph = vec(3);
        T* z = vec(4);
        T* t = vec(5);
        T* sl = vec(6);
        T* tl = vec(7);
        T* tmp = vec(8);
        const T *php = nullptr, *zp = nullptr, *slp = nullptr, *tlp = nullptr;
        PHB_TRY(dot2(ones, R, ones, R, B1_RHO, 0, 0));
        PHB_TRY(lin2(p, p, u, S_OMEGA, -1));           // p - omega u
        PHB_TRY(lin2(p, R, p, S_BETA, +1));            // r + beta (.)
        PHB_TRY(apply_pre(p, ph, tmp, &php));
        PHB_TRY(spmv_dot(php, u, ones, nullptr, B1_ALPHA, 0));
        PHB_TRY(lin2(X, X, php, S_ALPHA, +1));         // h = x + alpha ph
        PHB_TRY(lin2(R, R, u, S_ALPHA, -1));           // s = r - alpha u   (kept in R)
        PHB_TRY(apply_inv_l(R, sl, &slp));
        PHB_TRY(apply_pre(R, z, tmp, &zp));
        PHB_TRY(spmv_dot(zp, t, ones, nullptr, B1_FEV, 0));
        PHB_TRY(apply_inv_l(t, tl, &tlp));
        PHB_TRY(dot2(tlp, slp, tlp, tlp, B1_OMEGA, 0, 0));
        PHB_TRY(lin2(X, X, zp, S_OMEGA, +1));          // x = h + omega z
        PHB_TRY(lin2(R, R, t, S_OMEGA, -1));           // r = s - omega t
        return dot2(R, R, R, R, BI_FINAL, 0, 0);
    }

    int pass(bool refresh) {
        T* X = (T*)s->X;
        T* R = (T*)s->R;
        const T* Y = (const T*)s->Y;
        switch (s->method) {
            case PHB200_CG:
            case PHB200_CG_TORCH: return pass_cg(X, R, Y, refresh);
            case PHB200_CG_ADAPTIVE:
            case PHB200_CG_ADAPTIVE_TORCH: return pass_cg_adaptive(X, R, Y, refresh);
            case PHB200_BICGSTAB: return s->L == 1 ? pass_bicg_1(X, R) : pass_bicg_l(X, R);
        }
        return PHB200_ERR_UNSUPPORTED;
    }
};

// ------------------------------------------------------------------------------------------------ resident (on-chip) CG
// Eligibility and launch geometry of k_cg_resident for this solver; false -> streaming kernels.
template <typename T>
static bool resident_plan(const phb200_solver* s, int& cs, int& lpc, size_t& smem) {
    int mode = s->resident;
    if (mode < 0) { const char* e = getenv("PHB200_RESIDENT"); if (e && (e[0] == '0' || e[0] == '1')) mode = e[0] - '0'; }
    if (mode == 0) return false;
    const bool adaptive = s->method == PHB200_CG_ADAPTIVE || s->method == PHB200_CG_ADAPTIVE_TORCH;
    if (!(s->method == PHB200_CG || s->method == PHB200_CG_TORCH || adaptive)) return false;
    if (adaptive && s->pre != PHB200_PRE_NONE) return false;
    if (s->A->kind != PHB200_MAT_STENCIL || !s->A->is_star || s->totals) return false;
    if (pre_applied(s->pre) || (s->pre == PHB200_PRE_JACOBI && !s->M->is_const)) return false;
    const StarDesc& st = s->A->star;
    if (st.xb != 0 || st.xe != st.nx || st.store_halo) return false;
    if (st.nz % 4 != 0 || st.nz > (1 << 20) || st.nx * st.ny > (1 << 20)) return false;
    const int64_t lines = st.nx * st.ny, qpl = st.nz / res_pack<T>(), cap_quads = (int64_t)kResThreads * kResQpt;
    for (int c = 1; c <= kResMaxCluster; c *= 2) {
        const int64_t l = (lines + c - 1) / c;
        const size_t bytes = res_smem_bytes<T>((int)l, (int)st.nz);
        if (l * qpl <= cap_quads && bytes <= 225 * 1024) {
            // cost model: one on-chip iteration of a cluster ~1 us; a streaming pass moves 9 vectors at ~5 TB/s, >= 25 us
            const int64_t ncl = kSMs / c;
            const double t_res = (double)((s->batch + ncl - 1) / ncl) * 1.0e-6;
            const double t_str = fmax(25e-6, 9.0 * (double)s->batch * (double)s->n * sizeof(T) / 5e12);
            if (t_res > t_str && mode != 1) return false;
            cs = c; lpc = (int)l; smem = bytes;
            return true;
        }
    }
    return false;
}

template <typename T, int PRE, bool RANK3, bool CLUSTER, int ALG = 0>
static int launch_resident_t(const ResArgs& a, const ResCoef<T>& cf, size_t smem, cudaStream_t st) {
    auto kernel = k_cg_resident<T, PRE, RANK3, CLUSTER, ALG>;
    static bool attr_dev[kMaxDevices] = {};
    bool& attr = attr_dev[device_slot()];
    if (!attr) {
        PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        attr = true;
    }
    const int ncl_max = kSMs / a.cs;
    const int ncl = a.batch < ncl_max ? a.batch : ncl_max;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ncl * a.cs));
    cfg.blockDim = dim3(kResThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)a.cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = CLUSTER ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a, cf);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e == cudaSuccess && debug_sync()) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("kernel k_cg_resident failed: %s", cudaGetErrorString(e));
        return PHB200_ERR_CUDA;
    }
    return PHB200_OK;
}

template <typename T>
static int launch_resident(const phb200_solver* s, const ResArgs& a, size_t smem, cudaStream_t st) {
    const bool rank3 = a.nx > 1, cl = a.cs > 1, jac = s->pre == PHB200_PRE_JACOBI;
    const StarDesc& sd = s->A->star;
    ResCoef<T> cf;
    cf.c0 = (T)sd.c0; cf.cxm = (T)sd.cxm; cf.cxp = (T)sd.cxp; cf.cym = (T)sd.cym; cf.cyp = (T)sd.cyp; cf.czm = (T)sd.czm; cf.czp = (T)sd.czp;
    cf.cinv = jac ? (T)s->M->const_inv : T(1);
    if (s->method == PHB200_CG_ADAPTIVE || s->method == PHB200_CG_ADAPTIVE_TORCH) {
        if (rank3) return cl ? launch_resident_t<T, 0, true, true, 1>(a, cf, smem, st) : launch_resident_t<T, 0, true, false, 1>(a, cf, smem, st);
        return cl ? launch_resident_t<T, 0, false, true, 1>(a, cf, smem, st) : launch_resident_t<T, 0, false, false, 1>(a, cf, smem, st);
    }
    if (jac) {
        if (rank3) return cl ? launch_resident_t<T, 3, true, true>(a, cf, smem, st) : launch_resident_t<T, 3, true, false>(a, cf, smem, st);
        return cl ? launch_resident_t<T, 3, false, true>(a, cf, smem, st) : launch_resident_t<T, 3, false, false>(a, cf, smem, st);
    }
    if (rank3) return cl ? launch_resident_t<T, 0, true, true>(a, cf, smem, st) : launch_resident_t<T, 0, true, false>(a, cf, smem, st);
    return cl ? launch_resident_t<T, 0, false, true>(a, cf, smem, st) : launch_resident_t<T, 0, false, false>(a, cf, smem, st);
}

// Whole solve on chip.  Generic rules: systems are independent, one launch.  torch rules: finished systems keep
// iterating (and are refreshed) until the LAST one is finished, so the state of every system is needed at the common
// final pass P = first pass at which nobody continues.  Round k runs every system to the first pass >= lb_k after which
// it stops; lb_{k+1} = max pass reached; done when all systems sit at the same pass (no pass before it can be the end,
// because the system that got there last was still iterating).  Usually two rounds.
template <typename T>
static int run_resident(phb200_solver* s, int64_t cap, int cs, int lpc, size_t smem, cudaStream_t st) {
    ResArgs a;
    a.nx = (int)s->A->star.nx; a.ny = (int)s->A->star.ny; a.nz = (int)s->A->star.nz;
    a.X = s->X; a.R = s->R; a.D = s->ws; a.Y = s->Y;
    a.Q = s->ws + (size_t)s->batch * (size_t)s->n * sizeof(T);      // work vector 1
    a.ld = s->n;
    a.batch = (int)s->batch;
    a.sys = s->sys; a.glob = s->glob;
    a.torch_rules = (s->method == PHB200_CG_TORCH || s->method == PHB200_CG_ADAPTIVE_TORCH) ? 1 : 0;
    a.lb = 0;
    a.cap = (int)(cap > 0x3fffffff ? 0x3fffffff : cap);
    a.lpc = lpc; a.cs = cs;
    int* d_out = nullptr;
    PHB_CUDA(phb_malloc_async(&d_out, 2 * sizeof(int), st));
    int r = PHB200_OK;
    for (int round = 0; round < 64; ++round) {
        r = launch_resident<T>(s, a, smem, st);
        if (r != PHB200_OK) break;
        k_res_round<<<1, kBlock, 0, st>>>(s->sys, (int)s->batch, a.cap, s->glob, d_out);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaMemcpyAsync(s->h_flag, d_out, 2 * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("resident CG failed: %s", cudaGetErrorString(cudaGetLastError()));
            r = PHB200_ERR_CUDA;
            break;
        }
        s->passes = s->h_flag[0];
        if (!a.torch_rules || s->h_flag[1]) break;
        a.lb = s->h_flag[0];
    }
    cudaFreeAsync(d_out, st);
    return r;
}

template <typename T>
static Run<T> make_run(phb200_solver* s, cudaStream_t st) {
    Run<T> r;
    r.s = s;
    r.st = st;
    r.n = s->n;
    r.ld = s->n;
    r.batch = (int)s->batch;
    return r;
}

static int advance(phb200_solver* s, int n_iter, cudaStream_t st) {
    for (int k = 0; k < n_iter; ++k) {
        const int64_t pass_no = s->passes + 1;
        const bool torch = s->method == PHB200_CG_TORCH || s->method == PHB200_CG_ADAPTIVE_TORCH;
        const bool refresh = torch && (pass_no % 20 == 0);
        s->prof_slot = 0;
        int r;
        if (s->dtype == PHB200_F32) r = make_run<float>(s, st).pass(refresh);
        else r = make_run<double>(s, st).pass(refresh);
        if (r != PHB200_OK) return r;
        s->passes = pass_no;
    }
    if (n_iter > 0) {
        if (s->dtype == PHB200_F32) { auto run = make_run<float>(s, st); if (run.fused()) PHB_TRY(run.flush_x((float*)s->X)); }
        else { auto run = make_run<double>(s, st); if (run.fused()) PHB_TRY(run.flush_x((double*)s->X)); }
    }
    return PHB200_OK;
}

}  // namespace phb

// Solver scratch (per-system scalars, reduction partials, tickets, the pinned poll flag and its events) is recycled between solver
// objects: a create / destroy pair per solve_linear call cost ~8 ms of cudaMalloc / cudaMallocHost / cudaFree on a process that
// holds GBs of torch allocations — 10 % of a 256^3 solve.  At most kPoolMax blocks are kept; a block is reused for requests
// between a quarter of and its full size.
struct ScratchBlock {
    void* dev;
    size_t bytes;
    int* h_flag;
    cudaEvent_t ev[2];
    int device;
};
constexpr size_t kPoolMax = 4;
static std::mutex g_pool_mu;
static std::vector<ScratchBlock> g_scratch_pool;

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static cudaError_t scratch_acquire(size_t bytes, int device, ScratchBlock* out) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        for (size_t i = 0; i < g_scratch_pool.size(); ++i) {
            const ScratchBlock& b = g_scratch_pool[i];
            if (b.device == device && b.bytes >= bytes && b.bytes / 4 <= bytes) {
                *out = b;
                g_scratch_pool.erase(g_scratch_pool.begin() + (long)i);
                return cudaSuccess;
            }
        }
    }
    ScratchBlock b{};
    b.bytes = bytes;
    b.device = device;
    cudaError_t e = cudaMalloc(&b.dev, bytes);
    if (e == cudaSuccess) e = cudaMallocHost(&b.h_flag, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev[1], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        if (b.dev) cudaFree(b.dev);
        if (b.h_flag) cudaFreeHost(b.h_flag);
        if (b.ev[0]) cudaEventDestroy(b.ev[0]);
        if (b.ev[1]) cudaEventDestroy(b.ev[1]);
        return e;
    }
    *out = b;
    return cudaSuccess;
}

static void scratch_release(const ScratchBlock& b) {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur == b.device) {
        // same guarantee cudaFree gave: nothing queued by the old owner still uses the block when the next owner gets it
        if (cudaDeviceSynchronize() == cudaSuccess) {
            std::lock_guard<std::mutex> lock(g_pool_mu);
            if (g_scratch_pool.size() < kPoolMax) {
                g_scratch_pool.push_back(b);
                return;
            }
        }
    }
    cudaFree(b.dev);
    cudaFreeHost(b.h_flag);
    cudaEventDestroy(b.ev[0]);
    cudaEventDestroy(b.ev[1]);
}

extern "C" {

int phb200_solver_create(phb200_solver** out, const phb200_matrix* A, const phb200_precond* M, int method, int poly_order, int64_t batch) {
    PHB_ARG(out && A, "null argument");
    PHB_ARG(A->n_rows == A->n_cols, "matrix must be square");
    PHB_ARG(batch >= 1 && batch <= 65535, "batch must be 1..65535, got %lld", (long long)batch);
    PHB_ARG(method >= PHB200_CG && method <= PHB200_BICGSTAB, "unknown method %d", method);
    PHB_ARG(!(A->kind == PHB200_MAT_CSR && A->transposed), "solver needs a plain CSR or stencil matrix (convert CSC natives first)");
    if (method == PHB200_BICGSTAB) {
        if (poly_order == 0) { set_error("Generic non-stabilized biCG not supported. Use 'scipy-biCG' instead"); return PHB200_ERR_UNSUPPORTED; }
        PHB_ARG(poly_order >= 1 && poly_order <= 4, "biCG-stab(L) supports 1 <= L <= 4, got %d", poly_order);
    } else {
        poly_order = 0;
    }
    const int pre = M ? M->kind : PHB200_PRE_NONE;
    if (M) PHB_ARG(M->n == A->n_rows && M->dtype == A->dtype, "preconditioner does not match the matrix");
    if (method != PHB200_CG && method != PHB200_BICGSTAB && pre != PHB200_PRE_NONE) {
        set_error("method %d takes no preconditioner (the reference falls back to 'CG')", method);
        return PHB200_ERR_UNSUPPORTED;
    }
    phb200_solver* s = new phb200_solver();
    memset(s, 0, sizeof(*s));
    s->A = A;
    s->M = M;
    s->method = method;
    s->L = poly_order;
    s->dtype = A->dtype;
    s->pre = pre;
    s->batch = batch;
    s->n = A->n_rows;
    s->off = 0;
    s->n_act = A->n_rows;
    s->resident = -1;
    if (A->kind == PHB200_MAT_STENCIL && A->is_star) {
        s->off = A->star.xb * A->star.ny * A->star.nz;
        s->n_act = (A->star.xe - A->star.xb) * A->star.ny * A->star.nz;
    }
    s->prof_events = new std::vector<cudaEvent_t>();
    s->prof_slots = new std::vector<int>();
    const size_t b_sys = align256(sizeof(Sys) * (size_t)batch), b_glob = align256(sizeof(Glob));
    const size_t b_part = align256(sizeof(double) * kMaxDots * (size_t)max_blocks_for(batch) * (size_t)batch), b_tick = align256(sizeof(unsigned) * (size_t)batch);
    ScratchBlock blk{};
    cudaError_t e = cudaGetDevice(&s->device);
    if (e == cudaSuccess) e = scratch_acquire(b_sys + b_glob + b_part + b_tick, s->device, &blk);
    if (e == cudaSuccess) {
        char* base = (char*)blk.dev;
        s->scratch = blk.dev;
        s->scratch_bytes = blk.bytes;
        s->sys = (Sys*)base;
        s->glob = (Glob*)(base + b_sys);
        s->partials = (double*)(base + b_sys + b_glob);
        s->tickets = (unsigned*)(base + b_sys + b_glob + b_part);
        s->h_flag = blk.h_flag;
        s->ev[0] = blk.ev[0];
        s->ev[1] = blk.ev[1];
    }
    if (e != cudaSuccess) {
        set_error("solver allocation failed: %s", cudaGetErrorString(e));
        phb200_solver_destroy(s);
        return PHB200_ERR_CUDA;
    }
    *out = s;
    return PHB200_OK;
}

size_t phb200_solver_workspace_bytes(const phb200_solver* s) {
    if (!s) return 0;
    const size_t w = s->dtype == PHB200_F32 ? 4 : 8;
    return (size_t)vectors_needed(s->method, s->L, s->pre) * (size_t)s->batch * (size_t)s->n * w + 256;
}

int phb200_solver_start(phb200_solver* s, const void* Y, const void* X0, void* X, void* R, const void* rtol, const void* atol,
                        const int32_t* max_iter, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    PHB_ARG(s && Y && X0 && X && R && rtol && atol && max_iter, "null argument");
    if (workspace_bytes < phb200_solver_workspace_bytes(s) || (!workspace && workspace_bytes > 0)) {
        set_error("workspace too small: need %zu bytes, got %zu", phb200_solver_workspace_bytes(s), workspace_bytes);
        return PHB200_ERR_WORKSPACE;
    }
    PHB_ARG(((uintptr_t)workspace) % 256 == 0, "workspace must be 256-byte aligned");
    s->Y = Y;
    s->X = X;
    s->R = R;
    s->ws = (char*)workspace;
    s->ws_bytes = workspace_bytes;
    s->passes = 0;
    s->launches0 = g_launches.load();
    int r;
    if (s->dtype == PHB200_F32) r = make_run<float>(s, st).start((const float*)Y, (const float*)X0, (float*)X, (float*)R, (const float*)rtol, (const float*)atol, max_iter);
    else r = make_run<double>(s, st).start((const double*)Y, (const double*)X0, (double*)X, (double*)R, (const double*)rtol, (const double*)atol, max_iter);
    s->started = r == PHB200_OK;
    return r;
}

int phb200_solver_tol_min(phb200_solver* s, void* tol_min, int set, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    PHB_ARG(s && s->started && tol_min, "solver not started");
    if (!set) {
        static_assert(offsetof(Glob, yy_sum) == offsetof(Glob, tol_min) + sizeof(double), "tol_min and yy_sum are read as one pair");
        PHB_CUDA(cudaMemcpyAsync(tol_min, &s->glob->tol_min, 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return PHB200_OK;
    }
    PHB_ARG(set == 1 || set == 2, "set must be 0 (read), 1 (install the minimum tolerance) or 2 (install the batch-wide sum of |y|^2)");
    PHB_ARG(s->passes == 0, "the tolerance can only be replaced before the first pass");
    FBase f;
    fill_base(f, s, true, s->method == PHB200_CG || s->method == PHB200_CG_ADAPTIVE);
    const int rules = (s->method == PHB200_CG_TORCH || s->method == PHB200_CG_ADAPTIVE_TORCH) ? RULES_TORCH : RULES_GENERIC;
    const int adaptive = (s->method == PHB200_CG_ADAPTIVE || s->method == PHB200_CG_ADAPTIVE_TORCH) ? 1 : 0;
    const int sum_rule = (s->method != PHB200_CG && rules == RULES_GENERIC) ? 1 : 0;
    if (s->dtype == PHB200_F32) PHB_LAUNCH(k_recheck<float>, 1, kBlock, 0, st, f, rules, (const double*)tol_min, adaptive, set, (const float*)s->rtol, (const float*)s->atol, sum_rule);
    else PHB_LAUNCH(k_recheck<double>, 1, kBlock, 0, st, f, rules, (const double*)tol_min, adaptive, set, (const double*)s->rtol, (const double*)s->atol, sum_rule);
    return PHB200_OK;
}

int phb200_solver_set_dist(phb200_solver* s, void* totals) {
    PHB_ARG(s && !s->started, "set_dist must be called before start");
    PHB_ARG(s->method == PHB200_CG && s->A->kind == PHB200_MAT_STENCIL && s->A->is_star && !pre_applied(s->pre) && (s->pre == PHB200_PRE_NONE || s->M->is_const),
            "distributed mode supports 'CG' on a star stencil without or with constant-diagonal Jacobi preconditioner");
    s->totals = (double*)totals;
    return PHB200_OK;
}

int phb200_solver_set_dist_callbacks(phb200_solver* s, void* totals, phb200_allreduce_fn allreduce, phb200_halo_fn halo, void* user) {
    PHB_ARG(s && !s->started, "set_dist_callbacks must be called before start");
    PHB_ARG(totals && allreduce && halo, "null argument");
    PHB_ARG(s->A->kind == PHB200_MAT_STENCIL && s->A->is_star, "distributed solves need a star stencil in slab mode (phb200_matrix_set_slab)");
    PHB_ARG(!pre_applied(s->pre), "ILU across slabs is a pipeline along x (and 'cluster' a global coarse solve): not supported in distributed mode");
    s->totals = (double*)totals;
    s->ar_cb = allreduce;
    s->halo_cb = halo;
    s->cb_user = user;
    return PHB200_OK;
}

int phb200_solver_set_peer_native(phb200_solver* s, void* totals, void* const* mailboxes, int world, int rank,
                                  const phb200_peer_slab* lower, const phb200_peer_slab* upper, void* my_flags, void* stream_) {
    PHB_ARG(s && !s->started && !s->totals, "set_peer_native must be called before start, instead of the other distributed modes");
    PHB_ARG(totals && my_flags, "null argument");
    PHB_ARG(s->A->kind == PHB200_MAT_STENCIL && s->A->is_star, "distributed solves need a star stencil in slab mode (phb200_matrix_set_slab)");
    PHB_ARG(!pre_applied(s->pre), "ILU across slabs is a pipeline along x (and 'cluster' a global coarse solve): not supported in distributed mode");
    s->totals = (double*)totals;          // marks the solver as distributed; the native mode defers nothing into it
    int r = phb200_solver_set_peer_reduce(s, mailboxes, world, rank, stream_);
    if (r != PHB200_OK) { s->totals = nullptr; return r; }
    memset(&s->nb_lo, 0, sizeof(s->nb_lo));
    memset(&s->nb_hi, 0, sizeof(s->nb_hi));
    if (lower && lower->X) s->nb_lo = *lower;
    if (upper && upper->X) s->nb_hi = *upper;
    s->halo_flags = (unsigned long long*)my_flags;
    s->halo_seq = 0;
    if (!s->halo_ticket) {
        PHB_CUDA(cudaMalloc(&s->halo_ticket, sizeof(unsigned)));
        PHB_CUDA(cudaMemsetAsync(s->halo_ticket, 0, sizeof(unsigned), (cudaStream_t)stream_));
    }
    s->native_peer = 1;
    return PHB200_OK;
}

int phb200_solver_set_peer_halos(phb200_solver* s, void* lower_neighbour_halo, int64_t ld_lower, void* upper_neighbour_halo, int64_t ld_upper) {
    PHB_ARG(s && s->totals && !s->ar_cb, "peer halos belong to the step-driven distributed mode (phb200_solver_set_dist)");
    s->peer_lo = lower_neighbour_halo;
    s->peer_hi = upper_neighbour_halo;
    s->peer_ld_lo = ld_lower;
    s->peer_ld_hi = ld_upper;
    return PHB200_OK;
}

int phb200_solver_set_peer_reduce(phb200_solver* s, void* const* mailboxes, int world, int rank, void* stream_) {
    PHB_ARG(s && s->totals && !s->ar_cb && !s->started, "peer reduce belongs to the step-driven / native distributed modes and is set before start");
    PHB_ARG(mailboxes && world >= 1 && world <= 8 && rank >= 0 && rank < world, "bad world / rank (at most 8 ranks)");
    cudaStream_t st = (cudaStream_t)stream_;
    PeerReduce h;
    memset(&h, 0, sizeof(h));
    for (int p = 0; p < world; ++p) {
        PHB_ARG(mailboxes[p], "null mailbox");
        h.mail[p] = (PeerMail*)mailboxes[p];
    }
    h.world = world; h.rank = rank; h.batch = (int)s->batch;
    if (!s->peer_seq) PHB_CUDA(cudaMalloc(&s->peer_seq, sizeof(unsigned long long) * s->batch));
    PHB_CUDA(cudaMemsetAsync(s->peer_seq, 0, sizeof(unsigned long long) * s->batch, st));
    h.seq = s->peer_seq;
    if (!s->peer_reduce) PHB_CUDA(cudaMalloc(&s->peer_reduce, sizeof(PeerReduce)));
    PHB_CUDA(cudaMemcpyAsync(s->peer_reduce, &h, sizeof(PeerReduce), cudaMemcpyHostToDevice, st));
    PHB_CUDA(cudaStreamSynchronize(st));      // `h` is a stack object
    return PHB200_OK;
}

size_t phb200_peer_mailbox_bytes(int world, int64_t batch) { return sizeof(PeerMail) * 2 * (size_t)world * (size_t)batch; }

int phb200_solver_dist_step(phb200_solver* s, int step, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    PHB_ARG(s && s->started && s->totals && !s->ar_cb, "solver not started in (step-driven) distributed mode");
    if (!s->tma_ok) { set_error("distributed mode needs the TMA path (tensor map creation failed)"); return PHB200_ERR_UNSUPPORTED; }
    if (s->dtype == PHB200_F32) return make_run<float>(s, st).dist_step(step);
    return make_run<double>(s, st).dist_step(step);
}

int phb200_solver_dist_advance(phb200_solver* s, int n_iter, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    PHB_ARG(s && s->started && s->totals && !s->ar_cb, "solver not started in (step-driven) distributed mode");
    PHB_ARG(s->peer_reduce && (s->peer_lo || s->peer_hi), "dist_advance needs the in-kernel all-reduce and the peer-memory halo stores (set_peer_reduce, set_peer_halos)");
    PHB_ARG(n_iter >= 0, "n_iter must be >= 0");
    if (!s->tma_ok) { set_error("distributed mode needs the TMA path (tensor map creation failed)"); return PHB200_ERR_UNSUPPORTED; }
    for (int k = 0; k < n_iter; ++k) {
        for (int step = 0; step < 3; ++step) {
            const int r = s->dtype == PHB200_F32 ? make_run<float>(s, st).dist_step(step) : make_run<double>(s, st).dist_step(step);
            if (r != PHB200_OK) return r;
        }
    }
    return PHB200_OK;
}

int phb200_solver_advance(phb200_solver* s, int n_iter, void* stream_) {
    PHB_ARG(s && s->started, "solver not started");
    PHB_ARG(!s->totals || s->ar_cb || s->native_peer, "a step-driven distributed solver is advanced with phb200_solver_dist_step");
    PHB_ARG(n_iter >= 0, "n_iter must be >= 0");
    return advance(s, n_iter, (cudaStream_t)stream_);
}

int phb200_solver_option(phb200_solver* s, int option, int value, int* previous) {
    PHB_ARG(s, "null solver");
    if (option == PHB200_OPT_RESIDENT) {
        if (previous) *previous = s->resident;
        if (value >= -1 && value <= 1) s->resident = value;
        return PHB200_OK;
    }
    if (option == PHB200_OPT_RESIDENT_USED) {
        if (previous) *previous = s->resident_used;
        return PHB200_OK;
    }
    set_error("unknown solver option %d", option);
    return PHB200_ERR_ARG;
}

int phb200_solver_poll(phb200_solver* s, int* any_continue, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    PHB_ARG(s && s->started && any_continue, "solver not started");
    PHB_CUDA(cudaMemcpyAsync(&s->h_flag[0], &s->glob->any_go, sizeof(int), cudaMemcpyDeviceToHost, st));
    PHB_CUDA(cudaStreamSynchronize(st));
    *any_continue = s->h_flag[0];
    return PHB200_OK;
}

int phb200_solver_run(phb200_solver* s, int64_t iteration_cap, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    PHB_ARG(s && s->started, "solver not started");
    PHB_ARG(!s->totals || s->ar_cb || s->native_peer, "a step-driven distributed solver is advanced with phb200_solver_dist_step");
    if (s->passes == 0) {   // small systems: the whole loop on chip (resident.cuh)
        int cs = 0, lpc = 0;
        size_t smem = 0;
        s->resident_used = 0;
        if (s->dtype == PHB200_F32 && resident_plan<float>(s, cs, lpc, smem)) { s->resident_used = 1; return run_resident<float>(s, iteration_cap, cs, lpc, smem, st); }
        if (s->dtype == PHB200_F64 && resident_plan<double>(s, cs, lpc, smem)) { s->resident_used = 1; return run_resident<double>(s, iteration_cap, cs, lpc, smem, st); }
    }
    // The flag of chunk k is inspected while chunk k+1 is already queued, so the device never waits for the host.
    int chunk = 4;
    int slot = 0;
    bool have_prev = false;
    int64_t done = 0;
    while (done < iteration_cap) {
        const int n = (int)((iteration_cap - done) < chunk ? (iteration_cap - done) : chunk);
        PHB_TRY(advance(s, n, st));
        done += n;
        PHB_CUDA(cudaMemcpyAsync(&s->h_flag[slot], &s->glob->any_go, sizeof(int), cudaMemcpyDeviceToHost, st));
        PHB_CUDA(cudaEventRecord(s->ev[slot], st));
        if (have_prev) {
            PHB_CUDA(cudaEventSynchronize(s->ev[slot ^ 1]));
            if (s->h_flag[slot ^ 1] == 0) break;
        }
        have_prev = true;
        slot ^= 1;
        if (chunk < 32) chunk *= 2;
    }
    PHB_CUDA(cudaStreamSynchronize(st));
    return PHB200_OK;
}

int phb200_solver_result(phb200_solver* s, int32_t* iterations, int32_t* function_evaluations, uint8_t* converged, uint8_t* diverged, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    PHB_ARG(s && s->started, "solver not started");
    PHB_LAUNCH(k_collect, (unsigned)((s->batch + 255) / 256), 256, 0, st, s->sys, (int)s->batch, iterations, function_evaluations, converged, diverged);
    if ((s->method == PHB200_CG || s->method == PHB200_CG_ADAPTIVE) && s->batch > 1 && s->X && s->R && s->n_act > 0) {
        // stopped systems with a non-finite residual norm: the NaNs the reference's masked updates produce while the rest of the batch iterates
        // (almost always nothing to do: one CTA per system and 16 K rows keeps the check itself cheap for batches of thousands of small systems)
        const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((s->n_act + 16383) / 16384, 64));
        dim3 grid(gx, (unsigned)s->batch);
        if (s->dtype == PHB200_F32) PHB_LAUNCH(k_poison_nonfinite<float>, grid, 256, 0, st, s->sys, (int)s->batch, (float*)s->X, (float*)s->R, s->n, s->off, s->n_act);
        else PHB_LAUNCH(k_poison_nonfinite<double>, grid, 256, 0, st, s->sys, (int)s->batch, (double*)s->X, (double*)s->R, s->n, s->off, s->n_act);
    }
    return PHB200_OK;
}

int phb200_solver_stats(const phb200_solver* s, int64_t* passes_enqueued, int64_t* launches) {
    PHB_ARG(s, "null solver");
    if (passes_enqueued) *passes_enqueued = s->passes;
    if (launches) *launches = (int64_t)(g_launches.load() - s->launches0);
    return PHB200_OK;
}

int phb200_solver_profile(phb200_solver* s, int enable) {
    PHB_ARG(s, "null solver");
    s->prof = enable ? 1 : 0;
    return PHB200_OK;
}

int phb200_solver_profile_read(phb200_solver* s, double* ms_per_slot, int64_t* count_per_slot, int max_slots, int* n_slots) {
    PHB_ARG(s && ms_per_slot && count_per_slot && n_slots && max_slots > 0, "null argument");
    for (int k = 0; k < max_slots; ++k) { ms_per_slot[k] = 0.0; count_per_slot[k] = 0; }
    int used = 0;
    for (size_t k = 0; k < s->prof_slots->size(); ++k) {
        cudaEvent_t a = (*s->prof_events)[2 * k], b = (*s->prof_events)[2 * k + 1];
        PHB_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        PHB_CUDA(cudaEventElapsedTime(&ms, a, b));
        const int slot = (*s->prof_slots)[k];
        if (slot < max_slots) {
            ms_per_slot[slot] += ms;
            count_per_slot[slot] += 1;
            if (slot + 1 > used) used = slot + 1;
        }
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    s->prof_events->clear();
    s->prof_slots->clear();
    *n_slots = used;
    return PHB200_OK;
}

int phb200_solver_destroy(phb200_solver* s) {
    if (!s) return PHB200_OK;
    if (s->prof_events) {
        for (cudaEvent_t e : *s->prof_events) cudaEventDestroy(e);
        delete s->prof_events;
        delete s->prof_slots;
    }
    if (s->peer_reduce) cudaFree(s->peer_reduce);
    if (s->halo_ticket) cudaFree(s->halo_ticket);
    if (s->peer_seq) cudaFree(s->peer_seq);
    if (s->scratch) {
        ScratchBlock blk{s->scratch, s->scratch_bytes, s->h_flag, {s->ev[0], s->ev[1]}, s->device};
        scratch_release(blk);
    }
    delete s;
    return PHB200_OK;
}

}  // extern "C"
