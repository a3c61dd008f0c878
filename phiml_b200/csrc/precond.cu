// Preconditioners: Jacobi set-up, exact ILU(0) with a level schedule, level-scheduled sparse triangular solves.
#include "phase.cuh"
#include "precond.cuh"
#include "star_trsv.cuh"
#include "star_trsv_pipe.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <vector>
#include <stdlib.h>

namespace phb {

// ---------------------------------------------------------------------------------------------- Jacobi
template <typename T>
__global__ void __launch_bounds__(kBlock) k_jacobi_csr(CsrView A, T* __restrict__ inv_diag, int* missing) {
    const T* __restrict__ val = (const T*)A.val;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < A.n_rows; i += (int64_t)gridDim.x * kBlock) {
        T d = T(0);
        bool found = false;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
            if (A.col[k] == (int32_t)i) { d = d + val[k]; found = true; }
        if (!found) *missing = 1;
        inv_diag[i] = T(1) / d;
    }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_jacobi_const(T* __restrict__ inv_diag, int64_t n, T value) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) inv_diag[i] = value;
}

// ---------------------------------------------------------------------------------------------- level schedule
// level(i) = 1 + max level of the rows i depends on (columns < i for a lower, > i for an upper solve).  Rows of one
// level are independent.  The analysis runs once per matrix on the host (O(nnz)), the schedule lives on the device.
__global__ void __launch_bounds__(kBlock) k_pad_order(const int32_t* __restrict__ order, int64_t n, const int32_t* __restrict__ level_ptr, const int64_t* __restrict__ pad_ptr,
                                                      int32_t n_levels, int32_t* __restrict__ order_pad) {
    for (int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x; k < n; k += (int64_t)gridDim.x * kBlock) {
        int lo = 0, hi = n_levels;   // level l with level_ptr[l] <= k < level_ptr[l+1]
        while (lo + 1 < hi) {
            const int mid = (lo + hi) >> 1;
            if (level_ptr[mid] <= k) lo = mid;
            else hi = mid;
        }
        order_pad[pad_ptr[lo] + (k - level_ptr[lo])] = order[k];
    }
}

// p->order holds the rows sorted by level; level_ptr (host) the level boundaries.  Builds the device tables.
static int finish_plan(phb200_trsv_plan* p, const std::vector<int32_t>& level_ptr, cudaStream_t stream) {
    const int64_t n = p->n;
    const int32_t n_levels = (int32_t)level_ptr.size() - 1;
    p->n_levels = n_levels;
    std::vector<int64_t> pad_ptr(n_levels + 1, 0);
    int32_t max_rows = 0;
    for (int32_t l = 0; l < n_levels; ++l) {
        const int32_t cnt = level_ptr[l + 1] - level_ptr[l];
        if (cnt > max_rows) max_rows = cnt;
        pad_ptr[l + 1] = pad_ptr[l] + ((cnt + 31) / 32) * 32;
    }
    p->max_level_rows = max_rows;
    p->npad = pad_ptr[n_levels];
    p->epoch = 0;
    p->h_level_ptr = (int32_t*)malloc(sizeof(int32_t) * (n_levels + 1));
    memcpy(p->h_level_ptr, level_ptr.data(), sizeof(int32_t) * (n_levels + 1));
    int64_t* d_pad_ptr = nullptr;
    cudaError_t e = cudaMalloc(&p->level_ptr, sizeof(int32_t) * (n_levels + 1));
    if (e == cudaSuccess) e = cudaMalloc(&p->order_pad, sizeof(int32_t) * (p->npad > 0 ? p->npad : 1));
    const int64_t nwarps = p->npad / 32;
    std::vector<int32_t> warp_level((size_t)(nwarps > 0 ? nwarps : 1), 0), level_warps((size_t)(n_levels > 0 ? n_levels : 1), 0);
    for (int32_t l = 0; l < n_levels; ++l) {
        level_warps[l] = (int32_t)((pad_ptr[l + 1] - pad_ptr[l]) / 32);
        for (int64_t w = pad_ptr[l] / 32; w < pad_ptr[l + 1] / 32; ++w) warp_level[w] = l;
    }
    if (e == cudaSuccess) e = cudaMalloc(&p->warp_level, sizeof(int32_t) * warp_level.size());
    if (e == cudaSuccess) e = cudaMalloc(&p->level_warps, sizeof(int32_t) * level_warps.size());
    if (e == cudaSuccess) e = cudaMalloc(&p->level_cnt, sizeof(unsigned long long) * level_warps.size());
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->warp_level, warp_level.data(), sizeof(int32_t) * warp_level.size(), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->level_warps, level_warps.data(), sizeof(int32_t) * level_warps.size(), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->level_cnt, 0, sizeof(unsigned long long) * level_warps.size(), stream);
    if (e == cudaSuccess) e = cudaMalloc(&d_pad_ptr, sizeof(int64_t) * (n_levels + 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->level_ptr, level_ptr.data(), sizeof(int32_t) * (n_levels + 1), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_pad_ptr, pad_ptr.data(), sizeof(int64_t) * (n_levels + 1), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->order_pad, 0xFF, sizeof(int32_t) * (p->npad > 0 ? p->npad : 1), stream);
    if (e == cudaSuccess && n > 0) {
        k_pad_order<<<(unsigned)grid_x((n + kBlock - 1) / kBlock, 1), kBlock, 0, stream>>>(p->order, n, p->level_ptr, d_pad_ptr, n_levels, p->order_pad);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);   // pad_ptr (host vector) must outlive the copy
    if (d_pad_ptr) cudaFree(d_pad_ptr);
    if (e != cudaSuccess) {
        set_error("level schedule upload failed: %s", cudaGetErrorString(e));
        return PHB200_ERR_CUDA;
    }
    return PHB200_OK;
}

static int build_plan(phb200_trsv_plan** out, const CsrView& A, int lower, cudaStream_t stream) {
    const int64_t n = A.n_rows;
    std::vector<int32_t> rowptr(n + 1), col((size_t)A.nnz);
    PHB_CUDA(cudaMemcpyAsync(rowptr.data(), A.rowptr, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost, stream));
    if (A.nnz) PHB_CUDA(cudaMemcpyAsync(col.data(), A.col, sizeof(int32_t) * A.nnz, cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    std::vector<int32_t> level(n, 0);
    int32_t n_levels = 0;
    if (lower) {
        for (int64_t i = 0; i < n; ++i) {
            int32_t l = 0;
            for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
                if (col[k] < i && level[col[k]] + 1 > l) l = level[col[k]] + 1;
            level[i] = l;
            if (l + 1 > n_levels) n_levels = l + 1;
        }
    } else {
        for (int64_t i = n - 1; i >= 0; --i) {
            int32_t l = 0;
            for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
                if (col[k] > i && level[col[k]] + 1 > l) l = level[col[k]] + 1;
            level[i] = l;
            if (l + 1 > n_levels) n_levels = l + 1;
        }
    }
    std::vector<int32_t> level_ptr(n_levels + 1, 0), order(n);
    for (int64_t i = 0; i < n; ++i) level_ptr[level[i] + 1]++;
    int32_t max_rows = 0;
    for (int32_t l = 0; l < n_levels; ++l) {
        if (level_ptr[l + 1] > max_rows) max_rows = level_ptr[l + 1];
        level_ptr[l + 1] += level_ptr[l];
    }
    std::vector<int32_t> cursor(level_ptr.begin(), level_ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i) order[cursor[level[i]]++] = (int32_t)i;
    phb200_trsv_plan* p = new phb200_trsv_plan();
    memset(p, 0, sizeof(*p));
    p->lower = lower;
    p->n = n;
    cudaError_t e = cudaMalloc(&p->order, sizeof(int32_t) * (n > 0 ? n : 1));
    if (e == cudaSuccess) e = cudaMemcpy(p->order, order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("level schedule upload failed: %s", cudaGetErrorString(e));
        phb200_sptrsv_plan_destroy(p);
        return PHB200_ERR_CUDA;
    }
    int r = finish_plan(p, level_ptr, stream);
    if (r != PHB200_OK) {
        phb200_sptrsv_plan_destroy(p);
        return r;
    }
    *out = p;
    return PHB200_OK;
}

// ---- device-side analysis for star stencils: level = sum of the indices along the axes that carry a dependency ------
__global__ void __launch_bounds__(kBlock) k_star_levels(StarDesc st, int lower, int32_t* __restrict__ level, int32_t* __restrict__ row) {
    const int64_t n = st.nx * st.ny * st.nz;
    const bool dx = lower ? st.cxm != 0.0 : st.cxp != 0.0, dy = lower ? st.cym != 0.0 : st.cyp != 0.0, dz = lower ? st.czm != 0.0 : st.czp != 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        const int64_t iz = i % st.nz, iy = (i / st.nz) % st.ny, ix = i / (st.nz * st.ny);
        int64_t l = 0;
        if (dx) l += lower ? ix : st.nx - 1 - ix;
        if (dy) l += lower ? iy : st.ny - 1 - iy;
        if (dz) l += lower ? iz : st.nz - 1 - iz;
        level[i] = (int32_t)l;
        row[i] = (int32_t)i;
    }
}

__global__ void __launch_bounds__(kBlock) k_level_starts(const int32_t* __restrict__ sorted_level, int64_t n, int32_t n_levels, int32_t* __restrict__ level_ptr) {
    const int l = blockIdx.x * kBlock + threadIdx.x;
    if (l > n_levels) return;
    int64_t lo = 0, hi = n;   // first k with sorted_level[k] >= l
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_level[mid] < l) lo = mid + 1;
        else hi = mid;
    }
    level_ptr[l] = (int32_t)lo;
}

int build_plan_star(phb200_trsv_plan** out, const StarDesc& st, int lower, cudaStream_t stream) {
    const int64_t n = st.nx * st.ny * st.nz;
    const bool dx = lower ? st.cxm != 0.0 : st.cxp != 0.0, dy = lower ? st.cym != 0.0 : st.cyp != 0.0, dz = lower ? st.czm != 0.0 : st.czp != 0.0;
    const int64_t n_levels = 1 + (dx ? st.nx - 1 : 0) + (dy ? st.ny - 1 : 0) + (dz ? st.nz - 1 : 0);
    phb200_trsv_plan* p = new phb200_trsv_plan();
    memset(p, 0, sizeof(*p));
    p->lower = lower;
    p->n = n;
    int32_t *lev = nullptr, *row = nullptr, *lev_sorted = nullptr, *d_level_ptr = nullptr;
    void* tmp = nullptr;
    int r = PHB200_OK;
    std::vector<int32_t> level_ptr(n_levels + 1);
    do {
        cudaError_t e = cudaMalloc(&p->order, sizeof(int32_t) * (n > 0 ? n : 1));
        if (e == cudaSuccess) e = phb_malloc_async(&lev, sizeof(int32_t) * n, stream);
        if (e == cudaSuccess) e = phb_malloc_async(&row, sizeof(int32_t) * n, stream);
        if (e == cudaSuccess) e = phb_malloc_async(&lev_sorted, sizeof(int32_t) * n, stream);
        if (e == cudaSuccess) e = phb_malloc_async(&d_level_ptr, sizeof(int32_t) * (n_levels + 1), stream);
        if (e != cudaSuccess) { set_error("level analysis allocation failed: %s", cudaGetErrorString(e)); r = PHB200_ERR_CUDA; break; }
        k_star_levels<<<(unsigned)grid_x((n + kBlock - 1) / kBlock, 1), kBlock, 0, stream>>>(st, lower, lev, row);
        int bits = 1;
        while ((1ll << bits) < n_levels) ++bits;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, lev, lev_sorted, row, p->order, (int)n, 0, bits, stream);
        if (phb_malloc_async(&tmp, tmp_bytes, stream) != cudaSuccess) { set_error("sort scratch allocation failed"); r = PHB200_ERR_CUDA; break; }
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, lev, lev_sorted, row, p->order, (int)n, 0, bits, stream);   // stable: rows ascending inside a level
        k_level_starts<<<(unsigned)((n_levels + 1 + kBlock - 1) / kBlock), kBlock, 0, stream>>>(lev_sorted, n, (int32_t)n_levels, d_level_ptr);
        g_launches.fetch_add(4);
        e = cudaMemcpyAsync(level_ptr.data(), d_level_ptr, sizeof(int32_t) * (n_levels + 1), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { set_error("level analysis failed: %s", cudaGetErrorString(e)); r = PHB200_ERR_CUDA; break; }
        r = finish_plan(p, level_ptr, stream);
    } while (0);
    if (lev) cudaFreeAsync(lev, stream);
    if (row) cudaFreeAsync(row, stream);
    if (lev_sorted) cudaFreeAsync(lev_sorted, stream);
    if (d_level_ptr) cudaFreeAsync(d_level_ptr, stream);
    if (tmp) cudaFreeAsync(tmp, stream);
    if (r != PHB200_OK) {
        phb200_sptrsv_plan_destroy(p);
        return r;
    }
    *out = p;
    return PHB200_OK;
}

// ---------------------------------------------------------------------------------------------- triangular solve
// One launch per level: x_i = (b_i - sum_{j in deps} T_ij x_j) / T_ii for the rows of the level and every right-hand side.
template <typename T, bool LOWER>
__global__ void __launch_bounds__(kBlock) k_trsv_level(CsrView A, const int32_t* __restrict__ order, int r0, int r1, int unit_diagonal,
                                                       const T* __restrict__ B, T* __restrict__ X, int64_t ld) {
    const int b = blockIdx.y;
    const T* __restrict__ val = (const T*)A.val;
    const int k = r0 + blockIdx.x * kBlock + threadIdx.x;
    if (k >= r1) return;
    const int i = order[k];
    T acc = B[(int64_t)b * ld + i];
    T diag = T(1);
    const T* xb = X + (int64_t)b * ld;
    for (int e = A.rowptr[i]; e < A.rowptr[i + 1]; ++e) {
        const int j = A.col[e];
        if (LOWER ? (j < i) : (j > i)) acc -= val[e] * xb[j];
        else if (j == i) diag = val[e];
    }
    X[(int64_t)b * ld + i] = unit_diagonal ? acc : acc / diag;
}

// ---- single-launch execution ----------------------------------------------------------------------------------------
// Rows sorted by level, levels padded to warp multiples: warp w owns 32 rows of ONE level l.  Lane 0 polls a per-level
// counter until every warp of level l-1 has finished (by induction all earlier levels are then complete), the warp computes
// its rows and adds 1 to the counter of level l (one atomic per warp).  All waited-for warps sit at smaller warp indices,
// i.e. were dispatched earlier, so the chain always advances; one 8-byte poll per warp keeps the L2 quiet (per-row flags
// polled by every thread saturate it).  Counters accumulate over solves (target = epoch * warps of the level): no reset.
// A wait that lasts longer than ~2 s traps instead of hanging the GPU.
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ int g_trsv_max_ns = 1024;   // tuning knobs (PHB200_TRSV_NS / PHB200_TRSV_FENCE), set once from the host
__device__ int g_trsv_fence_all = 1;
__device__ __forceinline__ void wait_level(const unsigned long long* cnt, const int32_t* __restrict__ level_warps, int level, int epoch) {
    const int max_ns = g_trsv_max_ns, fence_all = g_trsv_fence_all;
    if (level > 0 && (threadIdx.x & 31) == 0) {
        const unsigned long long target = (unsigned long long)epoch * (unsigned long long)level_warps[level - 1];
        if (ld_relaxed_u64(cnt + level - 1) < target) {
            const long long t0 = clock64();
            unsigned ns = 32;
            while (ld_relaxed_u64(cnt + level - 1) < target) {
                if (max_ns > 0) __nanosleep(ns);
                if ((int)ns < max_ns) ns *= 2;
                if (clock64() - t0 > 4000000000ll) __trap();
            }
        }
        if (!fence_all) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncwarp();
    if (fence_all) asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void publish_level(unsigned long long* cnt, int level) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        __threadfence();
        atomicAdd(cnt + level, 1ull);
    }
}

template <typename T, bool LOWER>
__global__ void __launch_bounds__(kBlock) k_trsv_syncfree(CsrView A, const int32_t* __restrict__ order_pad, int64_t npad, int unit_diagonal,
                                                          const T* __restrict__ B, T* X, int64_t ld, int batch, const int32_t* __restrict__ warp_level,
                                                          const int32_t* __restrict__ level_warps, unsigned long long* level_cnt, int epoch) {
    const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (k >= npad) return;
    const int level = warp_level[k >> 5];
    const int i = order_pad[k];
    const T* __restrict__ val = (const T*)A.val;
    // Everything that does not depend on other rows is fetched BEFORE the wait: the row's entries (up to kPre of them in
    // registers) and its right-hand side.  After the wait only the x gathers remain, issued back to back (one L2 latency).
    constexpr int kPre = 8;
    int cols[kPre];
    T vals[kPre];
    int e0 = 0, e1 = 0;
    T rhs0 = T(0);
    T diag = T(1);
    if (i >= 0) {
        e0 = A.rowptr[i];
        e1 = A.rowptr[i + 1];
        rhs0 = B[i];
#pragma unroll
        for (int p = 0; p < kPre; ++p) {
            const bool in = e0 + p < e1;
            cols[p] = in ? A.col[e0 + p] : -1;
            vals[p] = in ? val[e0 + p] : T(0);
        }
#pragma unroll
        for (int p = 0; p < kPre; ++p) {
            if (cols[p] == i) diag = vals[p];
            if (cols[p] >= 0 && !(LOWER ? (cols[p] < i) : (cols[p] > i))) cols[p] = -1;   // keep dependencies only
        }
    }
    wait_level(level_cnt, level_warps, level, epoch);
    if (i >= 0) {
        for (int b = 0; b < batch; ++b) {
            T acc = b == 0 ? rhs0 : B[(int64_t)b * ld + i];
            const T* xb = X + (int64_t)b * ld;
            T xs[kPre];
#pragma unroll
            for (int p = 0; p < kPre; ++p) xs[p] = cols[p] >= 0 ? __ldcg(xb + cols[p]) : T(0);   // written by other SMs: read through L2
#pragma unroll
            for (int p = 0; p < kPre; ++p) acc -= vals[p] * xs[p];
            for (int e = e0 + kPre; e < e1; ++e) {   // rows longer than kPre entries
                const int j = A.col[e];
                if (LOWER ? (j < i) : (j > i)) acc -= val[e] * __ldcg(xb + j);
                else if (j == i) diag = val[e];
            }
            X[(int64_t)b * ld + i] = unit_diagonal ? acc : acc / diag;
        }
    }
    publish_level(level_cnt, level);
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_ilu0_syncfree(CsrView A, T* lu, const int32_t* __restrict__ diag_pos, const int32_t* __restrict__ order_pad,
                                                          int64_t npad, const int32_t* __restrict__ warp_level, const int32_t* __restrict__ level_warps,
                                                          unsigned long long* level_cnt, int epoch) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t >= npad) return;
    const int level = warp_level[t >> 5];
    const int i = order_pad[t];
    wait_level(level_cnt, level_warps, level, epoch);
    if (i >= 0) {
        const int beg = A.rowptr[i], end = A.rowptr[i + 1];
        for (int p = beg; p < end; ++p) {
            const int k = A.col[p];
            if (k >= i) break;
            const int dk = diag_pos[k];
            const T lik = lu[p] / __ldcg(lu + dk);
            lu[p] = lik;
            int q = dk + 1;
            const int qend = A.rowptr[k + 1];
            int c = p + 1;
            while (q < qend && c < end) {
                const int jq = A.col[q], jc = A.col[c];
                if (jq == jc) {
                    lu[c] = lu[c] - lik * __ldcg(lu + q);
                    ++q;
                    ++c;
                } else if (jq < jc) {
                    ++q;
                } else {
                    ++c;
                }
            }
        }
    }
    publish_level(level_cnt, level);
}

static bool use_level_launches() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PHB200_TRSV");
        v = (e && strcmp(e, "levels") == 0) ? 1 : 0;
        const char* ns = getenv("PHB200_TRSV_NS");
        const char* fa = getenv("PHB200_TRSV_FENCE");
        if (ns) { int x = atoi(ns); cudaMemcpyToSymbol(g_trsv_max_ns, &x, sizeof(int)); }
        if (fa) { int x = atoi(fa); cudaMemcpyToSymbol(g_trsv_fence_all, &x, sizeof(int)); }
    }
    return v == 1;
}

template <typename T>
static int trsv(const CsrView& A, const phb200_trsv_plan* plan, int lower, int unit_diagonal, const T* B, T* X, int64_t batch, int64_t ld, cudaStream_t stream) {
    if (!use_level_launches()) {
        if (plan->npad == 0) return PHB200_OK;
        const int epoch = ++const_cast<phb200_trsv_plan*>(plan)->epoch;
        const unsigned grid = (unsigned)((plan->npad + kBlock - 1) / kBlock);
        if (lower) PHB_LAUNCH((k_trsv_syncfree<T, true>), grid, kBlock, 0, stream, A, plan->order_pad, plan->npad, unit_diagonal, B, X, ld, (int)batch, plan->warp_level, plan->level_warps, plan->level_cnt, epoch);
        else PHB_LAUNCH((k_trsv_syncfree<T, false>), grid, kBlock, 0, stream, A, plan->order_pad, plan->npad, unit_diagonal, B, X, ld, (int)batch, plan->warp_level, plan->level_warps, plan->level_cnt, epoch);
        return PHB200_OK;
    }
    for (int64_t l = 0; l < plan->n_levels; ++l) {
        const int r0 = plan->h_level_ptr[l], r1 = plan->h_level_ptr[l + 1];
        if (r1 == r0) continue;
        dim3 grid((unsigned)((r1 - r0 + kBlock - 1) / kBlock), (unsigned)batch);
        if (lower) PHB_LAUNCH((k_trsv_level<T, true>), grid, kBlock, 0, stream, A, plan->order, r0, r1, unit_diagonal, B, X, ld);
        else PHB_LAUNCH((k_trsv_level<T, false>), grid, kBlock, 0, stream, A, plan->order, r0, r1, unit_diagonal, B, X, ld);
    }
    return PHB200_OK;
}

// ---------------------------------------------------------------------------------------------- ILU(0)
template <typename T>
__global__ void __launch_bounds__(kBlock) k_diag_pos(CsrView A, int32_t* __restrict__ diag_pos, int* missing) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < A.n_rows; i += (int64_t)gridDim.x * kBlock) {
        int32_t pos = -1;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
            if (A.col[k] == (int32_t)i) { pos = k; break; }
        if (pos < 0) *missing = 1;
        diag_pos[i] = pos;
    }
}

// IKJ elimination of the rows of one level (they only depend on rows of earlier levels).  Columns are sorted, so
// the update a_ij -= a_ik a_kj walks row i and row k with two cursors.
template <typename T>
__global__ void __launch_bounds__(kBlock) k_ilu0_level(CsrView A, T* __restrict__ lu, const int32_t* __restrict__ diag_pos,
                                                       const int32_t* __restrict__ order, int r0, int r1) {
    const int t = r0 + blockIdx.x * kBlock + threadIdx.x;
    if (t >= r1) return;
    const int i = order[t];
    const int beg = A.rowptr[i], end = A.rowptr[i + 1];
    for (int p = beg; p < end; ++p) {
        const int k = A.col[p];
        if (k >= i) break;
        const T lik = lu[p] / lu[diag_pos[k]];
        lu[p] = lik;
        int q = diag_pos[k] + 1;
        const int qend = A.rowptr[k + 1];
        int c = p + 1;
        while (q < qend && c < end) {
            const int jq = A.col[q], jc = A.col[c];
            if (jq == jc) {
                lu[c] = lu[c] - lik * lu[q];
                ++q;
                ++c;
            } else if (jq < jc) {
                ++q;
            } else {
                ++c;
            }
        }
    }
}

template <typename T>
static int ilu0_factor(const phb200_precond* p, cudaStream_t stream) {
    const phb200_trsv_plan* plan = p->plan_l;
    T* lu = (T*)p->lu.val;
    if (!use_level_launches()) {
        if (plan->npad == 0) return PHB200_OK;
        const int epoch = ++const_cast<phb200_trsv_plan*>(plan)->epoch;
        PHB_LAUNCH(k_ilu0_syncfree<T>, (unsigned)((plan->npad + kBlock - 1) / kBlock), kBlock, 0, stream, p->lu, lu, p->diag_pos, plan->order_pad, plan->npad, plan->warp_level, plan->level_warps, plan->level_cnt, epoch);
        return PHB200_OK;
    }
    for (int64_t l = 0; l < plan->n_levels; ++l) {
        const int r0 = plan->h_level_ptr[l], r1 = plan->h_level_ptr[l + 1];
        if (r1 == r0) continue;
        PHB_LAUNCH(k_ilu0_level<T>, (unsigned)((r1 - r0 + kBlock - 1) / kBlock), kBlock, 0, stream, p->lu, lu, p->diag_pos, plan->order, r0, r1);
    }
    return PHB200_OK;
}

// ---------------------------------------------------------------------------------------------- ILU fixed-point sweeps
// incomplete_lu_coo (phiml/backend/_linalg.py:524-594; T.K. Huckle, "Parallel approximate LU factorizations"): Jacobi-style
// sweeps on the stored pattern.  State lu: diagonal + strict upper = U, strict lower = L.
//   init  (:555-556)   lu = A on the diagonal, A_ij / A_jj below it, 0 above
//   sweep (:573-577)   s_ij = sum_{k < min(i,j)} L_ik U_kj  (entries present in the pattern, OLD values, ascending k — the order
//                      in which strict_lu_mm_pattern_coo enumerates the matches; products and sums rounded separately)
//                      lower: (A_ij - s_ij) / lu_old[j,j]   (divide_no_nan when `safe`);   else: A_ij - s_ij
// All entries are updated simultaneously from the previous iterate, so two buffers alternate.  One thread per row; rows have
// sorted columns, the partner entry (k, j) is found by bisection in row k's upper part.
template <typename T>
__global__ void __launch_bounds__(kBlock) k_ilu_sweep_init(CsrView A, const int32_t* __restrict__ diag_pos, T* __restrict__ lu) {
    const T* __restrict__ a = (const T*)A.val;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < A.n_rows; i += (int64_t)gridDim.x * kBlock) {
        for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
            const int j = A.col[p];
            lu[p] = j == (int)i ? a[p] : (j < (int)i ? a[p] / a[diag_pos[j]] : T(0));
        }
    }
}

// Sum of the products in the order NumPy's einsum adds them (the reference computes s_ij with
// einsum('bnkc,bnkc->bnc', ...) over K = half_density padded slots, _linalg.py:575): 128-bit SIMD lanes (V = 2 doubles / 4
// floats; slot t goes to lane t mod V), whole groups of 4 vectors are chained last-to-first, the remaining vectors first-to-last,
// then the lanes are added horizontally ((l0 + l1) + (l2 + l3)); multiply and add are rounded separately (no FMA in the SSE
// baseline).  Reproducing the order makes the factors BIT-identical to the reference's (tests/test_gpu_parity.py checks
// it against factors recorded from the reference).  Padded slots hold zero products and change nothing.
constexpr int kIluMaxTerms = 32;
template <typename T>
__device__ __forceinline__ T einsum_order_sum(const T* pr, int cnt, int K) {
    constexpr int V = sizeof(T) == 8 ? 2 : 4;
    T lane[V];
#pragma unroll
    for (int l = 0; l < V; ++l) lane[l] = T(0);
    const int nvec = (cnt + V - 1) / V;
    const int full = (K / (4 * V)) * 4;
    for (int g = 0; g < full && g < nvec; g += 4)
        for (int v = g + 3; v >= g; --v) {
            if (v >= nvec) continue;
#pragma unroll
            for (int l = 0; l < V; ++l) { const int t = v * V + l; if (t < cnt) lane[l] = add_rn(lane[l], pr[t]); }
        }
    for (int v = full; v < nvec; ++v) {
#pragma unroll
        for (int l = 0; l < V; ++l) { const int t = v * V + l; if (t < cnt) lane[l] = add_rn(lane[l], pr[t]); }
    }
    if constexpr (V == 2) return add_rn(lane[0], lane[1]);
    else return add_rn(add_rn(lane[0], lane[1]), add_rn(lane[2], lane[3]));
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_ilu_sweep(CsrView A, const int32_t* __restrict__ diag_pos, const T* __restrict__ lu_old, T* __restrict__ lu_new, int safe, int K) {
    const T* __restrict__ a = (const T*)A.val;
    const int32_t* __restrict__ col = A.col;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < A.n_rows; i += (int64_t)gridDim.x * kBlock) {
        const int beg = A.rowptr[i], end = A.rowptr[i + 1];
        for (int p = beg; p < end; ++p) {
            const int j = col[p];
            const int m = j < (int)i ? j : (int)i;
            T pr[kIluMaxTerms];
            int cnt = 0;
            T s_over = T(0);      // rows with more matches than kIluMaxTerms: the rest is added in ascending order
            for (int q = beg; q < end; ++q) {
                const int k = col[q];
                if (k >= m) break;
                int lo = diag_pos[k] + 1, hi = A.rowptr[k + 1];      // strict upper part of row k, sorted
                const int rend = hi;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (col[mid] < j) lo = mid + 1;
                    else hi = mid;
                }
                if (lo < rend && col[lo] == j) {
                    const T prod = mul_rn(lu_old[lo], lu_old[q]);
                    if (cnt < kIluMaxTerms) pr[cnt++] = prod;
                    else s_over = add_rn(s_over, prod);
                }
            }
            const T s = add_rn(einsum_order_sum<T>(pr, cnt, K), s_over);
            const T rest = add_rn(a[p], -s);
            if (j < (int)i) {
                const T d = lu_old[diag_pos[j]];
                lu_new[p] = (safe && d == T(0)) ? T(0) : rest / d;
            } else {
                lu_new[p] = rest;
            }
        }
    }
}

// half_density of strict_lu_mm_pattern_coo (_linalg.py:617): max(strict-lower entries in a row, strict-upper entries in a column)
__global__ void __launch_bounds__(kBlock) k_ilu_half_density(CsrView A, const int32_t* __restrict__ diag_pos, int32_t* __restrict__ col_upper, int* __restrict__ out, int phase) {
    int best = 0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < A.n_rows; i += (int64_t)gridDim.x * kBlock) {
        if (phase == 0) {
            const int lower = diag_pos[i] - A.rowptr[i];
            if (lower > best) best = lower;
            for (int p = diag_pos[i] + 1; p < A.rowptr[i + 1]; ++p) atomicAdd(&col_upper[A.col[p]], 1);
        } else if (col_upper[i] > best) {
            best = col_upper[i];
        }
    }
    if (best > 0) atomicMax(out, best);
}

// returns the iterate before the last sweep in *prev (pool memory, caller frees with cudaFreeAsync) — the star check needs it
template <typename T>
static int ilu_sweeps_factor(const phb200_precond* p, const CsrView& A, int sweeps, int safe, T** prev, cudaStream_t stream) {
    T* out = (T*)p->lu.val;
    T* tmp = nullptr;
    int32_t* col_upper = nullptr;
    int* d_k = nullptr;
    PHB_CUDA(phb_malloc_async(&tmp, sizeof(T) * (size_t)(A.nnz > 0 ? A.nnz : 1), stream));
    *prev = tmp;
    const unsigned grid = (unsigned)grid_x((A.n_rows + kBlock - 1) / kBlock, 1);
    PHB_CUDA(phb_malloc_async(&col_upper, sizeof(int32_t) * (size_t)(A.n_rows > 0 ? A.n_rows : 1), stream));
    PHB_CUDA(phb_malloc_async(&d_k, sizeof(int), stream));
    PHB_CUDA(cudaMemsetAsync(col_upper, 0, sizeof(int32_t) * (size_t)(A.n_rows > 0 ? A.n_rows : 1), stream));
    PHB_CUDA(cudaMemsetAsync(d_k, 0, sizeof(int), stream));
    PHB_LAUNCH(k_ilu_half_density, grid, kBlock, 0, stream, A, p->diag_pos, col_upper, d_k, 0);
    PHB_LAUNCH(k_ilu_half_density, grid, kBlock, 0, stream, A, p->diag_pos, col_upper, d_k, 1);
    int K = 0;
    PHB_CUDA(cudaMemcpyAsync(&K, d_k, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    PHB_CUDA(cudaFreeAsync(col_upper, stream));
    PHB_CUDA(cudaFreeAsync(d_k, stream));
    // the last sweep must write `out`: sweep s (1-based) writes buf[(sweeps - s) & 1], the initial state sits in buf[sweeps & 1]
    T* buf[2] = {out, tmp};
    PHB_LAUNCH(k_ilu_sweep_init<T>, grid, kBlock, 0, stream, A, p->diag_pos, buf[sweeps & 1]);
    for (int s = 1; s <= sweeps; ++s)
        PHB_LAUNCH(k_ilu_sweep<T>, grid, kBlock, 0, stream, A, p->diag_pos, (const T*)buf[(sweeps - s + 1) & 1], buf[(sweeps - s) & 1], safe, K);
    return PHB200_OK;
}

// ---------------------------------------------------------------------------------------------- star wavefront solves
template <typename T, bool LOWER>
static int star_trsv_launch(phb200_precond* p, const T* B, T* X, int64_t batch, int64_t ld, cudaStream_t stream) {
    const StarDesc& st = p->star;
    StarTrsvArgs a;
    a.nx = (int)st.nx; a.ny = (int)st.ny; a.nz = (int)st.nz;
    int tx_pref = 8;
    if (const char* e = getenv("PHB200_TRSV_TX")) { const int v = atoi(e); if (v >= 1 && v <= kTrMaxTX) tx_pref = v; }
    a.tx = (int)(st.nx < tx_pref ? st.nx : tx_pref);
    a.nsx = (a.nx + a.tx - 1) / a.tx;
    a.nsy = (a.ny + 31) / 32;
    a.nch = (a.nz + kTrZC - 1) / kTrZC;
    a.batch = (int)batch;
    a.ld = ld;
    a.rd = LOWER ? p->rd_l : p->rd;
    a.B = B;
    a.X = X;
    const int64_t need = batch * a.nsx * a.nsy;
    if (need > p->tr_flag_cap) {   // first use / larger batch: fresh counters (all solves so far have completed in stream order)
        PHB_CUDA(cudaStreamSynchronize(stream));
        if (p->tr_flags) cudaFree(p->tr_flags);
        p->tr_flags = nullptr;
        PHB_CUDA(cudaMalloc(&p->tr_flags, sizeof(unsigned long long) * need));
        PHB_CUDA(cudaMemsetAsync(p->tr_flags, 0, sizeof(unsigned long long) * need, stream));
        p->tr_flag_cap = need;
        p->tr_epoch = 0;
    }
    a.flags = p->tr_flags;
    a.epoch_base = p->tr_epoch;
    p->tr_epoch += (unsigned long long)a.nch;
    a.ticket = p->tr_ticket;
    constexpr int W = tr_warps<T>();
    a.n_ctas = (unsigned)((need + W - 1) / W);
    a.cx = LOWER ? st.cxm : st.cxp;
    a.cy = LOWER ? st.cym : st.cyp;
    a.cz = LOWER ? st.czm : st.czp;
    auto kernel = k_star_trsv<T, LOWER>;
    constexpr size_t smem = trsv_warp_smem<T>() * W;
    static bool attr_dev[kMaxDevices] = {};
    bool& attr = attr_dev[device_slot()];
    if (!attr) {
        PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    PHB_LAUNCH(kernel, a.n_ctas, W * 32, smem, stream, a);
    return PHB200_OK;
}

template <typename T, bool LOWER>
static int scan_trsv_launch(phb200_precond* p, const T* B, T* X, int64_t batch, int64_t ld, cudaStream_t stream) {
    const StarDesc& st = p->star;
    ScanTrsvArgs a;
    a.nx = (int)st.nx; a.ny = (int)st.ny; a.nz = (int)st.nz;
    a.ni = (a.nx + kScT - 1) / kScT;
    a.nj = (a.ny + kScT - 1) / kScT;
    a.nch = (a.nz + kScZ - 1) / kScZ;
    a.batch = (int)batch;
    a.ld = ld;
    a.rd = LOWER ? p->rd_l : p->rd;
    a.B = B;
    a.X = X;
    const int64_t need = batch * a.ni * a.nj;
    const int64_t n_flags = need * kScT;      // the dataflow kernel keeps one counter per tile row, the barrier kernel uses the first of each tile... of its own layout
    if (n_flags > p->tr_flag_cap) {
        PHB_CUDA(cudaStreamSynchronize(stream));
        if (p->tr_flags) cudaFree(p->tr_flags);
        p->tr_flags = nullptr;
        PHB_CUDA(cudaMalloc(&p->tr_flags, sizeof(unsigned long long) * n_flags));
        PHB_CUDA(cudaMemsetAsync(p->tr_flags, 0, sizeof(unsigned long long) * n_flags, stream));
        p->tr_flag_cap = n_flags;
        p->tr_epoch = 0;
    }
    a.flags = p->tr_flags;
    a.epoch_base = p->tr_epoch;
    p->tr_epoch += (unsigned long long)a.nch;
    a.ticket = p->tr_ticket;
    a.n_ctas = (unsigned)need;
    a.order = p->sc_order;
    a.dbg = nullptr;
    if (const char* e = getenv("PHB200_TRSV_DBG")) a.dbg = (unsigned long long*)strtoull(e, nullptr, 0);
    a.cx = LOWER ? st.cxm : st.cxp;
    a.cy = LOWER ? st.cym : st.cyp;
    a.cz = LOWER ? st.czm : st.czp;
    if (p->star_mode == 3) {
        auto kernel = k_star_trsv_flow<T, LOWER>;
        constexpr size_t smem = flow_smem_bytes<T>();
        static bool attr_dev[kMaxDevices] = {};
        bool& attr = attr_dev[device_slot()];
        if (!attr) {
            PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        PHB_LAUNCH(kernel, a.n_ctas, kScT * 32, smem, stream, a);
        return PHB200_OK;
    }
    auto kernel = k_star_trsv_scan<T, LOWER>;
    constexpr size_t smem = scan_smem_bytes<T>();
    static bool attr_dev[kMaxDevices] = {};
    bool& attr = attr_dev[device_slot()];
    if (!attr) {
        PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    PHB_LAUNCH(kernel, a.n_ctas, kScT * 32, smem, stream, a);
    return PHB200_OK;
}

// x-marching row pipelines (star_trsv_pipe.cuh): R rows per CTA, one CTA per (row tile, z chunk)
constexpr int kPpRows = 8;

template <typename T, bool LOWER>
static int pipe_trsv_launch(phb200_precond* p, const T* B, T* X, int64_t batch, int64_t ld, cudaStream_t stream) {
    const StarDesc& st = p->star;
    constexpr int R = kPpRows;
    PipeTrsvArgs a;
    a.nx = (int)st.nx; a.ny = (int)st.ny; a.nz = (int)st.nz;
    a.nj = (a.ny + R - 1) / R;
    constexpr int ZC = 32 * pipe_vec<T>();
    a.nch = (a.nz + ZC - 1) / ZC;
    a.batch = (int)batch;
    a.ld = ld;
    a.rd = LOWER ? p->rd_l : p->rd;
    a.B = B;
    a.X = X;
    const int64_t n_ctas = batch * a.nj * a.nch;
    constexpr int64_t W = sizeof(T) / 4;
    const int64_t y_words = batch * a.nj * (int64_t)a.nx * a.nch * ZC * W, z_words = batch * a.nch * (int64_t)a.nx * a.ny * W;
    const bool grow = y_words + z_words > p->pp_words;
    if (grow || (unsigned long long)p->pp_epoch + (unsigned long long)a.nx + 2ull >= (1ull << 32)) {
        // (re)start with empty hand-over words: first use, a larger batch, or the 32-bit tag space is used up
        if (grow) {
            PHB_CUDA(cudaStreamSynchronize(stream));
            if (p->pp_buf) cudaFree(p->pp_buf);
            p->pp_buf = nullptr;
            PHB_CUDA(cudaMalloc(&p->pp_buf, sizeof(unsigned long long) * (size_t)(y_words + z_words)));
            p->pp_words = y_words + z_words;
        }
        PHB_CUDA(cudaMemsetAsync(p->pp_buf, 0, sizeof(unsigned long long) * (size_t)p->pp_words, stream));
        p->pp_epoch = 0;
    }
    a.ybuf = (unsigned long long*)p->pp_buf;
    a.zbuf = a.ybuf + y_words;
    a.tag_base = p->pp_epoch;
    a.dbg = nullptr;
    if (const char* e = getenv("PHB200_TRSV_DBG")) a.dbg = (unsigned long long*)strtoull(e, nullptr, 0);
    p->pp_epoch += (unsigned)a.nx + 1u;
    a.ticket = p->tr_ticket;
    a.n_ctas = (unsigned)n_ctas;
    a.order = p->pp_order;
    a.cx = LOWER ? st.cxm : st.cxp;
    a.cy = LOWER ? st.cym : st.cyp;
    a.cz = LOWER ? st.czm : st.czp;
    auto kernel = k_star_trsv_pipe<T, LOWER, R>;
    constexpr size_t smem = pipe_smem_bytes<T, R>();
    static bool attr_dev[kMaxDevices] = {};
    bool& attr = attr_dev[device_slot()];
    if (!attr) {
        PHB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    PHB_LAUNCH(kernel, a.n_ctas, R * 32, smem, stream, a);
    return PHB200_OK;
}

static int pipe_order_setup(phb200_precond* p, const StarDesc& st, cudaStream_t stream) {
    const int zc = 32 * (p->dtype == PHB200_F32 ? 4 : 2);
    const int nj = (int)((st.ny + kPpRows - 1) / kPpRows), nch = (int)((st.nz + zc - 1) / zc);
    std::vector<int32_t> order;
    order.reserve((size_t)nj * nch);
    for (int dg = 0; dg < nj + nch - 1; ++dg)
        for (int j = dg < nch ? 0 : dg - nch + 1; j <= dg && j < nj; ++j) order.push_back(j * nch + (dg - j));
    PHB_CUDA(cudaMalloc(&p->pp_order, sizeof(int32_t) * order.size()));
    PHB_CUDA(cudaMemcpyAsync(p->pp_order, order.data(), sizeof(int32_t) * order.size(), cudaMemcpyHostToDevice, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    return PHB200_OK;
}

static int scan_order_setup(phb200_precond* p, const StarDesc& st, cudaStream_t stream) {
    const int ni = (int)((st.nx + kScT - 1) / kScT), nj = (int)((st.ny + kScT - 1) / kScT);
    std::vector<int32_t> order;
    order.reserve((size_t)ni * nj);
    for (int dg = 0; dg < ni + nj - 1; ++dg)
        for (int i = dg < nj ? 0 : dg - nj + 1; i <= dg && i < ni; ++i) order.push_back(i * nj + (dg - i));
    PHB_CUDA(cudaMalloc(&p->sc_order, sizeof(int32_t) * order.size()));
    PHB_CUDA(cudaMemcpyAsync(p->sc_order, order.data(), sizeof(int32_t) * order.size(), cudaMemcpyHostToDevice, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    return PHB200_OK;
}

template <typename T>
static int star_setup(phb200_precond* p, const phb200_matrix* twin, const T* lu_prev, cudaStream_t stream) {
    const StarDesc& st = twin->star;
    // closed form needs distinct, non-aliasing offsets: nz >= 4 (star kernels need nz % 4 == 0 anyway), ny != 2 in 3-D
    if (st.nz < 4 || st.nz % 4 != 0) return PHB200_OK;
    if (st.nx > 1 && st.ny < 3) return PHB200_OK;
    if (st.xb != 0 || st.xe != st.nx) return PHB200_OK;
    if (st.nx > 0x3fffffff / 8 || st.nx * st.ny * st.nz != p->n) return PHB200_OK;
    const char* e = getenv("PHB200_TRSV");
    if (e && (strcmp(e, "levels") == 0 || strcmp(e, "csr") == 0)) return PHB200_OK;
    int* d_ok = nullptr;
    PHB_CUDA(cudaMalloc(&p->rd, sizeof(T) * (size_t)p->n));
    p->rd_l = p->rd;
    if (lu_prev) PHB_CUDA(cudaMalloc(&p->rd_l, sizeof(T) * (size_t)p->n));
    PHB_CUDA(cudaMalloc(&p->tr_ticket, sizeof(unsigned)));
    PHB_CUDA(cudaMemsetAsync(p->tr_ticket, 0, sizeof(unsigned), stream));
    PHB_CUDA(phb_malloc_async(&d_ok, sizeof(int), stream));
    int one = 1;
    PHB_CUDA(cudaMemcpyAsync(d_ok, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
    PHB_LAUNCH(k_star_ilu_check<T>, (unsigned)grid_x((p->n + kBlock - 1) / kBlock, 1), kBlock, 0, stream, p->lu, (const T*)p->lu.val, lu_prev ? lu_prev : (const T*)p->lu.val,
               p->diag_pos, st, (T*)p->rd, (T*)p->rd_l, d_ok);
    int ok = 0;
    PHB_CUDA(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    PHB_CUDA(cudaFreeAsync(d_ok, stream));
    if (ok) {
        p->star = st;
        // default: x-marching row pipelines when there are planes to march over (3-D), row scans otherwise; PHB200_TRSV=scan|flow|skew|pipe
        const bool pipe_default = st.nx >= 16 && st.nx * st.ny * st.nz < (1ll << 31);
        p->star_mode = (e && strcmp(e, "skew") == 0) ? 1 : (e && strcmp(e, "flow") == 0) ? 3 : (e && strcmp(e, "scan") == 0) ? 2 : (e && strcmp(e, "pipe") == 0) ? 4
                       : (pipe_default ? 4 : 2);
        if (p->star_mode == 2 || p->star_mode == 3) PHB_TRY(scan_order_setup(p, st, stream));
        if (p->star_mode == 4) PHB_TRY(pipe_order_setup(p, st, stream));
    } else {
        if (p->rd_l != p->rd) cudaFree(p->rd_l);
        cudaFree(p->rd);
        p->rd = p->rd_l = nullptr;
    }
    return PHB200_OK;
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_scale_rows(const T* __restrict__ in, T* __restrict__ out, const T* __restrict__ w, int64_t n, int64_t ld) {
    const int b = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        out[(int64_t)b * ld + i] = mul_rn(in[(int64_t)b * ld + i], w[i]);
}

// ---------------------------------------------------------------------------------------------- 'cluster' preconditioner
// ExplicitClusterSolve.apply (phiml/backend/_linalg.py:741-753):  coarse = bincount(clusters, weights=vec);  delta = vec - (coarse / size)[clusters];
// result = (inv_coarse @ coarse)[clusters] + delta.
// k_cluster_restrict: one CTA per (cluster, system); members are summed in a fixed order (thread-strided partial sums in double, then a
// shared-memory tree), so the result does not depend on scheduling — the reference's bincount adds in element order, torch's with atomics.
template <typename T>
__global__ void __launch_bounds__(kBlock) k_cluster_restrict(const T* __restrict__ in, int64_t ld, const int32_t* __restrict__ members, const int32_t* __restrict__ ptr,
                                                              int count, T* __restrict__ coarse) {
    __shared__ double part[kBlock];
    const int c = blockIdx.x, b = blockIdx.y;
    const T* v = in + (int64_t)b * ld;
    double acc = 0.0;
    for (int j = ptr[c] + threadIdx.x; j < ptr[c + 1]; j += kBlock) acc += (double)v[members[j]];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int w = kBlock / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) coarse[(int64_t)b * count + c] = (T)part[0];
}

// coarse_solution = inv_coarse @ coarse: one warp per (row, system), lanes stride over the row, sums in double
template <typename T>
__global__ void __launch_bounds__(kBlock) k_cluster_coarse_solve(const T* __restrict__ inv, const T* __restrict__ coarse, int count, T* __restrict__ sol) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (kBlock / 32) + warp, b = blockIdx.y;
    if (i >= count) return;
    const T* row = inv + (int64_t)i * count;
    const T* cv = coarse + (int64_t)b * count;
    double acc = 0.0;
    for (int j = lane; j < count; j += 32) acc += (double)row[j] * (double)cv[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sol[(int64_t)b * count + i] = (T)acc;
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_cluster_prolong(const T* __restrict__ in, T* __restrict__ out, int64_t n, int64_t ld, const int32_t* __restrict__ cl_of,
                                                             const T* __restrict__ coarse, const T* __restrict__ sol, const T* __restrict__ size, int count) {
    const int b = blockIdx.y;
    const T* cv = coarse + (int64_t)b * count;
    const T* sv = sol + (int64_t)b * count;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        const int c = cl_of[i];
        const T mean = cv[c] / size[c];
        const T delta = add_rn(in[(int64_t)b * ld + i], -mean);
        out[(int64_t)b * ld + i] = add_rn(sv[c], delta);
    }
}

template <typename T>
static int cluster_apply(const phb200_precond* p, const T* in, T* out, int64_t batch, int64_t ld, cudaStream_t stream) {
    phb200_precond* q = const_cast<phb200_precond*>(p);     // scratch is bookkeeping
    if (batch > q->cl_batch_cap) {
        if (q->cl_scratch) PHB_CUDA(cudaFree(q->cl_scratch));
        q->cl_scratch = nullptr;
        PHB_CUDA(cudaMalloc(&q->cl_scratch, sizeof(T) * 2 * (size_t)batch * (size_t)p->cl_count));
        q->cl_batch_cap = batch;
    }
    T* coarse = (T*)q->cl_scratch;
    T* sol = coarse + (size_t)batch * p->cl_count;
    const int C = p->cl_count;
    PHB_LAUNCH(k_cluster_restrict<T>, dim3((unsigned)C, (unsigned)batch), kBlock, 0, stream, in, ld, p->cl_members, p->cl_ptr, C, coarse);
    PHB_LAUNCH(k_cluster_coarse_solve<T>, dim3((unsigned)((C + kBlock / 32 - 1) / (kBlock / 32)), (unsigned)batch), kBlock, 0, stream, (const T*)p->cl_inv, (const T*)coarse, C, sol);
    dim3 grid((unsigned)grid_x((p->n + kBlock - 1) / kBlock, batch), (unsigned)batch);
    PHB_LAUNCH(k_cluster_prolong<T>, grid, kBlock, 0, stream, in, out, p->n, ld, p->cl_of, (const T*)coarse, (const T*)sol, (const T*)p->cl_size, C);
    return PHB200_OK;
}

int precond_apply_inv_l(const phb200_precond* p, const void* in, void* out, int64_t batch, int64_t ld, cudaStream_t stream) {
    if (p->kind == PHB200_PRE_JACOBI) {
        dim3 grid((unsigned)grid_x((p->n + kBlock - 1) / kBlock, batch), (unsigned)batch);
        if (p->dtype == PHB200_F32) PHB_LAUNCH(k_scale_rows<float>, grid, kBlock, 0, stream, (const float*)in, (float*)out, (const float*)p->inv_diag, p->n, ld);
        else PHB_LAUNCH(k_scale_rows<double>, grid, kBlock, 0, stream, (const double*)in, (double*)out, (const double*)p->inv_diag, p->n, ld);
        return PHB200_OK;
    }
    if (p->kind == PHB200_PRE_ILU) {
        if (p->star_mode) {
            phb200_precond* q = const_cast<phb200_precond*>(p);   // counters / epoch are bookkeeping
            if (p->star_mode == 4) {
                if (p->dtype == PHB200_F32) return pipe_trsv_launch<float, true>(q, (const float*)in, (float*)out, batch, ld, stream);
                return pipe_trsv_launch<double, true>(q, (const double*)in, (double*)out, batch, ld, stream);
            }
            if (p->star_mode >= 2) {
                if (p->dtype == PHB200_F32) return scan_trsv_launch<float, true>(q, (const float*)in, (float*)out, batch, ld, stream);
                return scan_trsv_launch<double, true>(q, (const double*)in, (double*)out, batch, ld, stream);
            }
            if (p->dtype == PHB200_F32) return star_trsv_launch<float, true>(q, (const float*)in, (float*)out, batch, ld, stream);
            return star_trsv_launch<double, true>(q, (const double*)in, (double*)out, batch, ld, stream);
        }
        if (p->dtype == PHB200_F32) return trsv<float>(p->lu, p->plan_l, 1, 1, (const float*)in, (float*)out, batch, ld, stream);
        return trsv<double>(p->lu, p->plan_l, 1, 1, (const double*)in, (double*)out, batch, ld, stream);
    }
    if (p->kind == PHB200_PRE_CLUSTER) {          // ExplicitClusterSolve.apply_inv_l: identity (_linalg.py:758-759)
        const size_t w = p->dtype == PHB200_F32 ? 4 : 8;
        if (in != out) PHB_CUDA(cudaMemcpy2DAsync(out, (size_t)ld * w, in, (size_t)ld * w, (size_t)p->n * w, (size_t)batch, cudaMemcpyDeviceToDevice, stream));
        return PHB200_OK;
    }
    set_error("preconditioner kind %d has no apply_inv_l", p->kind);
    return PHB200_ERR_UNSUPPORTED;
}

int precond_apply(const phb200_precond* p, const void* in, void* out, void* tmp, int64_t batch, int64_t ld, cudaStream_t stream) {
    if (p->kind == PHB200_PRE_JACOBI) return precond_apply_inv_l(p, in, out, batch, ld, stream);
    if (p->kind == PHB200_PRE_ILU) {
        // out = U^-1 (L^-1 in); the upper solve may run in place, so `tmp` is not needed
        (void)tmp;
        PHB_TRY(precond_apply_inv_l(p, in, out, batch, ld, stream));
        if (p->star_mode) {
            phb200_precond* q = const_cast<phb200_precond*>(p);
            if (p->star_mode == 4) {
                if (p->dtype == PHB200_F32) return pipe_trsv_launch<float, false>(q, (const float*)out, (float*)out, batch, ld, stream);
                return pipe_trsv_launch<double, false>(q, (const double*)out, (double*)out, batch, ld, stream);
            }
            if (p->star_mode >= 2) {
                if (p->dtype == PHB200_F32) return scan_trsv_launch<float, false>(q, (const float*)out, (float*)out, batch, ld, stream);
                return scan_trsv_launch<double, false>(q, (const double*)out, (double*)out, batch, ld, stream);
            }
            if (p->dtype == PHB200_F32) return star_trsv_launch<float, false>(q, (const float*)out, (float*)out, batch, ld, stream);
            return star_trsv_launch<double, false>(q, (const double*)out, (double*)out, batch, ld, stream);
        }
        if (p->dtype == PHB200_F32) return trsv<float>(p->lu, p->plan_u, 0, 0, (const float*)out, (float*)out, batch, ld, stream);
        return trsv<double>(p->lu, p->plan_u, 0, 0, (const double*)out, (double*)out, batch, ld, stream);
    }
    if (p->kind == PHB200_PRE_CLUSTER) {
        if (p->dtype == PHB200_F32) return cluster_apply<float>(p, (const float*)in, (float*)out, batch, ld, stream);
        return cluster_apply<double>(p, (const double*)in, (double*)out, batch, ld, stream);
    }
    set_error("preconditioner kind %d cannot be applied", p->kind);
    return PHB200_ERR_UNSUPPORTED;
}


// ---------------------------------------------------------------------------------------------- CSR transposition
// A^T of a CSR matrix as CSR (= the CSC arrays of A): stable radix sort of the entry numbers by column (entries of one column
// keep their row order, so the rows of A^T come out sorted), row starts by bisection on the sorted keys, then one gather pass.
__global__ void __launch_bounds__(kBlock) k_iota(int32_t* __restrict__ out, int64_t n) {
    for (int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x; k < n; k += (int64_t)gridDim.x * kBlock) out[k] = (int32_t)k;
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_transpose_fill(const int32_t* __restrict__ rowptr, int64_t n_rows, const T* __restrict__ val, const int32_t* __restrict__ perm,
                                                            int64_t nnz, int32_t* __restrict__ t_col, T* __restrict__ t_val) {
    for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < nnz; j += (int64_t)gridDim.x * kBlock) {
        const int32_t k = perm[j];
        int64_t lo = 0, hi = n_rows;          // largest r with rowptr[r] <= k
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (rowptr[mid] <= k) lo = mid; else hi = mid;
        }
        t_col[j] = (int32_t)lo;
        t_val[j] = val[k];
    }
}

}  // namespace phb

using namespace phb;

extern "C" {

int phb200_jacobi_setup(const phb200_matrix* A, void* inv_diag, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(A && inv_diag, "null argument");
    PHB_ARG(A->n_rows == A->n_cols, "matrix must be square");
    const int64_t n = A->n_rows;
    const unsigned gridx = (unsigned)grid_x((n + kBlock - 1) / kBlock, 1);
    if (A->kind == PHB200_MAT_STENCIL) {
        PHB_ARG(A->st.diag >= 0, "stencil has no diagonal entry");
        const double c = A->st.coef[A->st.diag];
        if (A->dtype == PHB200_F32) PHB_LAUNCH(k_jacobi_const<float>, gridx, kBlock, 0, stream, (float*)inv_diag, n, 1.0f / (float)c);
        else PHB_LAUNCH(k_jacobi_const<double>, gridx, kBlock, 0, stream, (double*)inv_diag, n, 1.0 / c);
        return PHB200_OK;
    }
    // the diagonal of A^T is the diagonal of A, so transposed views need no special case
    int* d_missing = nullptr;
    PHB_CUDA(phb_malloc_async(&d_missing, sizeof(int), stream));
    PHB_CUDA(cudaMemsetAsync(d_missing, 0, sizeof(int), stream));
    if (A->dtype == PHB200_F32) PHB_LAUNCH(k_jacobi_csr<float>, gridx, kBlock, 0, stream, A->csr, (float*)inv_diag, d_missing);
    else PHB_LAUNCH(k_jacobi_csr<double>, gridx, kBlock, 0, stream, A->csr, (double*)inv_diag, d_missing);
    int h = 0;
    PHB_CUDA(cudaMemcpyAsync(&h, d_missing, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    PHB_CUDA(cudaFreeAsync(d_missing, stream));
    PHB_ARG(!h, "All diagonal elements must be present in sparse matrix.");
    return PHB200_OK;
}

int phb200_precond_jacobi(phb200_precond** out, const void* inv_diag, int64_t n, int dtype, int constant) {
    PHB_ARG(out && inv_diag && n >= 0, "null argument");
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    phb200_precond* p = new phb200_precond();
    memset(p, 0, sizeof(*p));
    p->kind = PHB200_PRE_JACOBI;
    p->dtype = dtype;
    p->n = n;
    p->inv_diag = inv_diag;
    if (constant && n > 0) {   // one element tells the whole diagonal (blocking copy, set-up only)
        float f = 0.f;
        double d = 0.0;
        cudaError_t e = dtype == PHB200_F32 ? cudaMemcpy(&f, inv_diag, sizeof(float), cudaMemcpyDeviceToHost) : cudaMemcpy(&d, inv_diag, sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            delete p;
            set_error("reading the constant inverse diagonal failed: %s", cudaGetErrorString(e));
            return PHB200_ERR_CUDA;
        }
        p->is_const = 1;
        p->const_inv = dtype == PHB200_F32 ? (double)f : d;
    }
    *out = p;
    return PHB200_OK;
}

int phb200_precond_cluster(phb200_precond** out, const int32_t* cluster_of, const int32_t* members, const int32_t* member_ptr, int64_t n,
                           int cluster_count, const void* inv_coarse, const void* cluster_size, int dtype) {
    PHB_ARG(out && cluster_of && members && member_ptr && inv_coarse && cluster_size, "null argument");
    PHB_ARG(n >= 1 && cluster_count >= 1 && cluster_count <= 65535, "need 1 <= cluster_count <= 65535, got %d", cluster_count);
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    phb200_precond* p = new phb200_precond();
    memset(p, 0, sizeof(*p));
    p->kind = PHB200_PRE_CLUSTER;
    p->dtype = dtype;
    p->n = n;
    p->cl_of = cluster_of;
    p->cl_members = members;
    p->cl_ptr = member_ptr;
    p->cl_inv = inv_coarse;
    p->cl_size = cluster_size;
    p->cl_count = cluster_count;
    *out = p;
    return PHB200_OK;
}

static int ilu_create(phb200_precond** out, const phb200_matrix* A, void* lu_val, int sweeps, int safe, const phb200_matrix* stencil_twin, cudaStream_t stream) {
    PHB_ARG(out && A && lu_val, "null argument");
    PHB_ARG(A->kind == PHB200_MAT_CSR && !A->transposed, "ILU(0) needs a plain CSR matrix");
    PHB_ARG(A->n_rows == A->n_cols, "incomplete_lu_coo only implemented for square matrices");
    PHB_ARG(lu_val != A->csr.val || sweeps < 0, "the sweeps read the matrix values: lu_val must not alias them");
    const size_t w = A->dtype == PHB200_F32 ? 4 : 8;
    phb200_precond* p = new phb200_precond();
    memset(p, 0, sizeof(*p));
    p->kind = PHB200_PRE_ILU;
    p->dtype = A->dtype;
    p->n = A->n_rows;
    p->lu = A->csr;
    p->lu.val = lu_val;
    p->sweeps = sweeps;
    int r = PHB200_OK;
    int* d_missing = nullptr;
    void* prev = nullptr;
    do {
        if (cudaMalloc(&p->diag_pos, sizeof(int32_t) * (p->n > 0 ? p->n : 1)) != cudaSuccess) { set_error("cudaMalloc failed"); r = PHB200_ERR_CUDA; break; }
        if (phb_malloc_async(&d_missing, sizeof(int), stream) != cudaSuccess) { set_error("cudaMallocAsync failed"); r = PHB200_ERR_CUDA; break; }
        cudaMemsetAsync(d_missing, 0, sizeof(int), stream);
        if (sweeps < 0 && lu_val != A->csr.val) cudaMemcpyAsync(lu_val, A->csr.val, w * A->nnz, cudaMemcpyDeviceToDevice, stream);
        const unsigned gridx = (unsigned)grid_x((p->n + kBlock - 1) / kBlock, 1);
        if (A->dtype == PHB200_F32) k_diag_pos<float><<<gridx, kBlock, 0, stream>>>(A->csr, p->diag_pos, d_missing);
        else k_diag_pos<double><<<gridx, kBlock, 0, stream>>>(A->csr, p->diag_pos, d_missing);
        g_launches.fetch_add(1);
        int h = 0;
        cudaMemcpyAsync(&h, d_missing, sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (cudaStreamSynchronize(stream) != cudaSuccess) { set_error("diagonal search failed: %s", cudaGetErrorString(cudaGetLastError())); r = PHB200_ERR_CUDA; break; }
        if (h) { set_error("All diagonal elements must be present in sparse matrix."); r = PHB200_ERR_ARG; break; }
        // a recognised star stencil has analytic dependency levels (computed on the device, no host copy of the pattern)
        const bool star = stencil_twin && stencil_twin->kind == PHB200_MAT_STENCIL && stencil_twin->is_star && stencil_twin->n_rows == A->n_rows;
        r = star ? build_plan_star(&p->plan_l, stencil_twin->star, 1, stream) : build_plan(&p->plan_l, A->csr, 1, stream);
        if (r != PHB200_OK) break;
        r = star ? build_plan_star(&p->plan_u, stencil_twin->star, 0, stream) : build_plan(&p->plan_u, A->csr, 0, stream);
        if (r != PHB200_OK) break;
        if (sweeps < 0) {
            r = A->dtype == PHB200_F32 ? ilu0_factor<float>(p, stream) : ilu0_factor<double>(p, stream);
        } else {
            r = A->dtype == PHB200_F32 ? ilu_sweeps_factor<float>(p, A->csr, sweeps, safe, (float**)&prev, stream) : ilu_sweeps_factor<double>(p, A->csr, sweeps, safe, (double**)&prev, stream);
        }
        if (r != PHB200_OK) break;
        if (star) r = A->dtype == PHB200_F32 ? star_setup<float>(p, stencil_twin, (const float*)prev, stream) : star_setup<double>(p, stencil_twin, (const double*)prev, stream);
    } while (0);
    if (d_missing) cudaFreeAsync(d_missing, stream);
    if (prev) cudaFreeAsync(prev, stream);
    if (r != PHB200_OK) {
        phb200_precond_destroy(p);
        return r;
    }
    *out = p;
    return PHB200_OK;
}

int phb200_precond_ilu0(phb200_precond** out, const phb200_matrix* A, void* lu_val, const phb200_matrix* stencil_twin, void* stream_) {
    return ilu_create(out, A, lu_val, -1, 0, stencil_twin, (cudaStream_t)stream_);
}

int phb200_precond_ilu_sweeps(phb200_precond** out, const phb200_matrix* A, void* lu_val, int sweeps, int safe, const phb200_matrix* stencil_twin, void* stream_) {
    PHB_ARG(sweeps >= 1, "incomplete_lu_coo needs at least one sweep, got %d", sweeps);
    return ilu_create(out, A, lu_val, sweeps, safe ? 1 : 0, stencil_twin, (cudaStream_t)stream_);
}

int phb200_precond_info(const phb200_precond* p, int* kind, int* star_wavefront) {
    PHB_ARG(p, "null preconditioner");
    if (kind) *kind = p->kind;
    if (star_wavefront) *star_wavefront = p->star_mode;
    return PHB200_OK;
}

int phb200_precond_levels(const phb200_precond* p, int64_t* lower_levels, int64_t* upper_levels) {
    PHB_ARG(p && p->kind == PHB200_PRE_ILU, "need an ILU preconditioner");
    if (lower_levels) *lower_levels = p->plan_l->n_levels;
    if (upper_levels) *upper_levels = p->plan_u->n_levels;
    return PHB200_OK;
}

int phb200_sptrsv_plan(phb200_trsv_plan** out, const phb200_matrix* T, int lower, void* stream_) {
    PHB_ARG(out && T, "null argument");
    PHB_ARG(T->kind == PHB200_MAT_CSR && !T->transposed, "triangular solves need a plain CSR matrix");
    PHB_ARG(T->n_rows == T->n_cols, "triangular matrix must be square");
    return build_plan(out, T->csr, lower ? 1 : 0, (cudaStream_t)stream_);
}

int phb200_sptrsv(const phb200_matrix* T, const phb200_trsv_plan* plan, int lower, int unit_diagonal,
                  const void* B, void* X, int64_t batch, int64_t ld, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(T && B && X, "null argument");
    PHB_ARG(T->kind == PHB200_MAT_CSR && !T->transposed, "triangular solves need a plain CSR matrix");
    PHB_ARG(batch >= 0 && batch <= 65535 && ld >= T->n_rows, "bad batch / leading dimension");
    if (batch == 0 || T->n_rows == 0) return PHB200_OK;
    phb200_trsv_plan* own = nullptr;
    if (!plan) {
        PHB_TRY(build_plan(&own, T->csr, lower ? 1 : 0, stream));
        plan = own;
    }
    int r;
    if (plan->lower != (lower ? 1 : 0) || plan->n != T->n_rows) {
        set_error("plan does not match this solve");
        r = PHB200_ERR_ARG;
    } else if (T->dtype == PHB200_F32) {
        r = trsv<float>(T->csr, plan, lower ? 1 : 0, unit_diagonal, (const float*)B, (float*)X, batch, ld, stream);
    } else {
        r = trsv<double>(T->csr, plan, lower ? 1 : 0, unit_diagonal, (const double*)B, (double*)X, batch, ld, stream);
    }
    if (own) {
        cudaStreamSynchronize(stream);
        phb200_sptrsv_plan_destroy(own);
    }
    return r;
}

int phb200_sptrsv_plan_destroy(phb200_trsv_plan* p) {
    if (!p) return PHB200_OK;
    if (p->order) cudaFree(p->order);
    if (p->order_pad) cudaFree(p->order_pad);
    if (p->warp_level) cudaFree(p->warp_level);
    if (p->level_warps) cudaFree(p->level_warps);
    if (p->level_cnt) cudaFree(p->level_cnt);
    if (p->level_ptr) cudaFree(p->level_ptr);
    if (p->h_level_ptr) free(p->h_level_ptr);
    delete p;
    return PHB200_OK;
}

int phb200_precond_apply(const phb200_precond* p, const void* in, void* out, int64_t batch, int64_t ld, void* stream_) {
    PHB_ARG(p && in && out, "null argument");
    PHB_ARG(batch >= 0 && batch <= 65535 && ld >= p->n, "bad batch / leading dimension");
    if (batch == 0 || p->n == 0) return PHB200_OK;
    return precond_apply(p, in, out, nullptr, batch, ld, (cudaStream_t)stream_);
}

int phb200_precond_destroy(phb200_precond* p) {
    if (!p) return PHB200_OK;
    if (p->diag_pos) cudaFree(p->diag_pos);
    if (p->rd_l && p->rd_l != p->rd) cudaFree(p->rd_l);
    if (p->rd) cudaFree(p->rd);
    if (p->tr_flags) cudaFree(p->tr_flags);
    if (p->tr_ticket) cudaFree(p->tr_ticket);
    if (p->sc_order) cudaFree(p->sc_order);
    if (p->pp_order) cudaFree(p->pp_order);
    if (p->pp_buf) cudaFree(p->pp_buf);
    if (p->cl_scratch) cudaFree(p->cl_scratch);
    phb200_sptrsv_plan_destroy(p->plan_l);
    phb200_sptrsv_plan_destroy(p->plan_u);
    delete p;
    return PHB200_OK;
}

int phb200_csr_transpose(const int32_t* rowptr, const int32_t* col, const void* val, int64_t n_rows, int64_t n_cols, int64_t nnz, int dtype,
                         int32_t* t_rowptr, int32_t* t_col, void* t_val, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(rowptr && t_rowptr && (nnz == 0 || (col && val && t_col && t_val)), "null argument");
    PHB_ARG(n_rows >= 0 && n_cols >= 0 && nnz >= 0 && nnz <= 2147483647LL && n_cols <= 2147483647LL - 2 * kBlock, "sizes out of range for int32 CSR");
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    if (nnz == 0) {
        PHB_CUDA(cudaMemsetAsync(t_rowptr, 0, sizeof(int32_t) * (size_t)(n_cols + 1), stream));
        return PHB200_OK;
    }
    int32_t *iota = nullptr, *keys = nullptr, *perm = nullptr;
    void* tmp = nullptr;
    int r = PHB200_OK;
    do {
        cudaError_t e = phb_malloc_async(&iota, sizeof(int32_t) * (size_t)nnz, stream);
        if (e == cudaSuccess) e = phb_malloc_async(&keys, sizeof(int32_t) * (size_t)nnz, stream);
        if (e == cudaSuccess) e = phb_malloc_async(&perm, sizeof(int32_t) * (size_t)nnz, stream);
        if (e != cudaSuccess) { set_error("transposition scratch allocation failed: %s", cudaGetErrorString(e)); r = PHB200_ERR_CUDA; break; }
        k_iota<<<(unsigned)grid_x((nnz + kBlock - 1) / kBlock, 1), kBlock, 0, stream>>>(iota, nnz);
        int bits = 1;
        while ((1ll << bits) < n_cols) ++bits;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, col, keys, iota, perm, (int)nnz, 0, bits, stream);
        if (phb_malloc_async(&tmp, tmp_bytes, stream) != cudaSuccess) { set_error("sort scratch allocation failed"); r = PHB200_ERR_CUDA; break; }
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, col, keys, iota, perm, (int)nnz, 0, bits, stream);   // stable: rows ascending inside a column
        k_level_starts<<<(unsigned)((n_cols + 1 + kBlock - 1) / kBlock), kBlock, 0, stream>>>(keys, nnz, (int32_t)n_cols, t_rowptr);
        const unsigned grid = (unsigned)grid_x((nnz + kBlock - 1) / kBlock, 1);
        if (dtype == PHB200_F32) k_transpose_fill<float><<<grid, kBlock, 0, stream>>>(rowptr, n_rows, (const float*)val, perm, nnz, t_col, (float*)t_val);
        else k_transpose_fill<double><<<grid, kBlock, 0, stream>>>(rowptr, n_rows, (const double*)val, perm, nnz, t_col, (double*)t_val);
        g_launches.fetch_add(5);
        if (cudaGetLastError() != cudaSuccess) { set_error("transposition launch failed"); r = PHB200_ERR_CUDA; break; }
    } while (0);
    if (iota) cudaFreeAsync(iota, stream);
    if (keys) cudaFreeAsync(keys, stream);
    if (perm) cudaFreeAsync(perm, stream);
    if (tmp) cudaFreeAsync(tmp, stream);
    return r;
}

}  // extern "C"
