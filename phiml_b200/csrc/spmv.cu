// Matrix handles, stand-alone products, stencil detection and device-side CSR assembly.
#include <cuda.h>
#include "dispatch.cuh"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <vector>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

namespace phb {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

bool debug_sync() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PHB200_DEBUG_SYNC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PHB200_PDL");   // off by default: measured 6 % SLOWER on the 2-kernel CG iteration (256^3)
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

L2Window& l2_window() {
    static thread_local L2Window w{nullptr, 0, 0.f};
    return w;
}

// PHB200_L2_PERSIST=q|r (experimental; 1 = q): sizes the persisting part of L2 once per device and returns the window for `bytes` at
// `base`; {nullptr} when the option is off or the device has no persisting L2.
int l2_persist_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("PHB200_L2_PERSIST");
        mode = !e ? 0 : (e[0] == '1' || e[0] == 'q') ? 1 : e[0] == 'r' ? 2 : 0;
    }
    return mode;
}

L2Window l2_persist_window(void* base, size_t bytes) {
    static int ready = -1;
    static size_t carve = 0, max_window = 0;
    if (ready < 0) {
        ready = 0;
        if (l2_persist_mode()) {
            int dev = 0;
            cudaDeviceProp prop;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
                carve = (size_t)prop.persistingL2CacheMaxSize;
                if (const char* f = getenv("PHB200_L2_CARVE_MB")) carve = std::min(carve, (size_t)atoll(f) << 20);
                max_window = (size_t)prop.accessPolicyMaxWindowSize;
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) ready = 1;
                else cudaGetLastError();
                if (getenv("PHB200_L2_VERBOSE")) fprintf(stderr, "[phb200] L2 %d MB, persisting max %zu MB (carve %zu MB), window max %zu MB\n", prop.l2CacheSize >> 20, (size_t)prop.persistingL2CacheMaxSize >> 20, carve >> 20, max_window >> 20);
            } else {
                cudaGetLastError();
            }
        }
    }
    if (ready != 1 || !base || bytes == 0) return L2Window{nullptr, 0, 0.f};
    const size_t win = bytes < max_window ? bytes : max_window;
    const float hit = win <= carve ? 1.0f : (float)((double)carve / (double)win);
    return L2Window{base, win, hit};
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int occupancy_of(const void* kernel) {
    static std::mutex mu;
    static std::unordered_map<const void*, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(kernel);
    if (it != cache.end()) return it->second;
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kBlock, 0) != cudaSuccess || v < 1) {
        cudaGetLastError();
        v = 4;
    }
    cache[kernel] = v;
    return v;
}

// ---------------------------------------------------------------------------------------------- CSR analysis
__global__ void k_csr_block_max(const int32_t* __restrict__ rowptr, int64_t n_rows, int* out) {
    int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r0 = chunk * kCsrRows;
    int v = 0;
    if (r0 < n_rows) {
        int64_t r1 = min(n_rows, r0 + kCsrRows);
        v = rowptr[r1] - rowptr[r0];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

int csr_analyse(const phb200_matrix* m, cudaStream_t stream) {
    if (m->analysed) return PHB200_OK;
    int* d = nullptr;
    PHB_CUDA(phb_malloc_async(&d, sizeof(int), stream));
    PHB_CUDA(cudaMemsetAsync(d, 0, sizeof(int), stream));
    int64_t chunks = (m->csr.n_rows + kCsrRows - 1) / kCsrRows;
    PHB_LAUNCH(k_csr_block_max, (unsigned)((chunks + 255) / 256), 256, 0, stream, m->csr.rowptr, m->csr.n_rows, d);
    int h = 0;
    PHB_CUDA(cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    PHB_CUDA(cudaFreeAsync(d, stream));
    m->max_block_nnz = h;
    m->analysed = 1;
    return PHB200_OK;
}

// ---------------------------------------------------------------------------------------------- transposed / COO
template <typename T>
__global__ void __launch_bounds__(kBlock) k_spmv_csr_scatter(CsrView A, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y, int64_t ldy) {
    // arrays describe A^T: entry (r, c, v) of the arrays contributes v * x[r] to y[c]
    const int b = blockIdx.y;
    const T* __restrict__ val = (const T*)A.val;
    for (int64_t r = (int64_t)blockIdx.x * kBlock + threadIdx.x; r < A.n_rows; r += (int64_t)gridDim.x * kBlock) {
        const T xv = X[(int64_t)b * ldx + r];
        for (int k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k) atomicAdd(&Y[(int64_t)b * ldy + A.col[k]], val[k] * xv);
    }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_spmv_coo(const int64_t* __restrict__ row, const int64_t* __restrict__ col, const T* __restrict__ val,
                                                     int64_t nnz, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y, int64_t ldy) {
    const int b = blockIdx.y;
    for (int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * kBlock)
        atomicAdd(&Y[(int64_t)b * ldy + row[k]], val[k] * X[(int64_t)b * ldx + col[k]]);
}

// ---------------------------------------------------------------------------------------------- batched CSR x dense
// `matrix @ x` outside the solver: Backend.mul_csr_dense (_backend.py:1366-1387; torch _torch_backend.py:912-922 builds one CSR tensor
// per (batch, channel) pair in a Python double loop, numpy _numpy_backend.py:510-519 one SciPy matrix per pair).  ONE launch for the
// whole (batch, rows, channels, dense_cols) result: a thread owns one output element, consecutive threads own consecutive (channel,
// dense column) slots of a row and then consecutive rows, so the gathers of `dense` are contiguous over the inner width and the
// index loads of a row are shared by its threads.  Every output is summed over its row's entries in entry order with separately
// rounded products, starting from 0 — SciPy's csr_matvecs (y += a * x per entry), bit for bit, per (batch, channel, column).
template <typename T>
__global__ void __launch_bounds__(kBlock) k_csr_dense_batched(const int32_t* __restrict__ col, const int32_t* __restrict__ rowptr, const T* __restrict__ val,
                                                              const T* __restrict__ dense, T* __restrict__ out, int64_t batch, int64_t rows, int64_t cols,
                                                              int64_t nnz, int channels, int dcols) {
    const int64_t width = (int64_t)channels * dcols, total = batch * rows * width;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < total; i += (int64_t)gridDim.x * kBlock) {
        const int64_t t = i / width, v = i - t * width, b = t / rows, r = t - b * rows;
        const int c = (int)(v / dcols);
        const int32_t* __restrict__ rp = rowptr + b * (rows + 1);
        const int32_t* __restrict__ cb = col + b * nnz;
        const T* __restrict__ vb = val + b * nnz * channels + c;
        const T* __restrict__ db = dense + b * cols * width + v;
        T acc = T(0);
        for (int k = rp[r], k1 = rp[r + 1]; k < k1; ++k) acc = add_rn(acc, mul_rn(vb[(int64_t)k * channels], db[(int64_t)cb[k] * width]));
        out[i] = acc;
    }
}

// ---------------------------------------------------------------------------------------------- matrix gradient
// Sampled outer product of the implicit-gradient step (phiml/math/_optimize.py:803-807):
//   dm[k] = -sum_b dy[b,row[k]] * x[b,col[k]]     on the stored entries, batch summed in order, products rounded separately.
// COO form: int64 coordinates as torch stores them.  CSR form: a CTA walks chunks of kBlock rows; thread t labels the entries of
// row t with the byte t in shared memory, then all threads walk the chunk's entries coalesced (index traffic 4 B instead of 16 B
// per entry).
template <typename T>
__global__ void __launch_bounds__(kBlock) k_matrix_gradient_coo(const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t nnz,
                                                                 const T* __restrict__ DY, int64_t ld_dy, const T* __restrict__ X, int64_t ld_x, int batch,
                                                                 T* __restrict__ out) {
    for (int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * kBlock) {
        const int64_t r = row[k], c = col[k];
        T acc = T(0);
        for (int b = 0; b < batch; ++b) acc = add_rn(acc, mul_rn(DY[(int64_t)b * ld_dy + r], X[(int64_t)b * ld_x + c]));
        out[k] = -acc;
    }
}

constexpr int kGradWindow = 8192;   // entries per window of the CSR matrix-gradient kernel (row offsets as bytes in shared memory)

template <typename T>
__global__ void __launch_bounds__(kBlock) k_matrix_gradient_csr(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                                                                 const T* __restrict__ DY, int64_t ld_dy, const T* __restrict__ X, int64_t ld_x, int batch,
                                                                 T* __restrict__ out) {
    static_assert(kBlock <= 256, "row offsets inside a chunk are stored as bytes");
    __shared__ uint8_t rowof[kGradWindow];
    __shared__ int32_t ends[2];
    __shared__ T dys[kBlock];
    const int64_t chunks = (n_rows + kBlock - 1) / kBlock;
    const int t = threadIdx.x;
    for (int64_t ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const int64_t r0 = ch * kBlock;
        const int nr = (int)min((int64_t)kBlock, n_rows - r0);
        // thread t owns row r0 + t: it labels the row's entries, window by window (rows longer than a window span several)
        const int32_t mine0 = t < nr ? rowptr[r0 + t] : 0, mine1 = t < nr ? rowptr[r0 + t + 1] : 0;
        if (t == 0) ends[0] = mine0;
        if (t == nr - 1) ends[1] = mine1;
        if (batch == 1 && t < nr) dys[t] = DY[r0 + t];     // one right-hand side: dy of the chunk's rows comes from shared memory
        __syncthreads();
        const int32_t p0 = ends[0], p1 = ends[1];
        __syncthreads();      // ends[] is rewritten for the next chunk
        for (int32_t w0 = p0; w0 < p1; w0 += kGradWindow) {
            const int32_t w1 = min(p1, w0 + kGradWindow);
            for (int32_t k = max(mine0, w0); k < min(mine1, w1); ++k) rowof[k - w0] = (uint8_t)t;
            __syncthreads();
            for (int32_t k = w0 + t; k < w1; k += kBlock) {
                const int lr = rowof[k - w0];
                const int64_t r = r0 + lr, c = col[k];
                T acc = T(0);
                if (batch == 1) acc = add_rn(acc, mul_rn(dys[lr], X[c]));
                else
                    for (int b = 0; b < batch; ++b) acc = add_rn(acc, mul_rn(DY[(int64_t)b * ld_dy + r], X[(int64_t)b * ld_x + c]));
                out[k] = -acc;
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------- stencil <-> CSR
template <typename T>
__global__ void __launch_bounds__(kBlock) k_stencil_verify(StencilDesc st, CsrView A, int* mismatch) {
    const T* __restrict__ val = (const T*)A.val;
    const int64_t ny = st.dims[1], nz = st.dims[2];
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < A.n_rows; i += (int64_t)gridDim.x * kBlock) {
        const int64_t iz = i % nz, iy = (i / nz) % ny, ix = i / (nz * ny);
        int k = A.rowptr[i];
        const int k1 = A.rowptr[i + 1];
        bool ok = true;
        for (int p = 0; p < st.npts; ++p) {
            int64_t jx = ix + st.off[p][0], jy = iy + st.off[p][1], jz = iz + st.off[p][2];
            if (jx < 0 || jx >= st.dims[0] || jy < 0 || jy >= ny || jz < 0 || jz >= nz) continue;
            int64_t j = (jx * ny + jy) * nz + jz;
            if (k >= k1 || A.col[k] != (int32_t)j || val[k] != (T)st.coef[p]) { ok = false; break; }
            ++k;
        }
        if (!ok || k != k1) *mismatch = 1;
    }
}

__global__ void __launch_bounds__(kBlock) k_stencil_count(StencilDesc st, int64_t n, int32_t* __restrict__ counts) {
    const int64_t ny = st.dims[1], nz = st.dims[2];
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i <= n; i += (int64_t)gridDim.x * kBlock) {
        int c = 0;
        if (i < n) {
            const int64_t iz = i % nz, iy = (i / nz) % ny, ix = i / (nz * ny);
            for (int p = 0; p < st.npts; ++p) {
                int64_t jx = ix + st.off[p][0], jy = iy + st.off[p][1], jz = iz + st.off[p][2];
                c += !(jx < 0 || jx >= st.dims[0] || jy < 0 || jy >= ny || jz < 0 || jz >= nz);
            }
        }
        counts[i] = c;
    }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_stencil_fill(StencilDesc st, int64_t n, const int32_t* __restrict__ rowptr,
                                                         int32_t* __restrict__ col, T* __restrict__ val) {
    const int64_t ny = st.dims[1], nz = st.dims[2];
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        const int64_t iz = i % nz, iy = (i / nz) % ny, ix = i / (nz * ny);
        int k = rowptr[i];
        for (int p = 0; p < st.npts; ++p) {
            int64_t jx = ix + st.off[p][0], jy = iy + st.off[p][1], jz = iz + st.off[p][2];
            if (jx < 0 || jx >= st.dims[0] || jy < 0 || jy >= ny || jz < 0 || jz >= nz) continue;
            col[k] = (int32_t)((jx * ny + jy) * nz + jz);
            val[k] = (T)st.coef[p];
            ++k;
        }
    }
}

// ---------------------------------------------------------------------------------------------- shift tracer -> CSR
// General form of the device assembly: a ShiftLinTracer holds one full-grid value array per shift (phiml/math/_lin_trace.py:57-99);
// tracer_to_coo (:750-807) keeps the entries whose value is non-zero and wraps the shifted index periodically ((idx + shift) % size, :775),
// SparseCoordinateTensor.compress (phiml/math/_sparse.py:347-390) orders every row by column through SciPy.  Here: one thread per row counts its
// non-zero shifts, an exclusive scan gives the row pointers, the fill pass gathers (wrapped column, value) per non-zero shift and sorts the <= 16
// pairs of the row by column.  Two shifts that wrap onto the same column (grids smaller than the stencil) set the failure flag — the reference
// asserts in that case.
struct ShiftSet {
    int rank, nshifts;
    int64_t dims[3];
    int off[kMaxStencil][3];
};

__global__ void __launch_bounds__(kBlock) k_narrow_i64(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) out[i] = (int32_t)in[i];
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_shift_count(ShiftSet sh, int64_t n, const T* __restrict__ vals, int32_t* __restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i <= n; i += (int64_t)gridDim.x * kBlock) {
        int c = 0;
        if (i < n)
            for (int s = 0; s < sh.nshifts; ++s) c += vals[(int64_t)s * n + i] != T(0);
        counts[i] = c;
    }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_shift_fill(ShiftSet sh, int64_t n, const T* __restrict__ vals, const int32_t* __restrict__ rowptr,
                                                       int32_t* __restrict__ col, T* __restrict__ val, int* fail) {
    const int64_t ny = sh.dims[1], nz = sh.dims[2], nx = sh.dims[0];
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        const int64_t iz = i % nz, iy = (i / nz) % ny, ix = i / (nz * ny);
        int32_t c[kMaxStencil];
        T v[kMaxStencil];
        int m = 0;
        for (int s = 0; s < sh.nshifts; ++s) {
            const T a = vals[(int64_t)s * n + i];
            if (a == T(0)) continue;
            int64_t jx = (ix + sh.off[s][0]) % nx, jy = (iy + sh.off[s][1]) % ny, jz = (iz + sh.off[s][2]) % nz;
            if (jx < 0) jx += nx;
            if (jy < 0) jy += ny;
            if (jz < 0) jz += nz;
            const int32_t cc = (int32_t)((jx * ny + jy) * nz + jz);
            int p = m++;
            while (p > 0 && c[p - 1] > cc) { c[p] = c[p - 1]; v[p] = v[p - 1]; --p; }     // insertion sort by column
            c[p] = cc;
            v[p] = a;
        }
        const int k0 = rowptr[i];
        for (int p = 0; p < m; ++p) {
            if (p > 0 && c[p] == c[p - 1]) *fail = 1;
            col[k0 + p] = c[p];
            val[k0 + p] = v[p];
        }
    }
}

static int make_stencil(phb200_matrix** out, int rank, const int64_t* grid, int npts, const int32_t* offsets, const double* coeffs, int dtype) {
    PHB_ARG(out && grid && offsets && coeffs, "null argument");
    PHB_ARG(rank >= 1 && rank <= 3, "stencil rank must be 1..3, got %d", rank);
    PHB_ARG(npts >= 1 && npts <= kMaxStencil, "stencil needs 1..%d points, got %d", kMaxStencil, npts);
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    phb200_matrix* m = new phb200_matrix();
    memset(m, 0, sizeof(*m));
    m->kind = PHB200_MAT_STENCIL;
    m->dtype = dtype;
    StencilDesc& st = m->st;
    st.rank = rank;
    st.npts = npts;
    for (int d = 0; d < 3; ++d) st.dims[d] = 1;
    for (int d = 0; d < rank; ++d) {
        if (grid[d] < 1) { delete m; set_error("grid size must be positive"); return PHB200_ERR_ARG; }
        st.dims[3 - rank + d] = grid[d];
    }
    struct Pt { int o[3]; double c; int64_t flat; };
    std::vector<Pt> pts(npts);
    for (int p = 0; p < npts; ++p) {
        pts[p].o[0] = pts[p].o[1] = pts[p].o[2] = 0;
        for (int d = 0; d < rank; ++d) pts[p].o[3 - rank + d] = offsets[p * rank + d];
        pts[p].c = coeffs[p];
        pts[p].flat = ((int64_t)pts[p].o[0] * st.dims[1] + pts[p].o[1]) * st.dims[2] + pts[p].o[2];
    }
    std::sort(pts.begin(), pts.end(), [](const Pt& a, const Pt& b) { return a.flat < b.flat; });
    st.diag = -1;
    for (int p = 0; p < npts; ++p) {
        if (p > 0 && pts[p].flat == pts[p - 1].flat) { delete m; set_error("duplicate stencil offsets"); return PHB200_ERR_ARG; }
        for (int d = 0; d < 3; ++d) {
            st.off[p][d] = pts[p].o[d];
        }
        st.coef[p] = pts[p].c;
        if (pts[p].o[0] == 0 && pts[p].o[1] == 0 && pts[p].o[2] == 0) st.diag = p;
    }
    m->n_rows = m->n_cols = st.dims[0] * st.dims[1] * st.dims[2];
    m->nnz = -1;
    m->is_star = star_from_stencil(st, m->star) ? 1 : 0;
    *out = m;
    return PHB200_OK;
}

}  // namespace phb

using namespace phb;

extern "C" {

int phb200_version(void) { return PHB200_VERSION; }
const char* phb200_last_error(void) { return g_err; }
int64_t phb200_launch_count(void) { return (int64_t)g_launches.load(); }

// ---- CUDA IPC: lets the slab neighbours (one process per GPU) store halo planes straight into each other's vectors
int phb200_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset) {
    PHB_ARG(dev_ptr && handle64 && offset, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CUdeviceptr base = 0;
    size_t size = 0;
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            set_error("cuMemGetAddressRange not available");
            return PHB200_ERR_CUDA;
        }
        fn = (RangeFn)p;
    }
    if (fn(&base, &size, (CUdeviceptr)dev_ptr) != CUDA_SUCCESS) { set_error("cuMemGetAddressRange failed"); return PHB200_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    PHB_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
    memcpy(handle64, &h, 64);
    *offset = (int64_t)((CUdeviceptr)dev_ptr - base);
    return PHB200_OK;
}

int phb200_ipc_open(const void* handle64, void** base_ptr) {
    PHB_ARG(handle64 && base_ptr, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PHB_CUDA(cudaIpcOpenMemHandle(base_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PHB200_OK;
}

int phb200_ipc_close(void* base_ptr) {
    if (base_ptr) PHB_CUDA(cudaIpcCloseMemHandle(base_ptr));
    return PHB200_OK;
}

int phb200_matrix_csr(phb200_matrix** out, const int32_t* rowptr, const int32_t* col, const void* val,
                      int64_t n_rows, int64_t n_cols, int64_t nnz, int dtype, int transposed) {
    PHB_ARG(out && rowptr && (nnz == 0 || (col && val)), "null argument");
    PHB_ARG(n_rows >= 0 && n_cols >= 0 && nnz >= 0 && nnz < (1ll << 31) && n_rows < (1ll << 31) && n_cols < (1ll << 31), "CSR sizes must fit int32 (nnz=%lld)", (long long)nnz);
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    phb200_matrix* m = new phb200_matrix();
    memset(m, 0, sizeof(*m));
    m->kind = PHB200_MAT_CSR;
    m->dtype = dtype;
    m->transposed = transposed ? 1 : 0;
    m->csr = CsrView{rowptr, col, val, n_rows, n_cols, nnz};
    // logical shape of A (not of the stored arrays)
    m->n_rows = transposed ? n_cols : n_rows;
    m->n_cols = transposed ? n_rows : n_cols;
    m->nnz = nnz;
    *out = m;
    return PHB200_OK;
}

int phb200_matrix_stencil(phb200_matrix** out, int rank, const int64_t* grid, int npts, const int32_t* offsets, const double* coeffs, int dtype) {
    return make_stencil(out, rank, grid, npts, offsets, coeffs, dtype);
}

int phb200_stencil_from_csr(phb200_matrix** out, const phb200_matrix* csr, int rank, const int64_t* grid, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(out && csr && grid, "null argument");
    PHB_ARG(csr->kind == PHB200_MAT_CSR && !csr->transposed, "stencil detection needs a plain CSR matrix");
    PHB_ARG(rank >= 1 && rank <= 3, "rank must be 1..3");
    int64_t dims[3] = {1, 1, 1};
    int64_t n = 1;
    for (int d = 0; d < rank; ++d) { dims[3 - rank + d] = grid[d]; n *= grid[d]; }
    if (n != csr->csr.n_rows || n != csr->csr.n_cols) { set_error("grid volume %lld != matrix size %lld", (long long)n, (long long)csr->csr.n_rows); return PHB200_ERR_NOT_STENCIL; }
    // candidate stencil = the row of the centre cell
    const int64_t c[3] = {dims[0] / 2, dims[1] / 2, dims[2] / 2};
    const int64_t ic = (c[0] * dims[1] + c[1]) * dims[2] + c[2];
    int32_t rp[2];
    PHB_CUDA(cudaMemcpyAsync(rp, csr->csr.rowptr + ic, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    const int npts = rp[1] - rp[0];
    if (npts < 1 || npts > kMaxStencil) { set_error("centre row has %d entries", npts); return PHB200_ERR_NOT_STENCIL; }
    int32_t cols[kMaxStencil];
    double vals64[kMaxStencil];
    float vals32[kMaxStencil];
    PHB_CUDA(cudaMemcpyAsync(cols, csr->csr.col + rp[0], npts * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    if (csr->dtype == PHB200_F32) PHB_CUDA(cudaMemcpyAsync(vals32, (const float*)csr->csr.val + rp[0], npts * sizeof(float), cudaMemcpyDeviceToHost, stream));
    else PHB_CUDA(cudaMemcpyAsync(vals64, (const double*)csr->csr.val + rp[0], npts * sizeof(double), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    int32_t offs[kMaxStencil * 3];
    double coef[kMaxStencil];
    for (int p = 0; p < npts; ++p) {
        int64_t j = cols[p];
        int64_t jz = j % dims[2], jy = (j / dims[2]) % dims[1], jx = j / (dims[2] * dims[1]);
        const int64_t o3[3] = {jx - c[0], jy - c[1], jz - c[2]};
        for (int d = 0; d < rank; ++d) offs[p * rank + d] = (int32_t)o3[3 - rank + d];
        coef[p] = csr->dtype == PHB200_F32 ? (double)vals32[p] : vals64[p];
    }
    phb200_matrix* cand = nullptr;
    int r = make_stencil(&cand, rank, grid, npts, offs, coef, csr->dtype);
    if (r != PHB200_OK) return PHB200_ERR_NOT_STENCIL;
    int* d_flag = nullptr;
    PHB_CUDA(phb_malloc_async(&d_flag, sizeof(int), stream));
    PHB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), stream));
    const unsigned gridx = (unsigned)grid_x((n + kBlock - 1) / kBlock, 1);
    if (csr->dtype == PHB200_F32) PHB_LAUNCH(k_stencil_verify<float>, gridx, kBlock, 0, stream, cand->st, csr->csr, d_flag);
    else PHB_LAUNCH(k_stencil_verify<double>, gridx, kBlock, 0, stream, cand->st, csr->csr, d_flag);
    int h = 0;
    PHB_CUDA(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    PHB_CUDA(cudaFreeAsync(d_flag, stream));
    if (h) { delete cand; set_error("matrix is not a constant-coefficient ZERO-boundary stencil on this grid"); return PHB200_ERR_NOT_STENCIL; }
    cand->nnz = csr->nnz;
    *out = cand;
    return PHB200_OK;
}

int phb200_coded_from_csr(phb200_matrix** out, const phb200_matrix* csr, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(out && csr, "null argument");
    PHB_ARG(csr->kind == PHB200_MAT_CSR && !csr->transposed, "coded-stencil recognition needs a plain CSR matrix");
    const int64_t n = csr->csr.n_rows;
    if (n != csr->csr.n_cols || n < 1) { set_error("coded stencils need a square, non-empty matrix"); return PHB200_ERR_NOT_STENCIL; }
    // pass 1: distinct offsets, distinct row hashes with their first rows
    CodedScan* h_scan = new CodedScan();
    for (int k = 0; k < kCodedOffSlots; ++k) h_scan->offs[k] = (unsigned long long)kCodedEmptyOff;
    for (int k = 0; k < kCodedHashSlots; ++k) { h_scan->hash[k] = 0ull; h_scan->rep[k] = 0x7fffffff; h_scan->cnt[k] = 0; }
    h_scan->fail = 0;
    CodedScan* d_scan = nullptr;
    cudaError_t ce = phb_malloc_async(&d_scan, sizeof(CodedScan), stream);
    if (ce != cudaSuccess) { delete h_scan; set_error("allocation failed: %s", cudaGetErrorString(ce)); return PHB200_ERR_CUDA; }
    struct Guard {
        CodedScan* h; void* d[4]; cudaStream_t st;
        ~Guard() { delete h; for (void* p : d) if (p) cudaFreeAsync(p, st); }
    } guard{h_scan, {d_scan, nullptr, nullptr, nullptr}, stream};
    PHB_CUDA(cudaMemcpyAsync(d_scan, h_scan, sizeof(CodedScan), cudaMemcpyHostToDevice, stream));
    const unsigned gridx = (unsigned)grid_x((n + kBlock - 1) / kBlock, 1, 4);
    if (csr->dtype == PHB200_F32) PHB_LAUNCH(k_coded_scan<float>, gridx, kBlock, 0, stream, csr->csr, d_scan);
    else PHB_LAUNCH(k_coded_scan<double>, gridx, kBlock, 0, stream, csr->csr, d_scan);
    PHB_CUDA(cudaMemcpyAsync(h_scan, d_scan, sizeof(CodedScan), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    if (h_scan->fail) { set_error("rows with unsorted columns, more than %d entries, or too many distinct offsets / rows", kMaxStencil); return PHB200_ERR_NOT_STENCIL; }
    std::vector<long long> offs;
    for (int k = 0; k < kCodedOffSlots; ++k)
        if (h_scan->offs[k] != (unsigned long long)kCodedEmptyOff) offs.push_back((long long)h_scan->offs[k]);
    std::sort(offs.begin(), offs.end());
    std::vector<std::pair<int, unsigned long long>> pats;      // (first row, hash)
    int best_cnt = -1, best_rep = 0;
    for (int k = 0; k < kCodedHashSlots; ++k)
        if (h_scan->hash[k] != 0ull) {
            pats.emplace_back(h_scan->rep[k], h_scan->hash[k]);
            if (h_scan->cnt[k] > best_cnt || (h_scan->cnt[k] == best_cnt && h_scan->rep[k] < best_rep)) { best_cnt = h_scan->cnt[k]; best_rep = h_scan->rep[k]; }
        }
    if (offs.empty() || (int)offs.size() > kMaxStencil || pats.empty() || (int)pats.size() > kCodedMaxPat) {
        set_error("%d distinct offsets / %d distinct rows: not a coded stencil (limits %d / %d)", (int)offs.size(), (int)pats.size(), kMaxStencil, kCodedMaxPat);
        return PHB200_ERR_NOT_STENCIL;
    }
    std::sort(pats.begin(), pats.end());
    for (size_t k = 0; k < pats.size(); ++k)                   // pattern 0 = the most frequent one (the interior stencil)
        if (pats[k].first == best_rep) { std::rotate(pats.begin(), pats.begin() + k, pats.begin() + k + 1); break; }
    CodedDict h_dict;
    memset(&h_dict, 0, sizeof(h_dict));
    h_dict.K = (int)offs.size();
    h_dict.npat = (int)pats.size();
    for (int k = 0; k < h_dict.K; ++k) h_dict.off[k] = offs[k];
    for (int k = 0; k < h_dict.npat; ++k) { h_dict.rep[k] = pats[k].first; h_dict.hash[k] = pats[k].second; }
    // pass 2 + 3: table of the distinct rows, code byte per row with bit-for-bit verification
    const size_t w = csr->dtype == PHB200_F32 ? 4 : 8;
    CodedDict* d_dict = nullptr;
    int* d_fail = nullptr;
    PHB_CUDA(phb_malloc_async(&d_dict, sizeof(CodedDict), stream));
    guard.d[1] = d_dict;
    PHB_CUDA(phb_malloc_async(&d_fail, sizeof(int), stream));
    guard.d[2] = d_fail;
    void* d_table = nullptr;
    uint8_t* d_code = nullptr;
    const size_t code_bytes = (size_t)((n + 3) / 4 * 4 + 16);
    PHB_CUDA(cudaMalloc(&d_table, (size_t)h_dict.npat * kMaxStencil * w));
    if (cudaMalloc((void**)&d_code, code_bytes) != cudaSuccess) { cudaFree(d_table); set_error("allocation of %zu code bytes failed", code_bytes); return PHB200_ERR_CUDA; }
    auto fail_free = [&](int rc) { cudaFree(d_table); cudaFree(d_code); return rc; };
    if (cudaMemcpyAsync(d_dict, &h_dict, sizeof(CodedDict), cudaMemcpyHostToDevice, stream) != cudaSuccess || cudaMemsetAsync(d_fail, 0, sizeof(int), stream) != cudaSuccess ||
        cudaMemsetAsync(d_code, 0, code_bytes, stream) != cudaSuccess) { set_error("set-up copies failed"); return fail_free(PHB200_ERR_CUDA); }
    if (csr->dtype == PHB200_F32) {
        k_coded_table<float><<<(h_dict.npat + 63) / 64, 64, 0, stream>>>(csr->csr, d_dict, (float*)d_table);
        k_coded_assign<float><<<gridx, kBlock, 0, stream>>>(csr->csr, d_dict, (const float*)d_table, d_code, d_fail);
    } else {
        k_coded_table<double><<<(h_dict.npat + 63) / 64, 64, 0, stream>>>(csr->csr, d_dict, (double*)d_table);
        k_coded_assign<double><<<gridx, kBlock, 0, stream>>>(csr->csr, d_dict, (const double*)d_table, d_code, d_fail);
    }
    g_launches.fetch_add(2);
    int h_fail = 0;
    if (cudaMemcpyAsync(&h_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
        set_error("coded-stencil kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail_free(PHB200_ERR_CUDA);
    }
    if (h_fail) { set_error("row verification failed (hash collision or inconsistent rows): the matrix stays CSR"); return fail_free(PHB200_ERR_NOT_STENCIL); }
    phb200_matrix* m = new phb200_matrix();
    memset(m, 0, sizeof(*m));
    m->kind = PHB200_MAT_CODED;
    m->dtype = csr->dtype;
    m->csr = csr->csr;
    m->n_rows = m->n_cols = n;
    m->nnz = csr->nnz;
    m->coded.K = h_dict.K;
    m->coded.npat = h_dict.npat;
    m->coded.k_m1 = m->coded.k_p1 = -1;
    for (int k = 0; k < h_dict.K; ++k) {
        m->coded.off[k] = offs[k];
        if (offs[k] == -1) m->coded.k_m1 = k;
        if (offs[k] == 1) m->coded.k_p1 = k;
    }
    // shape of the interior pattern (row 0 of the table): [na aligned offsets, -1, 0, +1, na aligned offsets] selects the specialised product
    {
        double c0[kMaxStencil];
        if (csr->dtype == PHB200_F32) {
            float t32[kMaxStencil];
            if (cudaMemcpy(t32, d_table, sizeof(t32), cudaMemcpyDeviceToHost) != cudaSuccess) { delete m; set_error("table copy failed"); return fail_free(PHB200_ERR_CUDA); }
            for (int k = 0; k < kMaxStencil; ++k) c0[k] = t32[k];
        } else if (cudaMemcpy(c0, d_table, sizeof(c0), cudaMemcpyDeviceToHost) != cudaSuccess) { delete m; set_error("table copy failed"); return fail_free(PHB200_ERR_CUDA); }
        const long long va = 16 / (long long)w;
        std::vector<long long> nzo;
        for (int k = 0; k < h_dict.K; ++k)
            if (c0[k] != 0.0) nzo.push_back(offs[k]);
        m->coded.fast_na = -1;
        const int cnt = (int)nzo.size();
        if (cnt >= 3 && (cnt - 3) % 2 == 0) {
            const int na = (cnt - 3) / 2;
            bool ok = nzo[na] == -1 && nzo[na + 1] == 0 && nzo[na + 2] == 1;
            for (int u = 0; u < na && ok; ++u) ok = nzo[u] % va == 0 && nzo[na + 3 + u] % va == 0;
            if (ok) m->coded.fast_na = na;
        }
    }
    m->coded.code = d_code;
    m->coded.table = d_table;
    m->owned[0] = d_code;
    m->owned[1] = d_table;
    // star with a per-row diagonal (ZERO_GRADIENT-type boundaries)?  Then the fused TMA CG kernels can run it (csrc/star_tma.cuh, DIAG form).
    m->coded.star_diag = 0;
    const int na = m->coded.fast_na;
    if ((na == 1 || na == 2) && !getenv("PHB200_CODED_STAR_OFF")) {
        // interior pattern [(-S2,) -S1, -1, 0, 1, S1 (, S2)]: S1 = row stride (nz), S2 = plane stride (ny nz)
        std::vector<long long> nzo;
        {
            double c0t[kMaxStencil];
            bool got = true;
            if (csr->dtype == PHB200_F32) {
                float t32[kMaxStencil];
                got = cudaMemcpy(t32, d_table, sizeof(t32), cudaMemcpyDeviceToHost) == cudaSuccess;
                for (int k = 0; k < kMaxStencil; ++k) c0t[k] = t32[k];
            } else got = cudaMemcpy(c0t, d_table, sizeof(c0t), cudaMemcpyDeviceToHost) == cudaSuccess;
            if (got)
                for (int k = 0; k < h_dict.K; ++k)
                    if (c0t[k] != 0.0) nzo.push_back(offs[k]);
        }
        auto index_of = [&](long long d) { for (int k = 0; k < h_dict.K; ++k) if (offs[k] == d) return k; return -1; };
        const bool shape_ok = (int)nzo.size() == 2 * na + 3;
        const long long s1 = shape_ok ? nzo[na + 3] : 0, s2 = (shape_ok && na == 2) ? nzo[na + 4] : 0;
        const long long nz = s1, ny = na == 2 ? (s1 > 0 ? s2 / s1 : 0) : n / (s1 > 0 ? s1 : 1), nx = na == 2 ? (s2 > 0 ? n / s2 : 0) : 1;
        const bool mirrored = shape_ok && nzo[na - 1] == -s1 && (na == 1 || nzo[0] == -s2);
        StarCheck sc;
        memset(&sc, 0, sizeof(sc));
        int wrap = 0, accounted = 2 * na + 3;
        if (mirrored && nz >= 4 && ny >= 1 && nx >= 1 && nx * ny * nz == n && (na == 1 || s2 == ny * nz) && nz % 4 == 0) {
            sc.nx = nx; sc.ny = ny; sc.nz = nz;
            sc.k_diag = index_of(0);
            const long long stride[3] = {s2, s1, 1}, ext[3] = {nx, ny, nz};
            for (int ax = 0; ax < 3; ++ax) {
                const bool present = ax > 0 || na == 2;
                sc.k_dir[2 * ax] = present ? index_of(-stride[ax]) : -1;
                sc.k_dir[2 * ax + 1] = present ? index_of(stride[ax]) : -1;
                sc.k_wrap[2 * ax] = sc.k_wrap[2 * ax + 1] = -1;
                if (present && ext[ax] >= 3) {
                    const int kw_m = index_of((ext[ax] - 1) * stride[ax]), kw_p = index_of(-(ext[ax] - 1) * stride[ax]);   // x- wraps to +(n-1) stride, x+ to -(n-1) stride
                    if (kw_m >= 0 && kw_p >= 0) { sc.k_wrap[2 * ax] = kw_m; sc.k_wrap[2 * ax + 1] = kw_p; wrap |= 1 << ax; accounted += 2; }
                }
            }
        }
        if (sc.nz > 0 && accounted == h_dict.K && sc.k_diag >= 0) {
            int* d_sf = nullptr;
            void* d_diag = nullptr;
            void* d_inv = nullptr;
            bool ok = phb_malloc_async(&d_sf, sizeof(int), stream) == cudaSuccess && cudaMemsetAsync(d_sf, 0, sizeof(int), stream) == cudaSuccess &&
                      cudaMalloc(&d_diag, (size_t)h_dict.npat * w) == cudaSuccess && cudaMalloc(&d_inv, (size_t)h_dict.npat * w) == cudaSuccess;
            const int kd = sc.k_diag, kzm = sc.k_dir[4], kzp = sc.k_dir[5], kym = sc.k_dir[2], kyp = sc.k_dir[3], kxm = sc.k_dir[0], kxp = sc.k_dir[1];
            if (ok) {
                if (csr->dtype == PHB200_F32) {
                    k_coded_star_check<float><<<gridx, kBlock, 0, stream>>>(m->coded, n, sc, d_sf);
                    k_coded_diag_tables<float><<<(h_dict.npat + 63) / 64, 64, 0, stream>>>(m->coded, kd, (float*)d_diag, (float*)d_inv);
                } else {
                    k_coded_star_check<double><<<gridx, kBlock, 0, stream>>>(m->coded, n, sc, d_sf);
                    k_coded_diag_tables<double><<<(h_dict.npat + 63) / 64, 64, 0, stream>>>(m->coded, kd, (double*)d_diag, (double*)d_inv);
                }
                g_launches.fetch_add(2);
                int h_sf = 1;
                ok = cudaMemcpyAsync(&h_sf, d_sf, sizeof(int), cudaMemcpyDeviceToHost, stream) == cudaSuccess && cudaStreamSynchronize(stream) == cudaSuccess && h_sf == 0;
            }
            if (d_sf) cudaFreeAsync(d_sf, stream);
            if (ok) {
                double c0[kMaxStencil];
                bool got = true;
                if (csr->dtype == PHB200_F32) {
                    float t32[kMaxStencil];
                    got = cudaMemcpy(t32, d_table, sizeof(t32), cudaMemcpyDeviceToHost) == cudaSuccess;
                    for (int k = 0; k < kMaxStencil; ++k) c0[k] = t32[k];
                } else got = cudaMemcpy(c0, d_table, sizeof(c0), cudaMemcpyDeviceToHost) == cudaSuccess;
                ok = got;
                if (ok) {
                    StarDesc& sd = m->star;
                    sd.nx = nx; sd.ny = ny; sd.nz = nz; sd.xb = 0; sd.xe = nx; sd.store_halo = 0;
                    sd.c0 = c0[kd]; sd.czm = c0[kzm]; sd.czp = c0[kzp]; sd.cym = c0[kym]; sd.cyp = c0[kyp];
                    sd.cxm = kxm >= 0 ? c0[kxm] : 0.0; sd.cxp = kxp >= 0 ? c0[kxp] : 0.0;
                    sd.wrap = wrap;
                    m->is_star = 1;
                    m->coded.star_diag = 1;
                    m->coded.diag_tab = d_diag;
                    m->coded.inv_tab = d_inv;
                    m->owned[2] = d_diag;
                    m->owned[3] = d_inv;
                }
            }
            if (!ok) { if (d_diag) cudaFree(d_diag); if (d_inv) cudaFree(d_inv); cudaGetLastError(); }
        }
    }
    *out = m;
    return PHB200_OK;
}

int phb200_coded_info(const phb200_matrix* m, int* n_offsets, int* n_patterns, int64_t* offsets, int* star_diag) {
    PHB_ARG(m && m->kind == PHB200_MAT_CODED, "need a coded matrix");
    if (star_diag) *star_diag = m->coded.star_diag;
    if (n_offsets) *n_offsets = m->coded.K;
    if (n_patterns) *n_patterns = m->coded.npat;
    if (offsets)
        for (int k = 0; k < m->coded.K; ++k) offsets[k] = m->coded.off[k];
    return PHB200_OK;
}

int phb200_stencil_nnz(const phb200_matrix* m, int64_t* nnz) {
    PHB_ARG(m && nnz && m->kind == PHB200_MAT_STENCIL, "need a stencil matrix");
    const StencilDesc& st = m->st;
    int64_t total = 0;
    for (int p = 0; p < st.npts; ++p) {
        int64_t cnt = 1;
        for (int d = 0; d < 3; ++d) {
            int64_t a = st.off[p][d] < 0 ? -st.off[p][d] : st.off[p][d];
            cnt *= (st.dims[d] - a > 0 ? st.dims[d] - a : 0);
        }
        total += cnt;
    }
    *nnz = total;
    return PHB200_OK;
}

int phb200_stencil_to_csr(const phb200_matrix* m, int32_t* rowptr, int32_t* col, void* val, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(m && rowptr && col && val && m->kind == PHB200_MAT_STENCIL, "need a stencil matrix and output arrays");
    int64_t nnz = 0;
    PHB_TRY(phb200_stencil_nnz(m, &nnz));
    PHB_ARG(nnz < (1ll << 31) && m->n_rows < (1ll << 31) - 1, "CSR with %lld entries does not fit int32 indices", (long long)nnz);
    const int64_t n = m->n_rows;
    const unsigned gridx = (unsigned)grid_x((n + kBlock) / kBlock, 1);
    int32_t* counts = nullptr;
    PHB_CUDA(phb_malloc_async(&counts, (n + 1) * sizeof(int32_t), stream));
    PHB_LAUNCH(k_stencil_count, gridx, kBlock, 0, stream, m->st, n, counts);
    size_t tmp_bytes = 0;
    PHB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, rowptr, (int)(n + 1), stream));
    void* tmp = nullptr;
    PHB_CUDA(phb_malloc_async(&tmp, tmp_bytes, stream));
    PHB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, rowptr, (int)(n + 1), stream));
    g_launches.fetch_add(2);
    if (m->dtype == PHB200_F32) PHB_LAUNCH(k_stencil_fill<float>, gridx, kBlock, 0, stream, m->st, n, rowptr, col, (float*)val);
    else PHB_LAUNCH(k_stencil_fill<double>, gridx, kBlock, 0, stream, m->st, n, rowptr, col, (double*)val);
    PHB_CUDA(cudaFreeAsync(tmp, stream));
    PHB_CUDA(cudaFreeAsync(counts, stream));
    return PHB200_OK;
}

static int shift_set(ShiftSet& sh, int rank, const int64_t* grid, int nshifts, const int32_t* shifts, int64_t* n_out) {
    PHB_ARG(grid && shifts, "null argument");
    PHB_ARG(rank >= 1 && rank <= 3, "rank must be 1..3, got %d", rank);
    PHB_ARG(nshifts >= 1 && nshifts <= kMaxStencil, "1..%d shifts, got %d", kMaxStencil, nshifts);
    memset(&sh, 0, sizeof(sh));
    sh.rank = rank;
    sh.nshifts = nshifts;
    int64_t n = 1;
    for (int d = 0; d < 3; ++d) sh.dims[d] = 1;
    for (int d = 0; d < rank; ++d) {
        PHB_ARG(grid[d] >= 1, "grid size must be positive");
        sh.dims[3 - rank + d] = grid[d];
        n *= grid[d];
    }
    for (int s = 0; s < nshifts; ++s)
        for (int d = 0; d < rank; ++d) sh.off[s][3 - rank + d] = shifts[s * rank + d];
    PHB_ARG(n < (1ll << 31) - 1, "CSR natives hold at most 2^31-1 rows (int32)");
    *n_out = n;
    return PHB200_OK;
}

int phb200_shift_count(int rank, const int64_t* grid, int nshifts, const int32_t* shifts, const void* values, int dtype, int32_t* rowptr, int64_t* nnz, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(values && rowptr && nnz, "null argument");
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    ShiftSet sh;
    int64_t n = 0;
    PHB_TRY(shift_set(sh, rank, grid, nshifts, shifts, &n));
    const unsigned gridx = (unsigned)grid_x((n + kBlock) / kBlock, 1);
    int32_t* counts = nullptr;
    PHB_CUDA(phb_malloc_async(&counts, (n + 1) * sizeof(int32_t), stream));
    if (dtype == PHB200_F32) PHB_LAUNCH(k_shift_count<float>, gridx, kBlock, 0, stream, sh, n, (const float*)values, counts);
    else PHB_LAUNCH(k_shift_count<double>, gridx, kBlock, 0, stream, sh, n, (const double*)values, counts);
    // the total may exceed int32 (then no CSR native exists): scan in 64 bits, narrow afterwards
    int64_t* wide = nullptr;
    PHB_CUDA(phb_malloc_async(&wide, (n + 1) * sizeof(int64_t), stream));
    size_t tmp_bytes = 0;
    PHB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, wide, (int)(n + 1), stream));
    void* tmp = nullptr;
    PHB_CUDA(phb_malloc_async(&tmp, tmp_bytes, stream));
    PHB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, wide, (int)(n + 1), stream));
    g_launches.fetch_add(2);
    int64_t total = 0;
    PHB_CUDA(cudaMemcpyAsync(&total, wide + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    *nnz = total;
    if (total < (1ll << 31)) PHB_LAUNCH(k_narrow_i64, gridx, kBlock, 0, stream, wide, rowptr, n + 1);
    PHB_CUDA(cudaFreeAsync(tmp, stream));
    PHB_CUDA(cudaFreeAsync(wide, stream));
    PHB_CUDA(cudaFreeAsync(counts, stream));
    PHB_ARG(total < (1ll << 31), "CSR with %lld entries does not fit int32 indices", (long long)total);
    return PHB200_OK;
}

int phb200_shift_fill(int rank, const int64_t* grid, int nshifts, const int32_t* shifts, const void* values, int dtype, const int32_t* rowptr, int32_t* col, void* val, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(values && rowptr && col && val, "null argument");
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    ShiftSet sh;
    int64_t n = 0;
    PHB_TRY(shift_set(sh, rank, grid, nshifts, shifts, &n));
    const unsigned gridx = (unsigned)grid_x((n + kBlock - 1) / kBlock, 1);
    int* d_fail = nullptr;
    PHB_CUDA(phb_malloc_async(&d_fail, sizeof(int), stream));
    PHB_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), stream));
    if (dtype == PHB200_F32) PHB_LAUNCH(k_shift_fill<float>, gridx, kBlock, 0, stream, sh, n, (const float*)values, rowptr, col, (float*)val, d_fail);
    else PHB_LAUNCH(k_shift_fill<double>, gridx, kBlock, 0, stream, sh, n, (const double*)values, rowptr, col, (double*)val, d_fail);
    int h = 0;
    PHB_CUDA(cudaMemcpyAsync(&h, d_fail, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PHB_CUDA(cudaStreamSynchronize(stream));
    PHB_CUDA(cudaFreeAsync(d_fail, stream));
    if (h) { set_error("two shifts wrap onto the same column (grid smaller than the stencil)"); return PHB200_ERR_UNSUPPORTED; }
    return PHB200_OK;
}

int phb200_matrix_set_slab(phb200_matrix* m, int64_t x_begin, int64_t x_end) {
    PHB_ARG(m && m->kind == PHB200_MAT_STENCIL && m->is_star, "slab ranges need a star stencil with nz % 4 == 0");
    PHB_ARG(x_begin >= 0 && x_begin < x_end && x_end <= m->star.nx, "bad slab range [%lld, %lld) for %lld planes", (long long)x_begin, (long long)x_end, (long long)m->star.nx);
    m->star.xb = x_begin;
    m->star.xe = x_end;
    m->star.store_halo = (x_begin > 0 || x_end < m->star.nx) ? 1 : 0;
    return PHB200_OK;
}

int phb200_matrix_info(const phb200_matrix* m, int* kind, int64_t* n_rows, int64_t* n_cols, int64_t* nnz, int* dtype) {
    PHB_ARG(m, "null matrix");
    if (kind) *kind = m->kind;
    if (n_rows) *n_rows = m->n_rows;
    if (n_cols) *n_cols = m->n_cols;
    if (dtype) *dtype = m->dtype;
    if (nnz) {
        if (m->kind == PHB200_MAT_STENCIL) PHB_TRY(phb200_stencil_nnz(m, nnz));
        else *nnz = m->nnz;
    }
    return PHB200_OK;
}

int phb200_matrix_destroy(phb200_matrix* m) {
    if (m)
        for (void* p : m->owned)
            if (p) cudaFree(p);
    delete m;
    return PHB200_OK;
}

int phb200_spmv(const phb200_matrix* A, const void* X, void* Y, int64_t batch, int64_t ldx, int64_t ldy, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(A && X && Y, "null argument");
    PHB_ARG(batch >= 0 && batch <= 65535, "batch must be <= 65535, got %lld", (long long)batch);
    PHB_ARG(ldx >= A->n_cols && ldy >= A->n_rows, "leading dimensions too small");
    if (batch == 0 || A->n_rows == 0) return PHB200_OK;
    if (A->kind == PHB200_MAT_CSR && A->transposed) {
        const size_t w = A->dtype == PHB200_F32 ? 4 : 8;
        PHB_CUDA(cudaMemset2DAsync(Y, ldy * w, 0, A->n_rows * w, batch, stream));
        dim3 grid((unsigned)grid_x((A->csr.n_rows + kBlock - 1) / kBlock, batch), (unsigned)batch);
        if (A->dtype == PHB200_F32) PHB_LAUNCH(k_spmv_csr_scatter<float>, grid, kBlock, 0, stream, A->csr, (const float*)X, ldx, (float*)Y, ldy);
        else PHB_LAUNCH(k_spmv_csr_scatter<double>, grid, kBlock, 0, stream, A->csr, (const double*)X, ldx, (double*)Y, ldy);
        return PHB200_OK;
    }
    if (A->dtype == PHB200_F32) {
        NoEpi<float> epi{};
        return launch_spmv<float>(A, (const float*)X, ldx, (float*)Y, ldy, (const float*)nullptr, (int)batch, epi, stream);
    }
    NoEpi<double> epi{};
    return launch_spmv<double>(A, (const double*)X, ldx, (double*)Y, ldy, (const double*)nullptr, (int)batch, epi, stream);
}

int phb200_spmv_coo(const int64_t* row, const int64_t* col, const void* val, int64_t nnz, int64_t n_rows, int64_t n_cols,
                    const void* X, void* Y, int64_t batch, int64_t ldx, int64_t ldy, int dtype, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG((nnz == 0 || (row && col && val)) && X && Y, "null argument");
    PHB_ARG(batch >= 0 && batch <= 65535, "batch must be <= 65535");
    PHB_ARG(ldx >= n_cols && ldy >= n_rows, "leading dimensions too small");
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    if (batch == 0 || nnz == 0) return PHB200_OK;
    dim3 grid((unsigned)grid_x((nnz + kBlock - 1) / kBlock, batch), (unsigned)batch);
    if (dtype == PHB200_F32) PHB_LAUNCH(k_spmv_coo<float>, grid, kBlock, 0, stream, row, col, (const float*)val, nnz, (const float*)X, ldx, (float*)Y, ldy);
    else PHB_LAUNCH(k_spmv_coo<double>, grid, kBlock, 0, stream, row, col, (const double*)val, nnz, (const double*)X, ldx, (double*)Y, ldy);
    return PHB200_OK;
}

int phb200_csr_dense_batched(const int32_t* col, const int32_t* rowptr, const void* val, const void* dense, void* out, int64_t batch, int64_t n_rows,
                             int64_t n_cols, int64_t nnz, int64_t channels, int64_t dense_cols, int dtype, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(batch >= 0 && n_rows >= 0 && n_cols >= 0 && nnz >= 0 && channels >= 0 && dense_cols >= 0, "negative size");
    PHB_ARG(nnz <= 2147483647LL && channels <= 2147483647LL && dense_cols <= 2147483647LL, "CSR natives hold at most 2^31-1 entries (int32 pointers)");
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    const int64_t total = batch * n_rows * channels * dense_cols;
    if (total == 0) return PHB200_OK;
    PHB_ARG(rowptr && out && (nnz == 0 || (col && val && dense)), "null argument");
    dim3 grid((unsigned)grid_x((total + kBlock - 1) / kBlock, 1));
    if (dtype == PHB200_F32) PHB_LAUNCH(k_csr_dense_batched<float>, grid, kBlock, 0, stream, col, rowptr, (const float*)val, (const float*)dense, (float*)out, batch, n_rows, n_cols, nnz, (int)channels, (int)dense_cols);
    else PHB_LAUNCH(k_csr_dense_batched<double>, grid, kBlock, 0, stream, col, rowptr, (const double*)val, (const double*)dense, (double*)out, batch, n_rows, n_cols, nnz, (int)channels, (int)dense_cols);
    return PHB200_OK;
}

int phb200_matrix_gradient_coo(const int64_t* row, const int64_t* col, int64_t nnz, int64_t n_rows, int64_t n_cols, const void* DY, const void* X,
                               int64_t batch, int64_t ld_dy, int64_t ld_x, void* out, int dtype, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(nnz == 0 || (row && col && DY && X && out), "null argument");
    PHB_ARG(batch >= 0 && batch <= 65535, "batch must be <= 65535");
    PHB_ARG(ld_dy >= n_rows && ld_x >= n_cols, "leading dimensions too small");
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    if (nnz == 0) return PHB200_OK;
    dim3 grid((unsigned)grid_x((nnz + kBlock - 1) / kBlock, 1));
    if (dtype == PHB200_F32) PHB_LAUNCH(k_matrix_gradient_coo<float>, grid, kBlock, 0, stream, row, col, nnz, (const float*)DY, ld_dy, (const float*)X, ld_x, (int)batch, (float*)out);
    else PHB_LAUNCH(k_matrix_gradient_coo<double>, grid, kBlock, 0, stream, row, col, nnz, (const double*)DY, ld_dy, (const double*)X, ld_x, (int)batch, (double*)out);
    return PHB200_OK;
}

int phb200_matrix_gradient_csr(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_cols, int64_t nnz, const void* DY, const void* X,
                               int64_t batch, int64_t ld_dy, int64_t ld_x, void* out, int dtype, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PHB_ARG(nnz == 0 || (rowptr && col && DY && X && out), "null argument");
    PHB_ARG(nnz <= 2147483647LL, "CSR natives hold at most 2^31-1 entries (int32 pointers)");
    PHB_ARG(batch >= 0 && batch <= 65535, "batch must be <= 65535");
    PHB_ARG(ld_dy >= n_rows && ld_x >= n_cols, "leading dimensions too small");
    PHB_ARG(dtype == PHB200_F32 || dtype == PHB200_F64, "bad dtype %d", dtype);
    if (nnz == 0 || n_rows == 0) return PHB200_OK;
    dim3 grid((unsigned)grid_x((n_rows + kBlock - 1) / kBlock, 1));
    if (dtype == PHB200_F32) PHB_LAUNCH(k_matrix_gradient_csr<float>, grid, kBlock, 0, stream, rowptr, col, n_rows, (const float*)DY, ld_dy, (const float*)X, ld_x, (int)batch, (float*)out);
    else PHB_LAUNCH(k_matrix_gradient_csr<double>, grid, kBlock, 0, stream, rowptr, col, n_rows, (const double*)DY, ld_dy, (const double*)X, ld_x, (int)batch, (double*)out);
    return PHB200_OK;
}

}  // extern "C"
