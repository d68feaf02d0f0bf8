// matrix.cu -- handle lifetime, derived compact arrays, scans, export/compare.
// Reference surface: include/bmSpMatrix.h:20-40, src/bmSpMatrix.cu:30-43 (adopting constructor),
// :320-363 (generate_coo), :381-432 (compare).
#include "common.cuh"
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace bmsp {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return BMSP_ERR_CUDA;
}

static bool g_pool_ready = false;
int dev_alloc(void** p, size_t bytes, cudaStream_t st) {
    if (!g_pool_ready) {
        int dev = 0;
        BMSP_CUDA(cudaGetDevice(&dev));
        cudaMemPool_t pool;
        BMSP_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t thr = UINT64_MAX;   // keep freed blocks cached: repeated products reuse them without cudaMalloc
        BMSP_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        g_pool_ready = true;
    }
    *p = nullptr;
    BMSP_CUDA(cudaMallocAsync(p, bytes + BMSP_PAD_BYTES, st));
    return BMSP_OK;
}
void dev_free(void* p, cudaStream_t st) {
    if (p) cudaFreeAsync(p, st);
}

// ------------------------------------------------------------------------------------ scans
// Three-kernel scan: per-tile sums -> (recursive) scan of sums -> per-tile scan with base.
template <typename T>
__global__ void scan_tile_sums(const T* __restrict__ in, T* __restrict__ sums, int64_t n) {
    constexpr int ITEMS = 8;
    int64_t base = (int64_t)blockIdx.x * (blockDim.x * ITEMS);
    T s = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        int64_t idx = base + (int64_t)i * blockDim.x + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    __shared__ T warp_s[32];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_s[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < (blockDim.x >> 5) ? warp_s[threadIdx.x] : 0;
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = v;
    }
}

template <typename T>
__global__ void scan_tiles(const T* in, T* out, const T* __restrict__ tile_base, int64_t n,
                           int write_total) {
    constexpr int ITEMS = 8;
    // blocked arrangement: thread t owns ITEMS consecutive elements
    int64_t base = (int64_t)blockIdx.x * (blockDim.x * ITEMS) + (int64_t)threadIdx.x * ITEMS;
    T v[ITEMS];
    T s = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        int64_t idx = base + i;
        v[i] = idx < n ? in[idx] : 0;
        s += v[i];
    }
    // exclusive scan of s across the block
    T incl = s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __shared__ T warp_s[32];
    if (lane == 31) warp_s[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        T w = lane < (blockDim.x >> 5) ? warp_s[lane] : 0;
        T wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_s[lane] = wi - w;
    }
    __syncthreads();
    T run = (tile_base ? tile_base[blockIdx.x] : 0) + warp_s[wid] + (incl - s);
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        int64_t idx = base + i;
        if (idx < n) out[idx] = run;
        run += v[i];
        if (write_total && idx == n - 1) out[n] = run;
    }
}

template <typename T>
static int exclusive_scan_impl(const T* in, T* out, int64_t n, cudaStream_t st, int write_total) {
    constexpr int THREADS = 256, TILE = THREADS * 8;
    if (n <= 0) {
        if (write_total) BMSP_CUDA(cudaMemsetAsync(out, 0, sizeof(T), st));
        return BMSP_OK;
    }
    int64_t tiles = ceil_div(n, TILE);
    if (tiles == 1) {
        scan_tiles<T><<<1, THREADS, 0, st>>>(in, out, nullptr, n, write_total);
        BMSP_KERNEL_CHECK();
        return BMSP_OK;
    }
    T* sums = nullptr;
    BMSP_TRY(dev_alloc_t(&sums, (size_t)tiles + 1, st));
    scan_tile_sums<T><<<(unsigned)tiles, THREADS, 0, st>>>(in, sums, n);
    BMSP_KERNEL_CHECK();
    int s = exclusive_scan_impl<T>(sums, sums, tiles, st, 0);
    if (s != BMSP_OK) { dev_free(sums, st); return s; }
    scan_tiles<T><<<(unsigned)tiles, THREADS, 0, st>>>(in, out, sums, n, write_total);
    BMSP_KERNEL_CHECK();
    dev_free(sums, st);
    return BMSP_OK;
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t st) {
    return exclusive_scan_impl<uint32_t>(in, out, n, st, 1);
}
int exclusive_scan_u64(const uint64_t* in, uint64_t* out, int64_t n, cudaStream_t st) {
    return exclusive_scan_impl<uint64_t>(in, out, n, st, 1);
}

// ------------------------------------------------------------------------------------ derive
__global__ void derive_blocks_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ bmps,
                                     int32_t* __restrict__ bcol, uint8_t* __restrict__ kmask, int64_t nblk) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    bcol[b] = (int32_t)(keys[b] & 0xFFFFFFFFull);
    kmask[b] = (uint8_t)kmask_of(bmps[b]);
}

// brp[br] = first block whose block row >= br (lower bound on keys); rvb[br] = its value offset.
__global__ void derive_rows_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ offsets,
                                   int32_t* __restrict__ brp, uint32_t* __restrict__ rvb, int64_t nblk, int32_t nbr,
                                   uint64_t nnz) {
    int32_t br = blockIdx.x * blockDim.x + threadIdx.x;
    if (br > nbr) return;
    uint64_t target = (uint64_t)br << 32;
    int64_t lo = 0, hi = nblk;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    brp[br] = (int32_t)lo;
    rvb[br] = (uint32_t)(lo < nblk ? offsets[lo] : nnz);
}

int derive_compact(bmsp_matrix_s* m, cudaStream_t st) {
    if (m->nblk > 0x7FFFFFFFll || m->nnz > 0xFFFFFFFFll) {
        set_error("matrix too large: %lld blocks, %lld values", (long long)m->nblk, (long long)m->nnz);
        return BMSP_ERR_TOO_LARGE;
    }
    m->nbr = (int32_t)ceil_div(m->rows, 8);
    BMSP_TRY(dev_alloc_t(&m->bcol, (size_t)m->nblk + 8, st));
    BMSP_TRY(dev_alloc_t(&m->kmask, (size_t)m->nblk + 16, st));
    BMSP_TRY(dev_alloc_t(&m->brp, (size_t)m->nbr + 1 + 8, st));
    BMSP_TRY(dev_alloc_t(&m->rvb, (size_t)m->nbr + 1 + 8, st));
    if (m->nblk > 0) {
        derive_blocks_kernel<<<(unsigned)ceil_div(m->nblk, 256), 256, 0, st>>>(m->keys, m->bmps, m->bcol, m->kmask, m->nblk);
        BMSP_KERNEL_CHECK();
    }
    derive_rows_kernel<<<(unsigned)ceil_div(m->nbr + 1, 256), 256, 0, st>>>(m->keys, m->offsets, m->brp, m->rvb, m->nblk,
                                                                          m->nbr, (uint64_t)m->nnz);
    BMSP_KERNEL_CHECK();
    m->spmv_path = -2;      // SpMV plan is built lazily by the first bmsp_spmv
    return BMSP_OK;
}

}  // namespace bmsp

using namespace bmsp;

static size_t dtype_size(int dtype) { return dtype == BMSP_F16 ? 2 : 4; }

extern "C" {

int bmsp_abi_version(void) { return BMSP_ABI_VERSION; }
const char* bmsp_last_error(void) { return g_err.c_str(); }

int bmsp_device_info(int32_t* sm_count, int64_t* l2_bytes, int64_t* hbm_bytes, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    BMSP_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    BMSP_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (l2_bytes) *l2_bytes = p.l2CacheSize;
    if (hbm_bytes) *hbm_bytes = (int64_t)p.totalGlobalMem;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return BMSP_OK;
}

int bmsp_create_from_arrays(int32_t rows, int32_t cols, int64_t block_num, int64_t nnz, const uint64_t* keys,
                            const uint64_t* bmps, const uint64_t* offsets, int64_t offsets_len, const void* values,
                            int32_t dtype, int32_t mem, int32_t transposed, void* stream, bmsp_matrix_t* out) {
    if (!out || rows < 0 || cols < 0 || block_num < 0 || nnz < 0 || (dtype != BMSP_F16 && dtype != BMSP_F32) ||
        (offsets_len != block_num && offsets_len != block_num + 1) ||
        (block_num > 0 && (!keys || !bmps || !offsets)) || (nnz > 0 && !values)) {
        set_error("bmsp_create_from_arrays: invalid argument");
        return BMSP_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    bmsp_matrix_s* m = new bmsp_matrix_s();
    touch(m, st);
    m->rows = rows; m->cols = cols; m->nblk = block_num; m->nnz = nnz; m->offsets_len = offsets_len;
    m->dtype = dtype; m->transposed = transposed;
    cudaMemcpyKind kind = mem == BMSP_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    int s = BMSP_OK;
    auto fail = [&](int code) { bmsp_destroy(m); return code; };
    if ((s = dev_alloc_t(&m->keys, (size_t)block_num + 2, st))) return fail(s);
    if ((s = dev_alloc_t(&m->bmps, (size_t)block_num + 2, st))) return fail(s);
    if ((s = dev_alloc_t(&m->offsets, (size_t)block_num + 2, st))) return fail(s);
    if ((s = dev_alloc(&m->values, (size_t)nnz * dtype_size(dtype) + 16, st))) return fail(s);
    if (block_num > 0) {
        if (cudaMemcpyAsync(m->keys, keys, sizeof(uint64_t) * block_num, kind, st) != cudaSuccess ||
            cudaMemcpyAsync(m->bmps, bmps, sizeof(uint64_t) * block_num, kind, st) != cudaSuccess ||
            cudaMemcpyAsync(m->offsets, offsets, sizeof(uint64_t) * offsets_len, kind, st) != cudaSuccess)
            return fail(cuda_fail(cudaGetLastError(), "copy structure", __FILE__, __LINE__));
    }
    if (nnz > 0 && cudaMemcpyAsync(m->values, values, (size_t)nnz * dtype_size(dtype), kind, st) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "copy values", __FILE__, __LINE__));
    if ((s = derive_compact(m, st))) return fail(s);
    if (mem == BMSP_HOST && cudaStreamSynchronize(st) != cudaSuccess)   // host buffers may be freed by the caller
        return fail(cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__));
    *out = m;
    return BMSP_OK;
}

int bmsp_destroy(bmsp_matrix_t m) {
    if (!m) return BMSP_OK;
    // frees are ordered behind the last work enqueued on the arrays (stream 0 does not order after non-blocking streams)
    if (m->multi_stream) cudaDeviceSynchronize();
    cudaStream_t st = m->last_stream;
    spmv_host_release(m);
    dev_free(m->keys, st); dev_free(m->bmps, st); dev_free(m->offsets, st); dev_free(m->values, st);
    dev_free(m->brp, st); dev_free(m->bcol, st); dev_free(m->rvb, st); dev_free(m->kmask, st);
    dev_free(m->work, st); dev_free(m->split_partial, st); dev_free(m->split_rows, st); dev_free(m->split_list, st); dev_free(m->pmeta, st);
    dev_free(m->tile_desc, st); dev_free(m->tile_rowpair, st); dev_free(m->tile_lines, st); dev_free(m->tile_xoff, st);
    delete m;
    return BMSP_OK;
}

int bmsp_get(bmsp_matrix_t m, bmsp_view* v) {
    if (!m || !v) { set_error("bmsp_get: null argument"); return BMSP_ERR_INVALID; }
    v->num_rows = m->rows; v->num_cols = m->cols; v->nnz = m->nnz; v->block_num = m->nblk;
    v->keys = m->keys; v->bmps = m->bmps; v->offsets = m->offsets; v->values = m->values;
    v->offsets_len = m->offsets_len; v->dtype = m->dtype; v->transposed = m->transposed;
    v->num_block_rows = m->nbr; v->block_row_ptr = m->brp; v->block_col = m->bcol; v->block_row_val = m->rvb;
    return BMSP_OK;
}

int bmsp_download(bmsp_matrix_t m, uint64_t* keys, uint64_t* bmps, uint64_t* offsets, void* values) {
    if (!m) { set_error("bmsp_download: null matrix"); return BMSP_ERR_INVALID; }
    BMSP_CUDA(cudaDeviceSynchronize());
    if (keys && m->nblk) BMSP_CUDA(cudaMemcpy(keys, m->keys, sizeof(uint64_t) * m->nblk, cudaMemcpyDeviceToHost));
    if (bmps && m->nblk) BMSP_CUDA(cudaMemcpy(bmps, m->bmps, sizeof(uint64_t) * m->nblk, cudaMemcpyDeviceToHost));
    if (offsets && m->offsets_len) BMSP_CUDA(cudaMemcpy(offsets, m->offsets, sizeof(uint64_t) * m->offsets_len, cudaMemcpyDeviceToHost));
    if (values && m->nnz) BMSP_CUDA(cudaMemcpy(values, m->values, (size_t)m->nnz * dtype_size(m->dtype), cudaMemcpyDeviceToHost));
    return BMSP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------ to_coo / compare
// One thread per block walks its bits MSB-first (bmSpMatrix.cu:334-353) and writes (row, col, fp32 value)
// at offsets[b] + rank.
template <typename T>
__global__ void to_coo_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ bmps,
                              const uint64_t* __restrict__ offsets, const T* __restrict__ values, int transposed,
                              int32_t* __restrict__ rows, int32_t* __restrict__ cols, float* __restrict__ vals, int64_t nblk) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    uint64_t bmp = bmps[b], key = keys[b];
    int32_t br = (int32_t)(key >> 32), bc = (int32_t)(key & 0xFFFFFFFFull);
    uint64_t w = offsets[b];
    while (bmp) {
        int p = __clzll((long long)bmp);
        bmp &= ~(0x8000000000000000ull >> p);
        int hi = p >> 3, lo = p & 7;
        rows[w] = br * 8 + (transposed ? lo : hi);
        cols[w] = bc * 8 + (transposed ? hi : lo);
        vals[w] = val_to_f32(values[w]);
        w++;
    }
}

extern "C" int bmsp_to_coo(bmsp_matrix_t m, int32_t* rows, int32_t* cols, float* vals) {
    if (!m || !rows || !cols || !vals) { set_error("bmsp_to_coo: null argument"); return BMSP_ERR_INVALID; }
    if (m->nnz == 0) return BMSP_OK;
    if (m->multi_stream) BMSP_CUDA(cudaDeviceSynchronize());
    cudaStream_t st = m->last_stream;          // ordered behind whatever produced the arrays
    int32_t *dr = nullptr, *dc = nullptr; float* dv = nullptr;
    BMSP_TRY(dev_alloc_t(&dr, (size_t)m->nnz, st));
    BMSP_TRY(dev_alloc_t(&dc, (size_t)m->nnz, st));
    BMSP_TRY(dev_alloc_t(&dv, (size_t)m->nnz, st));
    unsigned grid = (unsigned)ceil_div(m->nblk, 128);
    if (m->dtype == BMSP_F16)
        to_coo_kernel<__half><<<grid, 128, 0, st>>>(m->keys, m->bmps, m->offsets, (const __half*)m->values, m->transposed, dr, dc, dv, m->nblk);
    else
        to_coo_kernel<float><<<grid, 128, 0, st>>>(m->keys, m->bmps, m->offsets, (const float*)m->values, m->transposed, dr, dc, dv, m->nblk);
    BMSP_KERNEL_CHECK();
    BMSP_CUDA(cudaMemcpyAsync(rows, dr, sizeof(int32_t) * m->nnz, cudaMemcpyDeviceToHost, st));
    BMSP_CUDA(cudaMemcpyAsync(cols, dc, sizeof(int32_t) * m->nnz, cudaMemcpyDeviceToHost, st));
    BMSP_CUDA(cudaMemcpyAsync(vals, dv, sizeof(float) * m->nnz, cudaMemcpyDeviceToHost, st));
    BMSP_CUDA(cudaStreamSynchronize(st));
    dev_free(dr, st); dev_free(dc, st); dev_free(dv, st);
    return BMSP_OK;
}

extern "C" int bmsp_compare(bmsp_matrix_t m, int64_t nnz, const int32_t* rows, const int32_t* cols, const float* vals,
                            int64_t* only_in_m, int64_t* only_in_coo, double* mean_rel_err, double* max_rel_err) {
    if (!m || (nnz > 0 && (!rows || !cols || !vals))) { set_error("bmsp_compare: null argument"); return BMSP_ERR_INVALID; }
    // decode on the device, merge on the host (test utility, like the reference's host walk :398-428)
    std::vector<int32_t> r((size_t)m->nnz), c((size_t)m->nnz);
    std::vector<float> v((size_t)m->nnz);
    if (m->nnz) BMSP_TRY(bmsp_to_coo(m, r.data(), c.data(), v.data()));
    std::vector<int64_t> pa((size_t)m->nnz), pb((size_t)nnz);
    for (int64_t i = 0; i < m->nnz; i++) pa[i] = i;
    for (int64_t i = 0; i < nnz; i++) pb[i] = i;
    std::sort(pa.begin(), pa.end(), [&](int64_t x, int64_t y) { return r[x] != r[y] ? r[x] < r[y] : c[x] < c[y]; });
    std::sort(pb.begin(), pb.end(), [&](int64_t x, int64_t y) { return rows[x] != rows[y] ? rows[x] < rows[y] : cols[x] < cols[y]; });
    int64_t i = 0, j = 0, oa = 0, ob = 0, common = 0;
    double sum = 0, mx = 0;
    const double eps = 1e-8;
    while (i < m->nnz || j < nnz) {
        int cmp;
        if (i >= m->nnz) cmp = 1;
        else if (j >= nnz) cmp = -1;
        else {
            int64_t a = pa[i], b = pb[j];
            cmp = r[a] != rows[b] ? (r[a] < rows[b] ? -1 : 1) : (c[a] != cols[b] ? (c[a] < cols[b] ? -1 : 1) : 0);
        }
        if (cmp < 0) { oa++; i++; }
        else if (cmp > 0) { ob++; j++; }
        else {
            double e = vals[pb[j]], g = v[pa[i]];
            double rel = std::fabs(e - g) / std::max(std::fabs(e), eps);
            sum += rel; mx = std::max(mx, rel); common++;
            i++; j++;
        }
    }
    if (only_in_m) *only_in_m = oa;
    if (only_in_coo) *only_in_coo = ob;
    if (mean_rel_err) *mean_rel_err = common ? sum / common : 0.0;
    if (max_rel_err) *max_rel_err = mx;
    return BMSP_OK;
}
