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
    if (m->halo_side) { cudaStreamSynchronize(m->halo_side); cudaStreamDestroy(m->halo_side); }
    if (m->halo_fork) cudaEventDestroy(m->halo_fork);
    if (m->halo_join) cudaEventDestroy(m->halo_join);
    dev_free(m->keys, st); dev_free(m->bmps, st); dev_free(m->offsets, st); dev_free(m->values, st);
    dev_free(m->brp, st); dev_free(m->bcol, st); dev_free(m->rvb, st); dev_free(m->kmask, st);
    dev_free(m->work, st); dev_free(m->split_partial, st); dev_free(m->split_rows, st); dev_free(m->split_list, st); dev_free(m->pmeta, st);
    dev_free(m->halo_tiles, st); dev_free(m->fine_ptr, st); dev_free(m->fine_bcol, st); dev_free(m->fine_kmask, st); dev_free(m->fine_rec, st);
    dev_free(m->tile_desc, st); dev_free(m->spmv_sched, st); dev_free(m->tile_rowpair, st); dev_free(m->tile_lines, st); dev_free(m->tile_xoff, st);
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

// ------------------------------------------------------------------------------------ bmSparse -> CSR on the device
// The blocks of a block row are already in ascending block column and a bitmap is row-major inside a block, so CSR order is an
// 8-way merge in reverse: entry (r, c) of block b goes to row_ptr[r] + (cells of row r in the earlier blocks of the block row) +
// (cells of row r left of c in this block).  Pass 1 counts the cells per matrix row (one thread per block), a scan gives row_ptr,
// pass 2 runs one warp per block row: lanes <-> blocks, 32 at a time; the eight per-row counts of a block travel as two 64-bit
// words of four 16-bit fields through one warp scan each.
__global__ void csr_count_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ bmps, int64_t nblk, int32_t rows,
                                 uint32_t* __restrict__ row_cnt) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint64_t bmp = bmps[b];
    const int64_t r0 = (int64_t)(keys[b] >> 32) * 8;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t c = __popc((uint32_t)(bmp >> (56 - 8 * r)) & 0xFFu);
        if (c && r0 + r < rows) atomicAdd(row_cnt + r0 + r, c);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) csr_fill_kernel(const int32_t* __restrict__ brp, const int32_t* __restrict__ bcol, const uint64_t* __restrict__ bmps,
                                                       const uint64_t* __restrict__ offsets, const T* __restrict__ values, int32_t nbr, int32_t rows,
                                                       const uint32_t* __restrict__ row_ptr, int32_t* __restrict__ col_idx, float* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int br = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (br >= nbr) return;
    uint32_t base[8];                     // next free slot of each of the block row's eight matrix rows (warp-uniform)
#pragma unroll
    for (int r = 0; r < 8; r++) base[r] = (int64_t)br * 8 + r < rows ? row_ptr[(int64_t)br * 8 + r] : 0u;
    for (int b0 = brp[br]; b0 < brp[br + 1]; b0 += 32) {
        const int b = b0 + lane;
        const bool valid = b < brp[br + 1];
        const uint64_t bmp = valid ? bmps[b] : 0ull;
        uint64_t lo = 0, hi = 0;          // per-row counts, 16-bit fields: rows 0-3 in lo, 4-7 in hi
#pragma unroll
        for (int r = 0; r < 4; r++) {
            lo |= (uint64_t)__popc((uint32_t)(bmp >> (56 - 8 * r)) & 0xFFu) << (16 * r);
            hi |= (uint64_t)__popc((uint32_t)(bmp >> (24 - 8 * r)) & 0xFFu) << (16 * r);
        }
        uint64_t ilo = lo, ihi = hi;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t tl = __shfl_up_sync(0xffffffffu, ilo, o), th = __shfl_up_sync(0xffffffffu, ihi, o);
            if (lane >= o) { ilo += tl; ihi += th; }
        }
        const uint64_t elo = ilo - lo, ehi = ihi - hi;             // exclusive: cells of each row in the earlier blocks of this chunk
        if (valid) {
            uint64_t k = offsets[b];
            const int32_t c0 = bcol[b] * 8;
            uint64_t rem = bmp;
            uint32_t in_row[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            while (rem) {
                const int p = __clzll((long long)rem);
                rem &= ~(0x8000000000000000ull >> p);
                const int r = p >> 3;
                const uint32_t ex = (uint32_t)(((r < 4 ? elo : ehi) >> (16 * (r & 3))) & 0xFFFFu);
                uint32_t slot = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) if (q == r) slot = base[q] + ex + in_row[q]++;
                col_idx[slot] = c0 + (p & 7);
                vals[slot] = val_to_f32(values[k++]);
            }
        }
        const uint64_t tlo = __shfl_sync(0xffffffffu, ilo, 31), thi = __shfl_sync(0xffffffffu, ihi, 31);
#pragma unroll
        for (int r = 0; r < 4; r++) { base[r] += (uint32_t)((tlo >> (16 * r)) & 0xFFFFu); base[4 + r] += (uint32_t)((thi >> (16 * r)) & 0xFFFFu); }
    }
}

// device CSR of m into caller-provided device arrays (row_ptr: rows + 1)
static int to_csr_device(bmsp_matrix_s* m, int32_t* d_rp, int32_t* d_ci, float* d_v, cudaStream_t st) {
    if (m->transposed) { set_error("bmsp_to_csr: matrix is in transposed-operand form (bmsp_block_transpose first)"); return BMSP_ERR_UNSUPPORTED; }
    BMSP_CUDA(cudaMemsetAsync(d_rp, 0, sizeof(int32_t) * ((size_t)m->rows + 1), st));
    if (m->nblk == 0) return BMSP_OK;
    csr_count_kernel<<<(unsigned)ceil_div(m->nblk, 256), 256, 0, st>>>(m->keys, m->bmps, m->nblk, m->rows, (uint32_t*)d_rp);
    BMSP_KERNEL_CHECK();
    BMSP_TRY(exclusive_scan_u32((const uint32_t*)d_rp, (uint32_t*)d_rp, m->rows, st));
    const unsigned grid = (unsigned)ceil_div(m->nbr, 8);
    if (m->dtype == BMSP_F16)
        csr_fill_kernel<__half><<<grid, 256, 0, st>>>(m->brp, m->bcol, m->bmps, m->offsets, (const __half*)m->values, m->nbr, m->rows, (const uint32_t*)d_rp, d_ci, d_v);
    else
        csr_fill_kernel<float><<<grid, 256, 0, st>>>(m->brp, m->bcol, m->bmps, m->offsets, (const float*)m->values, m->nbr, m->rows, (const uint32_t*)d_rp, d_ci, d_v);
    BMSP_KERNEL_CHECK();
    return BMSP_OK;
}

extern "C" int bmsp_to_csr(bmsp_matrix_t m, int32_t* row_ptr, int32_t* col_idx, float* vals, int32_t mem, void* stream) {
    if (!m || !row_ptr || (m->nnz > 0 && (!col_idx || !vals))) { set_error("bmsp_to_csr: null argument"); return BMSP_ERR_INVALID; }
    if (m->nnz > 0x7FFFFFFFll) { set_error("bmsp_to_csr: %lld values exceed int32 row pointers", (long long)m->nnz); return BMSP_ERR_TOO_LARGE; }
    cudaStream_t st = (cudaStream_t)stream;
    touch(m, st);
    if (mem == BMSP_DEVICE) return to_csr_device(m, row_ptr, col_idx, vals, st);
    int32_t *d_rp = nullptr, *d_ci = nullptr; float* d_v = nullptr;
    BMSP_TRY(dev_alloc_t(&d_rp, (size_t)m->rows + 2, st));
    BMSP_TRY(dev_alloc_t(&d_ci, (size_t)m->nnz + 1, st));
    BMSP_TRY(dev_alloc_t(&d_v, (size_t)m->nnz + 1, st));
    int s = to_csr_device(m, d_rp, d_ci, d_v, st);
    if (s == BMSP_OK) {
        cudaError_t e = cudaMemcpyAsync(row_ptr, d_rp, sizeof(int32_t) * ((size_t)m->rows + 1), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && m->nnz) e = cudaMemcpyAsync(col_idx, d_ci, sizeof(int32_t) * (size_t)m->nnz, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && m->nnz) e = cudaMemcpyAsync(vals, d_v, sizeof(float) * (size_t)m->nnz, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) s = cuda_fail(e, "bmsp_to_csr: copy out", __FILE__, __LINE__);
    }
    dev_free(d_rp, st); dev_free(d_ci, st); dev_free(d_v, st);
    return s;
}

// ------------------------------------------------------------------------------------ compare on the device
// Both sides as CSR (columns ascending in every row): one thread per matrix row merges the two column lists -- no sort.  Counts
// and error sums are reduced per CTA, then with one atomic per CTA (doubles: the mean is order-dependent only in the last bits).
__global__ void __launch_bounds__(256) compare_rows_kernel(int32_t rows, const int32_t* __restrict__ a_rp, const int32_t* __restrict__ a_ci,
                                                           const float* __restrict__ a_v, const int32_t* __restrict__ b_rp,
                                                           const int32_t* __restrict__ b_ci, const float* __restrict__ b_v,
                                                           unsigned long long* __restrict__ counts, double* __restrict__ sums) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long only_a = 0, only_b = 0, common = 0;
    double sum = 0.0, mx = 0.0;
    if (r < rows) {
        int i = a_rp[r], j = b_rp[r];
        const int ie = a_rp[r + 1], je = b_rp[r + 1];
        while (i < ie || j < je) {
            const int ca = i < ie ? a_ci[i] : 0x7FFFFFFF, cb = j < je ? b_ci[j] : 0x7FFFFFFF;
            if (ca < cb) { only_a++; i++; }
            else if (cb < ca) { only_b++; j++; }
            else {
                const double e = (double)b_v[j], g = (double)a_v[i];
                const double rel = fabs(e - g) / fmax(fabs(e), 1e-8);       // the reference's metric, bmSpMatrix.cu:418
                sum += rel; mx = fmax(mx, rel); common++;
                i++; j++;
            }
        }
    }
    __shared__ unsigned long long s_c[3];
    __shared__ double s_sum;
    __shared__ unsigned long long s_mx;
    if (threadIdx.x == 0) { s_c[0] = s_c[1] = s_c[2] = 0; s_sum = 0.0; s_mx = 0ull; }
    __syncthreads();
    for (int o = 16; o; o >>= 1) {
        only_a += __shfl_xor_sync(0xffffffffu, only_a, o); only_b += __shfl_xor_sync(0xffffffffu, only_b, o);
        common += __shfl_xor_sync(0xffffffffu, common, o); sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_c[0], only_a); atomicAdd(&s_c[1], only_b); atomicAdd(&s_c[2], common); atomicAdd(&s_sum, sum);
        atomicMax(&s_mx, (unsigned long long)__double_as_longlong(mx));       // non-negative doubles order like their bit patterns
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_c[0]) atomicAdd(counts + 0, s_c[0]);
        if (s_c[1]) atomicAdd(counts + 1, s_c[1]);
        if (s_c[2]) atomicAdd(counts + 2, s_c[2]);
        if (s_sum != 0.0) atomicAdd(sums, s_sum);
        if (s_mx) atomicMax(counts + 3, s_mx);
    }
}

extern "C" int bmsp_compare_csr(bmsp_matrix_t m, const int32_t* row_ptr, const int32_t* col_idx, const float* vals, int32_t mem, void* stream,
                                int64_t* only_in_m, int64_t* only_in_csr, double* mean_rel_err, double* max_rel_err) {
    if (!m || !row_ptr) { set_error("bmsp_compare_csr: null argument"); return BMSP_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    touch(m, st);
    int32_t *a_rp = nullptr, *a_ci = nullptr, *b_rp = nullptr, *b_ci = nullptr; float *a_v = nullptr, *b_v = nullptr;
    unsigned long long* acc = nullptr;
    int s = BMSP_OK;
    auto cleanup = [&]() { dev_free(a_rp, st); dev_free(a_ci, st); dev_free(a_v, st); dev_free(acc, st); if (mem != BMSP_DEVICE) { dev_free(b_rp, st); dev_free(b_ci, st); dev_free(b_v, st); } };
#define CMP_TRY(x) do { s = (x); if (s != BMSP_OK) { cleanup(); return s; } } while (0)
#define CMP_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { cleanup(); return cuda_fail(e__, #x, __FILE__, __LINE__); } } while (0)
    CMP_TRY(dev_alloc_t(&a_rp, (size_t)m->rows + 2, st));
    CMP_TRY(dev_alloc_t(&a_ci, (size_t)m->nnz + 1, st));
    CMP_TRY(dev_alloc_t(&a_v, (size_t)m->nnz + 1, st));
    CMP_TRY(dev_alloc_t(&acc, 8, st));
    CMP_CUDA(cudaMemsetAsync(acc, 0, 8 * sizeof(unsigned long long), st));
    CMP_TRY(to_csr_device(m, a_rp, a_ci, a_v, st));
    if (mem == BMSP_DEVICE) { b_rp = const_cast<int32_t*>(row_ptr); b_ci = const_cast<int32_t*>(col_idx); b_v = const_cast<float*>(vals); }
    else {
        const int64_t n = row_ptr[m->rows];
        CMP_TRY(dev_alloc_t(&b_rp, (size_t)m->rows + 2, st));
        CMP_TRY(dev_alloc_t(&b_ci, (size_t)n + 1, st));
        CMP_TRY(dev_alloc_t(&b_v, (size_t)n + 1, st));
        CMP_CUDA(cudaMemcpyAsync(b_rp, row_ptr, sizeof(int32_t) * ((size_t)m->rows + 1), cudaMemcpyHostToDevice, st));
        if (n) {
            CMP_CUDA(cudaMemcpyAsync(b_ci, col_idx, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
            CMP_CUDA(cudaMemcpyAsync(b_v, vals, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
        }
    }
    if (m->rows > 0) {
        compare_rows_kernel<<<(unsigned)ceil_div(m->rows, 256), 256, 0, st>>>(m->rows, a_rp, a_ci, a_v, b_rp, b_ci, b_v, acc, reinterpret_cast<double*>(acc + 4));
        CMP_CUDA(cudaGetLastError());
    }
    unsigned long long h[8];
    CMP_CUDA(cudaMemcpyAsync(h, acc, sizeof(h), cudaMemcpyDeviceToHost, st));
    CMP_CUDA(cudaStreamSynchronize(st));
    double sum, mx;
    memcpy(&sum, &h[4], 8); memcpy(&mx, &h[3], 8);
    if (only_in_m) *only_in_m = (int64_t)h[0];
    if (only_in_csr) *only_in_csr = (int64_t)h[1];
    if (mean_rel_err) *mean_rel_err = h[2] ? sum / (double)h[2] : 0.0;
    if (max_rel_err) *max_rel_err = mx;
    cleanup();
    return BMSP_OK;
#undef CMP_TRY
#undef CMP_CUDA
}

extern "C" int bmsp_compare(bmsp_matrix_t m, int64_t nnz, const int32_t* rows, const int32_t* cols, const float* vals,
                            int64_t* only_in_m, int64_t* only_in_coo, double* mean_rel_err, double* max_rel_err) {
    if (!m || (nnz > 0 && (!rows || !cols || !vals))) { set_error("bmsp_compare: null argument"); return BMSP_ERR_INVALID; }
    // the COO (any order) becomes CSR on the host -- counting sort by row, columns sorted inside each row -- and the comparison
    // itself runs on the device (bmsp_compare_csr); entries outside the matrix count as "only in the COO"
    std::vector<int32_t> rp((size_t)m->rows + 1, 0);
    int64_t outside = 0;
    for (int64_t i = 0; i < nnz; i++) {
        if (rows[i] < 0 || rows[i] >= m->rows) { outside++; continue; }
        rp[(size_t)rows[i] + 1]++;
    }
    for (int32_t r = 0; r < m->rows; r++) rp[r + 1] += rp[r];
    const int64_t n_in = rp[m->rows];
    std::vector<int64_t> perm((size_t)n_in);
    {
        std::vector<int32_t> cur(rp.begin(), rp.end() - 1);
        for (int64_t i = 0; i < nnz; i++) if (rows[i] >= 0 && rows[i] < m->rows) perm[(size_t)cur[rows[i]]++] = i;
    }
    for (int32_t r = 0; r < m->rows; r++)
        std::sort(perm.begin() + rp[r], perm.begin() + rp[r + 1], [&](int64_t a, int64_t b) { return cols[a] < cols[b]; });
    std::vector<int32_t> ci((size_t)n_in); std::vector<float> v((size_t)n_in);
    for (int64_t i = 0; i < n_in; i++) { ci[i] = cols[perm[i]]; v[i] = vals[perm[i]]; }
    int64_t ob = 0;
    BMSP_TRY(bmsp_compare_csr(m, rp.data(), ci.data(), v.data(), BMSP_HOST, m->last_stream, only_in_m, &ob, mean_rel_err, max_rel_err));
    if (only_in_coo) *only_in_coo = ob + outside;
    return BMSP_OK;
}
