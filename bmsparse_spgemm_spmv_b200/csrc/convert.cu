// convert.cu -- CSR / COO / MatrixMarket -> bmSparse on the device, and the block transpose.
//
// Replaces the reference's ingest path src/bmSpMatrix.cu:111-219 (host parse -> thrust::sort with a
// 4-level comparator -> 2 reduce_by_key + scan + 2 transforms) and src/reader.cu:49-110.
// Same mapping (Appendix A of SURVEY.md), different algorithm: CSR rows are already sorted, so the
// block-row order of an entry is a sum of eight binary searches (an 8-way merge rank) -- no global
// sort.  Bitmaps are then built warp-cooperatively (head flags by ballot, segmented OR by shuffle,
// ranks by popc) and values are compacted with 16-byte stores.
#include "common.cuh"
#include <vector>
#include <algorithm>
#include <numeric>
#include <fstream>
#include <string>
#include <cstring>
#include <cstdlib>
#include <cctype>
#include <thread>

namespace bmsp {

enum { FLAG_UNSORTED = 1, FLAG_DUP = 2, FLAG_RANGE = 4, FLAG_ROWPTR = 8 };

// row_ptr of a device CSR: starts at 0, ends at nnz, never decreases (rank_kernel binary-searches it and would read out of
// bounds on anything else)
__global__ void check_rowptr_kernel(const int32_t* __restrict__ rp, int32_t rows, int64_t nnz, int32_t* __restrict__ flags) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > rows) return;
    const int32_t v = rp[r];
    bool bad = v < 0 || (int64_t)v > nnz;
    if (r == 0) bad |= v != 0;
    if (r == rows) bad |= (int64_t)v != nnz;
    else bad |= rp[r + 1] < v;
    if (bad) atomicOr(flags, FLAG_ROWPTR);
}

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int lo, int hi, int target) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// K0: the row of the first entry of every chunk of RANK_T entries (last r with rp[r] <= e), one thread per chunk.  The only
// place that searches all of row_ptr: a search per ENTRY was 24 dependent L2 round trips on a 16 M-row matrix and two thirds of the
// whole conversion (ncu launch list, profiles/r2_bench_launches_summary.txt).
constexpr int RANK_T = 256, RANK_SPAN = 1024;
__global__ void chunk_row_kernel(const int32_t* __restrict__ rp, int32_t rows, int64_t nnz, int32_t* __restrict__ chunk_row, int64_t nchunks) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nchunks) return;
    if (b == nchunks) { chunk_row[b] = rows - 1; return; }
    const int64_t e = b * RANK_T;
    int lo = 0, hi = rows;   // invariant: rp[lo] <= e < rp[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rp + mid) <= e) lo = mid; else hi = mid;
    }
    chunk_row[b] = lo;
}

// K1: one thread per CSR entry, a CTA per chunk of RANK_T entries.  The chunk's rows lie between the rows of its first entry and
// of the next chunk's first entry: their row_ptr values go to shared memory and every thread finds its row there.  It then
// validates, computes its position in the block-row order (block column, then row-major or column-major inside the block:
// block_order, bmSpMatrix.cu:45-74) and scatters (key, cell position, source index) there.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(RANK_T) rank_kernel(const int32_t* __restrict__ rp, const int32_t* __restrict__ ci, int32_t rows,
                                                      int32_t cols, int64_t nnz, const int32_t* __restrict__ chunk_row,
                                                      uint64_t* __restrict__ s_key, uint8_t* __restrict__ s_p, int32_t* __restrict__ s_src,
                                                      int32_t* __restrict__ flags) {
    __shared__ int32_t srp[RANK_SPAN + 1];
    const int r_lo = chunk_row[blockIdx.x], r_hi = chunk_row[blockIdx.x + 1];
    const int span = r_hi - r_lo + 1;                         // rows r_lo .. r_hi; rp[r_hi + 1] is past every entry of the chunk
    const bool in_smem = span <= RANK_SPAN;
    if (in_smem)
        for (int i = threadIdx.x; i <= span; i += RANK_T) srp[i] = __ldg(rp + r_lo + i);
    __syncthreads();
    int64_t e = (int64_t)blockIdx.x * RANK_T + threadIdx.x;
    if (e >= nnz) return;
    // row = last r with rp[r] <= e
    int lo, hi;              // invariant: rp[lo] <= e < rp[hi]
    if (in_smem) {
        lo = 0; hi = span;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if ((int64_t)srp[mid] <= e) lo = mid; else hi = mid;
        }
        lo += r_lo;
    } else {
        lo = r_lo; hi = r_hi + 1;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if ((int64_t)__ldg(rp + mid) <= e) lo = mid; else hi = mid;
        }
    }
    const int r = lo;
    const int c = ci[e];
    if (c < 0 || c >= cols) { atomicOr(flags, FLAG_RANGE); return; }
    const int rs = __ldg(rp + r);
    if (e > rs) {
        int prev = ci[e - 1];
        if (prev == c) atomicOr(flags, FLAG_DUP);
        else if (prev > c) atomicOr(flags, FLAG_UNSORTED);
    }
    const int br = r >> 3, bc = c >> 3, ri = r & 7, cidx = c & 7;
    const int r0 = br << 3;
    int pos = (int)(e - rs);   // entries of the own row with a smaller column
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int rr = r0 + q;
        if (q == ri || rr >= rows) continue;
        int s = __ldg(rp + rr), t = __ldg(rp + rr + 1);
        int target;
        if (TRANSPOSED) target = q < ri ? c + 1 : c;                   // (bc, ci, ri) order
        else            target = q < ri ? (bc + 1) << 3 : bc << 3;      // (bc, ri, ci) order
        pos += lower_bound_i32(ci, s, t, target) - s;
    }
    int64_t g = (int64_t)__ldg(rp + r0) + pos;
    s_key[g] = ((uint64_t)(uint32_t)br << 32) | (uint32_t)bc;
    s_p[g] = (uint8_t)(TRANSPOSED ? cidx * 8 + ri : ri * 8 + cidx);
    s_src[g] = (int32_t)e;
}

// K2: heads per tile of 1024 sorted entries (ballot + popc).
__global__ void __launch_bounds__(1024) head_count_kernel(const uint64_t* __restrict__ s_key, int64_t nnz,
                                                          uint32_t* __restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    bool head = false;
    if (i < nnz) head = (i == 0) || (s_key[i] != s_key[i - 1]);
    __shared__ uint32_t wc[32];
    uint32_t b = __ballot_sync(0xffffffffu, head);
    if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = wc[threadIdx.x];
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) counts[blockIdx.x] = v;
    }
}

// K3: block ids from the scanned head counts; keys/offsets/block columns by the head lanes; bitmaps by
// a warp-segmented OR (entries of one block are contiguous) finished with one atomicOr per segment.
__global__ void __launch_bounds__(1024) emit_kernel(const uint64_t* __restrict__ s_key, const uint8_t* __restrict__ s_p,
                                                    int64_t nnz, const uint32_t* __restrict__ tile_base,
                                                    uint64_t* __restrict__ keys, uint64_t* __restrict__ bmps,
                                                    uint64_t* __restrict__ offsets) {
    int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t key = 0;
    bool head = false;
    if (i < nnz) { key = s_key[i]; head = (i == 0) || (key != s_key[i - 1]); }
    uint32_t b = __ballot_sync(0xffffffffu, head);
    __shared__ uint32_t wc[32];
    if (lane == 0) wc[wid] = __popc(b);
    __syncthreads();
    if (wid == 0) {
        uint32_t v = wc[lane], inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        wc[lane] = inc - v;
    }
    __syncthreads();
    // inclusive head count up to this lane, minus one = block id
    int32_t blk = (int32_t)(tile_base[blockIdx.x] + wc[wid] + __popc(b & (0xffffffffu >> (31 - lane)))) - 1;
    if (i >= nnz) blk = -1;
    uint32_t hi = 0, lo = 0;
    if (i < nnz) {
        int p = s_p[i];
        if (p < 32) hi = 0x80000000u >> p; else lo = 0x80000000u >> (p - 32);
        if (head) { keys[blk] = key; offsets[blk] = (uint64_t)i; }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t th = __shfl_up_sync(0xffffffffu, hi, o), tl = __shfl_up_sync(0xffffffffu, lo, o);
        int32_t tb = __shfl_up_sync(0xffffffffu, blk, o);
        if (lane >= o && tb == blk) { hi |= th; lo |= tl; }
    }
    int32_t nb = __shfl_down_sync(0xffffffffu, blk, 1);
    if (blk >= 0 && (lane == 31 || nb != blk))
        atomicOr((unsigned long long*)(bmps + blk), ((unsigned long long)hi << 32) | lo);
}

// K4: values in sorted order, 8 per thread, 16-byte stores.
template <typename Tin, typename Tout> __device__ __forceinline__ Tout conv(Tin v);
template <> __device__ __forceinline__ __half conv<float, __half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __half conv<__half, __half>(__half v) { return v; }
template <> __device__ __forceinline__ float conv<float, float>(float v) { return v; }
template <> __device__ __forceinline__ float conv<__half, float>(__half v) { return __half2float(v); }

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) values_kernel(const Tin* __restrict__ vals, const int32_t* __restrict__ s_src, int64_t nnz,
                                                     Tout* __restrict__ out) {
    int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= nnz) return;
    Tout v[8];
    if (i0 + 8 <= nnz) {
        int4 a = *reinterpret_cast<const int4*>(s_src + i0), b = *reinterpret_cast<const int4*>(s_src + i0 + 4);
        int idx[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = conv<Tin, Tout>(vals[idx[k]]);
        if (sizeof(Tout) == 2) {
            *reinterpret_cast<int4*>(out + i0) = *reinterpret_cast<int4*>(v);
        } else {
            *reinterpret_cast<int4*>(out + i0) = *reinterpret_cast<int4*>(v);
            *reinterpret_cast<int4*>(out + i0 + 4) = *reinterpret_cast<int4*>(v + 4);
        }
    } else {
        for (int64_t i = i0; i < nnz; i++) out[i] = conv<Tin, Tout>(vals[s_src[i]]);
    }
}

static size_t dsize(int dtype) { return dtype == BMSP_F16 ? 2 : 4; }

// device CSR -> new matrix
static int convert_device_csr(int32_t rows, int32_t cols, int64_t nnz, const int32_t* rp, const int32_t* ci, const void* vals,
                              int vals_dtype, int transposed, int out_dtype, cudaStream_t st, bmsp_matrix_t* out) {
    if (nnz > 0xFFFFFFFFll) { set_error("nnz %lld exceeds 2^32-1", (long long)nnz); return BMSP_ERR_TOO_LARGE; }
    bmsp_matrix_s* m = new bmsp_matrix_s();
    touch(m, st);
    m->rows = rows; m->cols = cols; m->nnz = nnz; m->dtype = out_dtype; m->transposed = transposed;
    uint64_t* s_key = nullptr; uint8_t* s_p = nullptr; int32_t* s_src = nullptr; int32_t* flags = nullptr;
    uint32_t* counts = nullptr; int32_t* chunk_row = nullptr;
    int status = BMSP_OK;
    auto cleanup = [&]() { dev_free(s_key, st); dev_free(s_p, st); dev_free(s_src, st); dev_free(flags, st); dev_free(counts, st); dev_free(chunk_row, st); };
    auto fail = [&](int code) { cleanup(); bmsp_destroy(m); return code; };
#define CV_TRY(x) do { status = (x); if (status != BMSP_OK) return fail(status); } while (0)
#define CV_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return fail(cuda_fail(e__, #x, __FILE__, __LINE__)); } while (0)
    int64_t tiles = ceil_div(nnz, 1024);
    uint32_t nblk32 = 0;
    CV_TRY(dev_alloc_t(&flags, 1, st));
    CV_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
    {
        check_rowptr_kernel<<<(unsigned)ceil_div((int64_t)rows + 1, 256), 256, 0, st>>>(rp, rows, nnz, flags);
        CV_CUDA(cudaGetLastError());
        int32_t hflags = 0;
        CV_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CV_CUDA(cudaStreamSynchronize(st));
        if (hflags & FLAG_ROWPTR) { set_error("row_ptr must start at 0, end at nnz and never decrease"); return fail(BMSP_ERR_INVALID); }
    }
    if (nnz > 0) {
        CV_TRY(dev_alloc_t(&s_key, (size_t)nnz, st));
        CV_TRY(dev_alloc_t(&s_p, (size_t)nnz, st));
        CV_TRY(dev_alloc_t(&s_src, (size_t)nnz + 8, st));
        CV_TRY(dev_alloc_t(&counts, (size_t)tiles + 1, st));
        const int64_t nchunks = ceil_div(nnz, RANK_T);
        CV_TRY(dev_alloc_t(&chunk_row, (size_t)nchunks + 1, st));
        chunk_row_kernel<<<(unsigned)ceil_div(nchunks + 1, 256), 256, 0, st>>>(rp, rows, nnz, chunk_row, nchunks);
        CV_CUDA(cudaGetLastError());
        unsigned grid = (unsigned)nchunks;
        if (transposed) rank_kernel<true><<<grid, RANK_T, 0, st>>>(rp, ci, rows, cols, nnz, chunk_row, s_key, s_p, s_src, flags);
        else            rank_kernel<false><<<grid, RANK_T, 0, st>>>(rp, ci, rows, cols, nnz, chunk_row, s_key, s_p, s_src, flags);
        CV_CUDA(cudaGetLastError());
        int32_t hflags = 0;
        CV_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CV_CUDA(cudaStreamSynchronize(st));
        if (hflags & FLAG_RANGE) { set_error("column index outside [0, %d)", cols); return fail(BMSP_ERR_RANGE); }
        if (hflags & FLAG_UNSORTED) { set_error("CSR column indices must ascend inside each row"); return fail(BMSP_ERR_UNSORTED); }
        if (hflags & FLAG_DUP) { set_error("duplicate (row, col) entries are not supported"); return fail(BMSP_ERR_DUPLICATE); }
        head_count_kernel<<<(unsigned)tiles, 1024, 0, st>>>(s_key, nnz, counts);
        CV_CUDA(cudaGetLastError());
        CV_TRY(exclusive_scan_u32(counts, counts, tiles, st));
        CV_CUDA(cudaMemcpyAsync(&nblk32, counts + tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CV_CUDA(cudaStreamSynchronize(st));
    }
    m->nblk = nblk32; m->offsets_len = m->nblk;
    CV_TRY(dev_alloc_t(&m->keys, (size_t)m->nblk + 2, st));
    CV_TRY(dev_alloc_t(&m->bmps, (size_t)m->nblk + 2, st));
    CV_TRY(dev_alloc_t(&m->offsets, (size_t)m->nblk + 2, st));
    CV_TRY(dev_alloc(&m->values, (size_t)nnz * dsize(out_dtype) + 16, st));
    if (nnz > 0) {
        CV_CUDA(cudaMemsetAsync(m->bmps, 0, sizeof(uint64_t) * m->nblk, st));
        emit_kernel<<<(unsigned)tiles, 1024, 0, st>>>(s_key, s_p, nnz, counts, m->keys, m->bmps, m->offsets);
        CV_CUDA(cudaGetLastError());
        unsigned grid = (unsigned)ceil_div(ceil_div(nnz, 8), 256);
        if (vals_dtype == BMSP_F32 && out_dtype == BMSP_F16) values_kernel<float, __half><<<grid, 256, 0, st>>>((const float*)vals, s_src, nnz, (__half*)m->values);
        else if (vals_dtype == BMSP_F16 && out_dtype == BMSP_F16) values_kernel<__half, __half><<<grid, 256, 0, st>>>((const __half*)vals, s_src, nnz, (__half*)m->values);
        else if (vals_dtype == BMSP_F32 && out_dtype == BMSP_F32) values_kernel<float, float><<<grid, 256, 0, st>>>((const float*)vals, s_src, nnz, (float*)m->values);
        else values_kernel<__half, float><<<grid, 256, 0, st>>>((const __half*)vals, s_src, nnz, (float*)m->values);
        CV_CUDA(cudaGetLastError());
    }
    CV_TRY(derive_compact(m, st));
    cleanup();
    *out = m;
    return BMSP_OK;
#undef CV_TRY
#undef CV_CUDA
}

// ---------------------------------------------------------------------------- block transpose
// 8x8 bit-matrix transpose (three masked swaps); with MSB-first cell numbering the main diagonal is
// bits 63,54,...,0, so the classic delta-swap constants apply unchanged.
__device__ __forceinline__ uint64_t transpose8x8(uint64_t x) {
    uint64_t t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;  x = x ^ t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x = x ^ t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x = x ^ t ^ (t << 28);
    return x;
}

// One thread per block: new bitmap = transpose; value of cell p=(hi,lo) moves to the rank of cell (lo,hi).
template <typename Tin, typename Tout>
__global__ void block_transpose_kernel(const uint64_t* __restrict__ bmps, const uint64_t* __restrict__ offsets,
                                       const Tin* __restrict__ vin, uint64_t* __restrict__ bmps_out,
                                       Tout* __restrict__ vout, int64_t nblk) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    uint64_t bmp = bmps[b], t = transpose8x8(bmp), off = offsets[b];
    bmps_out[b] = t;
    uint64_t rem = bmp;
    int k = 0;
    while (rem) {
        int p = __clzll((long long)rem);
        rem &= ~(0x8000000000000000ull >> p);
        int q = ((p & 7) << 3) | (p >> 3);
        int rank = q == 0 ? 0 : __popcll(t >> (64 - q));
        vout[off + rank] = conv<Tin, Tout>(vin[off + k]);
        k++;
    }
}

}  // namespace bmsp

using namespace bmsp;

extern "C" {

int bmsp_create_from_csr(int32_t rows, int32_t cols, int64_t nnz, const int32_t* row_ptr, const int32_t* col_idx,
                         const void* vals, int32_t vals_dtype, int32_t mem, int32_t transposed, int32_t out_dtype,
                         void* stream, bmsp_matrix_t* out) {
    if (!out || rows < 0 || cols < 0 || nnz < 0 || !row_ptr || (nnz > 0 && (!col_idx || !vals)) ||
        (vals_dtype != BMSP_F16 && vals_dtype != BMSP_F32) || (out_dtype != BMSP_F16 && out_dtype != BMSP_F32)) {
        set_error("bmsp_create_from_csr: invalid argument");
        return BMSP_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (mem == BMSP_DEVICE)
        return convert_device_csr(rows, cols, nnz, row_ptr, col_idx, vals, vals_dtype, transposed, out_dtype, st, out);
    if (row_ptr[0] != 0 || (int64_t)row_ptr[rows] != nnz) { set_error("row_ptr[0] must be 0 and row_ptr[rows] == nnz"); return BMSP_ERR_INVALID; }
    int32_t *d_rp = nullptr, *d_ci = nullptr; void* d_v = nullptr;
    BMSP_TRY(dev_alloc_t(&d_rp, (size_t)rows + 1, st));
    BMSP_TRY(dev_alloc_t(&d_ci, (size_t)nnz, st));
    BMSP_TRY(dev_alloc(&d_v, (size_t)nnz * dsize(vals_dtype), st));
    BMSP_CUDA(cudaMemcpyAsync(d_rp, row_ptr, sizeof(int32_t) * ((size_t)rows + 1), cudaMemcpyHostToDevice, st));
    if (nnz) {
        BMSP_CUDA(cudaMemcpyAsync(d_ci, col_idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
        BMSP_CUDA(cudaMemcpyAsync(d_v, vals, (size_t)nnz * dsize(vals_dtype), cudaMemcpyHostToDevice, st));
    }
    int s = convert_device_csr(rows, cols, nnz, d_rp, d_ci, d_v, vals_dtype, transposed, out_dtype, st, out);
    cudaStreamSynchronize(st);
    dev_free(d_rp, st); dev_free(d_ci, st); dev_free(d_v, st);
    return s;
}

// host COO -> CSR (counting sort by row, columns sorted inside each row); duplicates summed when `merge`
static int coo_to_matrix(int32_t rows, int32_t cols, int64_t nnz, const int32_t* row_idx, const int32_t* col_idx, const double* vals,
                         int32_t transposed, int32_t out_dtype, int32_t flags, void* stream, bmsp_matrix_t* out) {
    std::vector<int32_t> rp((size_t)rows + 1, 0);
    for (int64_t i = 0; i < nnz; i++) {
        if (row_idx[i] < 0 || row_idx[i] >= rows) { set_error("row index outside [0, %d)", rows); return BMSP_ERR_RANGE; }
        rp[(size_t)row_idx[i] + 1]++;
    }
    for (int32_t r = 0; r < rows; r++) rp[r + 1] += rp[r];
    std::vector<int64_t> perm((size_t)nnz);
    {
        std::vector<int32_t> cur(rp.begin(), rp.end() - 1);
        for (int64_t i = 0; i < nnz; i++) perm[(size_t)cur[row_idx[i]]++] = i;
    }
    for (int32_t r = 0; r < rows; r++)
        std::stable_sort(perm.begin() + rp[r], perm.begin() + rp[r + 1], [&](int64_t a, int64_t b) { return col_idx[a] < col_idx[b]; });
    std::vector<int32_t> ci((size_t)nnz);
    std::vector<double> dv((size_t)nnz);
    for (int64_t i = 0; i < nnz; i++) { ci[i] = col_idx[perm[i]]; dv[i] = vals[perm[i]]; }
    int64_t n = nnz;
    if (flags & BMSP_MERGE_DUPLICATES) {
        // repeated (row, col) entries are summed in file order (what assembling a finite-element matrix means by them); the
        // reference keeps them apart and corrupts its offsets, without this flag they are rejected
        std::vector<int32_t> nrp((size_t)rows + 1, 0);
        int64_t w = 0;
        for (int32_t r = 0; r < rows; r++) {
            for (int64_t i = rp[r]; i < rp[r + 1]; i++) {
                if (i > rp[r] && ci[i] == ci[w - 1] && w > nrp[r]) dv[w - 1] += dv[i];
                else { ci[w] = ci[i]; dv[w] = dv[i]; w++; }
            }
            nrp[r + 1] = (int32_t)w;
        }
        rp.swap(nrp); n = w;
    }
    if (out_dtype == BMSP_F16) {
        std::vector<__half> v((size_t)n);
        for (int64_t i = 0; i < n; i++) v[i] = __double2half(dv[i]);
        return bmsp_create_from_csr(rows, cols, n, rp.data(), ci.data(), v.data(), BMSP_F16, BMSP_HOST, transposed, BMSP_F16, stream, out);
    }
    std::vector<float> v((size_t)n);
    for (int64_t i = 0; i < n; i++) v[i] = (float)dv[i];
    return bmsp_create_from_csr(rows, cols, n, rp.data(), ci.data(), v.data(), BMSP_F32, BMSP_HOST, transposed, out_dtype, stream, out);
}

int bmsp_create_from_coo_ex(int32_t rows, int32_t cols, int64_t nnz, const int32_t* row_idx, const int32_t* col_idx,
                            const double* vals, int32_t transposed, int32_t out_dtype, int32_t flags, void* stream, bmsp_matrix_t* out) {
    if (!out || rows < 0 || cols < 0 || nnz < 0 || (nnz > 0 && (!row_idx || !col_idx || !vals))) {
        set_error("bmsp_create_from_coo: invalid argument");
        return BMSP_ERR_INVALID;
    }
    // values cast straight from double (bmSpMatrix.cu:136-157 casts the parsed double to valueType in one rounding)
    return coo_to_matrix(rows, cols, nnz, row_idx, col_idx, vals, transposed, out_dtype, flags, stream, out);
}

int bmsp_create_from_coo(int32_t rows, int32_t cols, int64_t nnz, const int32_t* row_idx, const int32_t* col_idx,
                         const double* vals, int32_t transposed, int32_t out_dtype, void* stream, bmsp_matrix_t* out) {
    return bmsp_create_from_coo_ex(rows, cols, nnz, row_idx, col_idx, vals, transposed, out_dtype, 0, stream, out);
}

// MatrixMarket coordinate files.  Banner rules follow cusp's reader (cusp/io/detail/matrix_market.inl:70-95): five tokens,
// storage coordinate | array, type real | integer | pattern | complex, symmetry general | symmetric | skew-symmetric | hermitian.
//   pattern            value 1 (:171)                       integer           parsed like real (:177-190)
//   symmetric          off-diagonal entries mirrored (:257-279; bmSpMatrix.cu:113-149 does the same)
//   skew-symmetric     mirrored with the sign flipped (cusp throws not_implemented :286-290; the reference's own reader would mirror it
//                      with the SAME sign because "skew-symmetric" contains "symmetric", bmSpMatrix.cu:113-120)
//   complex, hermitian, array storage: BMSP_ERR_UNSUPPORTED (a real-valued bmSpMatrix cannot hold them; cusp keeps only the real part)
// Indices are checked against the size line (:222-235).  The body is parsed by all host cores: the text is cut at line ends into
// one slice per thread (strtol / strtod, no streams).
int bmsp_create_from_mtx_ex(const char* path, int32_t transposed, int32_t out_dtype, int32_t flags, void* stream, bmsp_matrix_t* out) {
    if (!path || !out) { set_error("bmsp_create_from_mtx: null argument"); return BMSP_ERR_INVALID; }
    std::ifstream f(path, std::ios::binary);
    if (!f) { set_error("cannot open %s", path); return BMSP_ERR_IO; }
    std::string content((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const char* p = content.c_str();
    const char* end = p + content.size();
    auto line_end = [&](const char* q) { while (q < end && *q != '\n') q++; return q; };
    const char* le = line_end(p);
    std::string banner(p, le);
    std::vector<std::string> tok;
    {
        size_t i = 0;
        while (i < banner.size()) {
            while (i < banner.size() && isspace((unsigned char)banner[i])) i++;
            size_t j = i;
            while (j < banner.size() && !isspace((unsigned char)banner[j])) j++;
            if (j > i) tok.push_back(banner.substr(i, j - i));
            i = j;
        }
        for (auto& t : tok) for (auto& ch : t) ch = (char)tolower((unsigned char)ch);
    }
    bool symmetric = false, skew = false, pattern = false;
    if (tok.size() == 5 && tok[0] == "%%matrixmarket") {
        if (tok[1] != "matrix") { set_error("%s: invalid MatrixMarket banner", path); return BMSP_ERR_IO; }
        if (tok[2] == "array") { set_error("%s: array storage is not supported (coordinate files only)", path); return BMSP_ERR_UNSUPPORTED; }
        if (tok[2] != "coordinate") { set_error("%s: invalid MatrixMarket storage format [%s]", path, tok[2].c_str()); return BMSP_ERR_IO; }
        if (tok[3] == "complex") { set_error("%s: complex matrices are not supported", path); return BMSP_ERR_UNSUPPORTED; }
        if (tok[3] == "pattern") pattern = true;
        else if (tok[3] != "real" && tok[3] != "integer") { set_error("%s: invalid MatrixMarket data type [%s]", path, tok[3].c_str()); return BMSP_ERR_IO; }
        if (tok[4] == "hermitian") { set_error("%s: hermitian matrices are not supported", path); return BMSP_ERR_UNSUPPORTED; }
        if (tok[4] == "symmetric") symmetric = true;
        else if (tok[4] == "skew-symmetric") skew = true;
        else if (tok[4] != "general") { set_error("%s: invalid MatrixMarket symmetry [%s]", path, tok[4].c_str()); return BMSP_ERR_IO; }
    } else {
        // no well-formed banner: the reference's rule -- "symmetric" anywhere in the first line (bmSpMatrix.cu:113-120)
        symmetric = banner.find("symmetric") != std::string::npos;
        pattern = banner.find("pattern") != std::string::npos;
        if (!banner.empty() && banner[0] != '%') le = p - 1;          // the first line already is the size line
    }
    p = le < end ? le + 1 : end;
    while (p < end && (*p == '%' || *p == '\n' || *p == '\r')) { le = line_end(p); p = le < end ? le + 1 : end; }
    char* q = nullptr;
    long nr = strtol(p, &q, 10); p = q;
    long nc = strtol(p, &q, 10); p = q;
    long long nl = strtoll(p, &q, 10); p = q;
    if (nr <= 0 || nc <= 0 || nl < 0 || nr > 0x7FFFFFFFl || nc > 0x7FFFFFFFl) { set_error("%s: bad size line", path); return BMSP_ERR_IO; }
    le = line_end(p); p = le < end ? le + 1 : end;
    // slices of whole lines, one per thread
    int nth = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if ((end - p) < (1 << 20)) nth = 1;
    std::vector<const char*> cut((size_t)nth + 1);
    cut[0] = p; cut[nth] = end;
    for (int t = 1; t < nth; t++) { const char* c = p + (end - p) * t / nth; c = line_end(c); cut[t] = c < end ? c + 1 : end; }
    struct Part { std::vector<int32_t> r, c; std::vector<double> v; int err = 0; long long lines = 0; };
    std::vector<Part> parts((size_t)nth);
    auto parse = [&](int t) {
        Part& P = parts[(size_t)t];
        const char* s = cut[t]; const char* e = cut[t + 1];
        P.r.reserve((size_t)((e - s) / 12 + 16)); P.c.reserve(P.r.capacity()); P.v.reserve(P.r.capacity());
        char* qq = nullptr;
        while (s < e) {
            while (s < e && (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n')) s++;
            if (s >= e) break;
            if (*s == '%') { while (s < e && *s != '\n') s++; continue; }
            long i = strtol(s, &qq, 10); if (qq == s) { P.err = BMSP_ERR_IO; return; } s = qq;
            long j = strtol(s, &qq, 10); if (qq == s) { P.err = BMSP_ERR_IO; return; } s = qq;
            double x = 1.0;
            if (!pattern) { x = strtod(s, &qq); if (qq == s) { P.err = BMSP_ERR_IO; return; } s = qq; }
            if (i < 1 || i > nr || j < 1 || j > nc) { P.err = BMSP_ERR_RANGE; return; }
            P.lines++;
            P.r.push_back((int32_t)(i - 1)); P.c.push_back((int32_t)(j - 1)); P.v.push_back(x);
            if ((symmetric || skew) && i != j) { P.r.push_back((int32_t)(j - 1)); P.c.push_back((int32_t)(i - 1)); P.v.push_back(skew ? -x : x); }
            while (s < e && *s != '\n') s++;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nth; t++) th.emplace_back(parse, t);
        parse(0);
        for (auto& t : th) t.join();
    }
    size_t total = 0; long long lines = 0;
    for (auto& P : parts) {
        if (P.err == BMSP_ERR_RANGE) { set_error("%s: index outside the %ld x %ld matrix", path, nr, nc); return BMSP_ERR_RANGE; }
        if (P.err) { set_error("%s: malformed entry line", path); return BMSP_ERR_IO; }
        total += P.r.size(); lines += P.lines;
    }
    std::vector<int32_t> r(total), c(total); std::vector<double> v(total);
    size_t w = 0;
    for (auto& P : parts) {
        std::copy(P.r.begin(), P.r.end(), r.begin() + w); std::copy(P.c.begin(), P.c.end(), c.begin() + w); std::copy(P.v.begin(), P.v.end(), v.begin() + w);
        w += P.r.size();
    }
    if (lines != nl) { set_error("%s: %lld entry lines, the size line announces %lld", path, lines, nl); return BMSP_ERR_IO; }
    return coo_to_matrix((int32_t)nr, (int32_t)nc, (int64_t)total, r.data(), c.data(), v.data(), transposed, out_dtype, flags, stream, out);
}

int bmsp_create_from_mtx(const char* path, int32_t transposed, int32_t out_dtype, void* stream, bmsp_matrix_t* out) {
    return bmsp_create_from_mtx_ex(path, transposed, out_dtype, 0, stream, out);
}

int bmsp_block_transpose(bmsp_matrix_t A, int32_t out_dtype, void* stream, bmsp_matrix_t* At) {
    if (!A || !At || (out_dtype != BMSP_F16 && out_dtype != BMSP_F32)) { set_error("bmsp_block_transpose: invalid argument"); return BMSP_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    touch(A, st);
    bmsp_matrix_s* m = new bmsp_matrix_s();
    touch(m, st);
    m->rows = A->rows; m->cols = A->cols; m->nnz = A->nnz; m->nblk = A->nblk; m->offsets_len = A->offsets_len;
    m->dtype = out_dtype; m->transposed = !A->transposed;
    int s;
    auto fail = [&](int code) { bmsp_destroy(m); return code; };
    if ((s = dev_alloc_t(&m->keys, (size_t)m->nblk + 2, st))) return fail(s);
    if ((s = dev_alloc_t(&m->bmps, (size_t)m->nblk + 2, st))) return fail(s);
    if ((s = dev_alloc_t(&m->offsets, (size_t)m->nblk + 2, st))) return fail(s);
    if ((s = dev_alloc(&m->values, (size_t)m->nnz * dsize(out_dtype) + 16, st))) return fail(s);
    if (m->nblk) {
        cudaError_t ce = cudaMemcpyAsync(m->keys, A->keys, sizeof(uint64_t) * m->nblk, cudaMemcpyDeviceToDevice, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(m->offsets, A->offsets, sizeof(uint64_t) * m->offsets_len, cudaMemcpyDeviceToDevice, st);
        if (ce != cudaSuccess) return fail(cuda_fail(ce, "block_transpose: copy keys/offsets", __FILE__, __LINE__));
        unsigned grid = (unsigned)ceil_div(m->nblk, 128);
        if (A->dtype == BMSP_F16 && out_dtype == BMSP_F16) block_transpose_kernel<__half, __half><<<grid, 128, 0, st>>>(A->bmps, A->offsets, (const __half*)A->values, m->bmps, (__half*)m->values, m->nblk);
        else if (A->dtype == BMSP_F32 && out_dtype == BMSP_F16) block_transpose_kernel<float, __half><<<grid, 128, 0, st>>>(A->bmps, A->offsets, (const float*)A->values, m->bmps, (__half*)m->values, m->nblk);
        else if (A->dtype == BMSP_F32 && out_dtype == BMSP_F32) block_transpose_kernel<float, float><<<grid, 128, 0, st>>>(A->bmps, A->offsets, (const float*)A->values, m->bmps, (float*)m->values, m->nblk);
        else block_transpose_kernel<__half, float><<<grid, 128, 0, st>>>(A->bmps, A->offsets, (const __half*)A->values, m->bmps, (float*)m->values, m->nblk);
        ce = cudaGetLastError();
        if (ce != cudaSuccess) return fail(cuda_fail(ce, "block_transpose", __FILE__, __LINE__));
    }
    if ((s = derive_compact(m, st))) return fail(s);
    *At = m;
    return BMSP_OK;
}

}  // extern "C"
