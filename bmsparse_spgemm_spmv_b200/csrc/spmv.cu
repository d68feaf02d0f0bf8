// spmv.cu -- y = A x on the bmSparse form, HBM-streaming, no tensor cores.
//
// Replaces bmSparse_SpMV / spmv_kernel (src/bmSparse_SPMV.cu:153-230).  The reference rebuilds the
// block-row pointers with reduce_by_key + scan on every call (:196-206), runs one 64-thread CTA per
// block row with thread <-> cell and re-reads key/bitmap/offset (24 B) from all 64 threads per block.
// Here the matrix is streamed exactly once through the compact surface (8 B bitmap + 4 B block column
// per block, 2 B per value, 8 B per block row):
//
//   path 0 "row-tiled"  (dense-ish blocks: Poisson, block-clustered)
//       persistent CTAs; a tile = 64 block rows; its bitmaps / block columns / values / row pointers are
//       bulk-copied (cp.async.bulk + mbarrier, a 4-stage ring) into shared memory; one thread per matrix
//       row walks the blocks of its block row, decodes its 8-bit row mask, ranks with popc and
//       accumulates in fp32; y is stored coalesced.  x is read through L1/L2 (.nc).
//   path 1 "block-parallel" (about one value per block: uniform random, R-MAT)
//       one warp per work item (a block row, or a 4096-block slice of a long one); lane <-> block,
//       coalesced 8 B + 4 B metadata loads, value offsets by a warp scan of popc, eight per-row partial
//       sums per lane kept in shared memory, shuffle reduction at the end; sliced rows are finished by
//       a deterministic fix-up kernel.
#include "common.cuh"

namespace bmsp {

constexpr int RT = 64;             // block rows per tile (path 0)
constexpr int SPMV_THREADS = 2 * RT;   // two threads per block row: one per 32-bit bitmap half (4 matrix rows each)
constexpr int SLICE = 4096;        // blocks per work item (path 1)
constexpr int ROWSLOT = RT + 4;    // staged row-pointer slice, padded to a 16-byte multiple

struct TileMeta { int32_t p0a, p0c; uint32_t v0a; int32_t staged; };

template <typename T>
struct SpmvArgs {
    const uint64_t* bmps; const int32_t* bcol; const T* values;
    const int32_t* brp; const uint32_t* rvb;
    int32_t rows, nbr, ntiles, cap_blk, cap_val;
};

__host__ __device__ inline size_t stage_bytes(int cap_blk, int cap_val, int vsize) {
    return (size_t)(cap_blk + 4) * 8 + (size_t)(cap_blk + 4) * 4 + (size_t)(cap_val + 8) * vsize + 2 * ROWSLOT * 4 + 16 + 16;
}

template <typename X> __device__ __forceinline__ float ld_x(const X* x, uint32_t i);
template <> __device__ __forceinline__ float ld_x<float>(const float* x, uint32_t i) { return __ldg(x + i); }
template <> __device__ __forceinline__ float ld_x<__half>(const __half* x, uint32_t i) { return __half2float(__ldg(x + i)); }

// One 8-bit row mask: walk its set bits MSB-first (column 0 first), values are consecutive from kk.
template <typename T, typename X>
__device__ __forceinline__ void row_bits(uint32_t byte, const T* __restrict__ vals, uint32_t& kk, const X* __restrict__ x,
                                         uint32_t xb7, float& acc) {
    while (byte) {
        const uint32_t f = 31u - __clz(byte);          // highest set bit, f = 7 - column
        acc = fmaf(val_to_f32(vals[kk]), ld_x<X>(x, xb7 - f), acc);
        kk++;
        byte ^= 1u << f;
    }
}

// The blocks [pb, pe) of one block row, seen by the thread that owns bitmap half `h` (rows 4h..4h+3).
// bm / bc / vals are indexed relative to the tile's aligned bases (shared memory for staged tiles).
template <typename T, typename X>
__device__ __forceinline__ void half_block_row(const uint64_t* __restrict__ bm, const int32_t* __restrict__ bc,
                                               const T* __restrict__ vals, int pb, int pe, uint32_t k, const int h,
                                               const X* __restrict__ x, float (&acc)[4]) {
    for (int b = pb; b < pe; b++) {
        const uint2 w2 = *reinterpret_cast<const uint2*>(bm + b);     // .y = rows 0-3, .x = rows 4-7
        const uint32_t nhi = __popc(w2.y), nlo = __popc(w2.x);
        const uint32_t w = h ? w2.x : w2.y;
        uint32_t kk = k + (h ? nhi : 0u);
        k += nhi + nlo;
        if (w) {
            const uint32_t xb7 = (uint32_t)bc[b] * 8u + 7u;
            row_bits<T, X>(w >> 24, vals, kk, x, xb7, acc[0]);
            row_bits<T, X>((w >> 16) & 0xFFu, vals, kk, x, xb7, acc[1]);
            row_bits<T, X>((w >> 8) & 0xFFu, vals, kk, x, xb7, acc[2]);
            row_bits<T, X>(w & 0xFFu, vals, kk, x, xb7, acc[3]);
        }
    }
}

// One CTA per tile, a single staging buffer: latency is hidden by the 16 CTAs (64 warps) resident per SM,
// each in a different phase (row pointers -> bulk copies in flight -> compute -> store).
template <typename T, typename X>
__global__ void __launch_bounds__(SPMV_THREADS, 16) spmv_rowtile_kernel(SpmvArgs<T> a, const X* __restrict__ x, float* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VA = 16 / sizeof(T);   // values per 16 bytes
    const int tid = threadIdx.x;
    const size_t off_bc = (size_t)(a.cap_blk + 4) * 8, off_val = (size_t)(a.cap_blk + 4) * 12;
    const size_t off_brp = off_val + (size_t)(a.cap_val + 8) * sizeof(T);
    uint64_t* s_bm = reinterpret_cast<uint64_t*>(smem);
    int32_t* s_bc = reinterpret_cast<int32_t*>(smem + off_bc);
    T* s_val = reinterpret_cast<T*>(smem + off_val);
    int32_t* s_brp = reinterpret_cast<int32_t*>(smem + off_brp);
    uint32_t* s_rvb = reinterpret_cast<uint32_t*>(smem + off_brp + ROWSLOT * 4);
    TileMeta* s_meta = reinterpret_cast<TileMeta*>(smem + off_brp + 2 * ROWSLOT * 4);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + off_brp + 2 * ROWSLOT * 4 + 16);

    const int t = blockIdx.x;
    const int r0 = t * RT, r1 = min(r0 + RT, a.nbr);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        const int p0 = a.brp[r0], p1 = a.brp[r1];
        const uint32_t v0 = a.rvb[r0], v1 = a.rvb[r1];
        TileMeta m;
        m.p0a = p0 & ~1; m.p0c = p0 & ~3; m.v0a = v0 & ~(uint32_t)(VA - 1);
        m.staged = (p1 - p0 <= a.cap_blk) && ((int64_t)v1 - v0 <= a.cap_val);
        *s_meta = m;
        const uint32_t nrow = (uint32_t)(((r1 - r0 + 1) + 3) & ~3) * 4;
        uint32_t n8 = 0, n4 = 0, nv = 0;
        if (m.staged) {
            n8 = (uint32_t)((p1 - m.p0a + 1) & ~1) * 8;
            n4 = (uint32_t)((p1 - m.p0c + 3) & ~3) * 4;
            nv = (uint32_t)((v1 - m.v0a + VA - 1) & ~(uint32_t)(VA - 1)) * sizeof(T);
        }
        mbar_arrive_expect_tx(bar, n8 + n4 + nv + 2 * nrow);
        bulk_g2s(s_brp, a.brp + r0, nrow, bar);
        bulk_g2s(s_rvb, a.rvb + r0, nrow, bar);
        if (n8) bulk_g2s(s_bm, a.bmps + m.p0a, n8, bar);
        if (n4) bulk_g2s(s_bc, a.bcol + m.p0c, n4, bar);
        if (nv) bulk_g2s(s_val, a.values + m.v0a, nv, bar);
    }
    __syncthreads();                 // barrier initialised and armed before anyone polls it
    mbar_wait(bar, 0);

    const int lbr = tid >> 1, h = tid & 1;
    const TileMeta m = *s_meta;
    const int64_t row = ((int64_t)(r0 + lbr)) * 8 + h * 4;
    if (row >= a.rows) return;
    const int pb = s_brp[lbr], pe = s_brp[lbr + 1];
    const uint32_t k = s_rvb[lbr] - m.v0a;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (m.staged) half_block_row<T, X>(s_bm - m.p0a, s_bc - m.p0c, s_val, pb, pe, k, h, x, acc);
    else          half_block_row<T, X>(a.bmps, a.bcol, a.values + m.v0a, pb, pe, k, h, x, acc);
    if (row + 4 <= a.rows) *reinterpret_cast<float4*>(y + row) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else for (int q = 0; q < 4; q++) if (row + q < a.rows) y[row + q] = acc[q];
}

// ------------------------------------------------------------------------------------ path 1
// work item: x = block row, y = first block, z = end block, w = 1 when the block row is sliced
template <typename T, typename X>
__global__ void __launch_bounds__(256) spmv_blockpar_kernel(const uint64_t* __restrict__ bmps, const int32_t* __restrict__ bcol,
                                                           const uint64_t* __restrict__ offsets, const T* __restrict__ values,
                                                           const int4* __restrict__ work, int n_work, int rows,
                                                           const X* __restrict__ x, float* __restrict__ y,
                                                           float* __restrict__ partial) {
    __shared__ float s_acc[8][8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int item = blockIdx.x * 8 + wid;
    if (item >= n_work) return;
    const int4 w = work[item];
    float (*acc)[32] = s_acc[wid];
#pragma unroll
    for (int r = 0; r < 8; r++) acc[r][lane] = 0.f;
    uint64_t vbase = w.y < w.z ? offsets[w.y] : 0;
    for (int b0 = w.y; b0 < w.z; b0 += 32) {
        const int b = b0 + lane;
        const bool valid = b < w.z;
        uint64_t bmp = valid ? ld_stream_u64(bmps + b) : 0ull;
        const uint32_t xb = valid ? (uint32_t)ld_stream_s32(bcol + b) * 8u : 0u;
        const uint32_t cnt = __popcll(bmp);
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        uint64_t k = vbase + inc - cnt;
        vbase += __shfl_sync(0xffffffffu, inc, 31);
        while (bmp) {
            const int p = __clzll((long long)bmp);
            bmp &= ~(0x8000000000000000ull >> p);
            acc[p >> 3][lane] += val_to_f32(values[k]) * ld_x<X>(x, xb + (uint32_t)(p & 7));
            k++;
        }
    }
    float res = 0.f;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        float v = acc[r][lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == r) res = v;
    }
    if (lane < 8) {
        if (w.w) partial[(int64_t)item * 8 + lane] = res;
        else {
            const int64_t row = (int64_t)w.x * 8 + lane;
            if (row < rows) y[row] = res;
        }
    }
}

// sliced block rows: sum the slices' partials in slice order (deterministic), one thread per matrix row
__global__ void spmv_fixup_kernel(const int32_t* __restrict__ item_ofs, const float* __restrict__ partial, int nbr, int rows,
                                  float* __restrict__ y) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int br = (int)(row >> 3), r = (int)(row & 7);
    const int i0 = item_ofs[br], i1 = item_ofs[br + 1];
    if (i1 - i0 <= 1) return;
    float s = 0.f;
    for (int i = i0; i < i1; i++) s += partial[(int64_t)i * 8 + r];
    y[row] = s;
}

__global__ void work_count_kernel(const int32_t* __restrict__ brp, int nbr, uint32_t* __restrict__ cnt) {
    int br = blockIdx.x * blockDim.x + threadIdx.x;
    if (br >= nbr) return;
    int nb = brp[br + 1] - brp[br];
    cnt[br] = nb <= SLICE ? 1u : (uint32_t)((nb + SLICE - 1) / SLICE);
}
__global__ void work_fill_kernel(const int32_t* __restrict__ brp, int nbr, const uint32_t* __restrict__ ofs, int4* __restrict__ work) {
    int br = blockIdx.x * blockDim.x + threadIdx.x;
    if (br >= nbr) return;
    int b0 = brp[br], b1 = brp[br + 1];
    uint32_t o = ofs[br], n = ofs[br + 1] - o;
    for (uint32_t i = 0; i < n; i++) {
        int s = b0 + (int)i * SLICE;
        work[o + i] = make_int4(br, s, min(s + SLICE, b1), n > 1 ? 1 : 0);
    }
}

int plan_spmv(bmsp_matrix_s* m, cudaStream_t st) {
    if (m->transposed || m->nbr == 0) { m->spmv_path = -1; return BMSP_OK; }
    const double per_blk = m->nblk ? (double)m->nnz / (double)m->nblk : 0.0;
    m->spmv_path = per_blk >= 2.5 ? 0 : 1;
    if (m->spmv_path == 0) {
        const int vsize = m->dtype == BMSP_F16 ? 2 : 4;
        double ab = (double)m->nblk / m->nbr * RT, av = (double)m->nnz / m->nbr * RT;
        int cb = (int)(ab * 1.25) + 32, cv = (int)(av * 1.25) + 128;
        cb = (cb + 3) & ~3; cv = (cv + 7) & ~7;
        const size_t budget = 48 * 1024;
        while (stage_bytes(cb, cv, vsize) > budget && (cb > 64 || cv > 256)) {
            cb = max(64, ((cb * 3 / 4) + 3) & ~3);
            cv = max(256, ((cv * 3 / 4) + 7) & ~7);
        }
        m->cap_blk = cb; m->cap_val = cv;
        return BMSP_OK;
    }
    uint32_t* cnt = nullptr;
    BMSP_TRY(dev_alloc_t(&cnt, (size_t)m->nbr + 1, st));
    work_count_kernel<<<(unsigned)ceil_div(m->nbr, 256), 256, 0, st>>>(m->brp, m->nbr, cnt);
    BMSP_KERNEL_CHECK();
    BMSP_TRY(exclusive_scan_u32(cnt, cnt, m->nbr, st));
    uint32_t total = 0;
    BMSP_CUDA(cudaMemcpyAsync(&total, cnt + m->nbr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    BMSP_CUDA(cudaStreamSynchronize(st));
    m->n_work = (int32_t)total;
    m->n_split = (int32_t)total - m->nbr;   // > 0 iff some block row is sliced
    BMSP_TRY(dev_alloc((void**)&m->work, sizeof(int4) * (size_t)total, st));
    work_fill_kernel<<<(unsigned)ceil_div(m->nbr, 256), 256, 0, st>>>(m->brp, m->nbr, cnt, (int4*)m->work);
    BMSP_KERNEL_CHECK();
    m->split_rows = (int32_t*)cnt;           // item offsets per block row, kept for the fix-up
    if (m->n_split > 0) BMSP_TRY(dev_alloc_t(&m->split_partial, (size_t)total * 8, st));
    return BMSP_OK;
}

template <typename T, typename X>
static int launch_spmv(bmsp_matrix_s* A, const X* x, float* y, cudaStream_t st) {
    if (A->spmv_path == 0) {
        SpmvArgs<T> a;
        a.bmps = A->bmps; a.bcol = A->bcol; a.values = (const T*)A->values; a.brp = A->brp; a.rvb = A->rvb;
        a.rows = A->rows; a.nbr = A->nbr; a.ntiles = (int)ceil_div(A->nbr, RT); a.cap_blk = A->cap_blk; a.cap_val = A->cap_val;
        const size_t smem = stage_bytes(a.cap_blk, a.cap_val, sizeof(T));
        static size_t configured[4] = {0, 0, 0, 0};
        const int inst = (sizeof(T) == 2 ? 0 : 1) * 2 + (sizeof(X) == 2 ? 1 : 0);
        auto kern = spmv_rowtile_kernel<T, X>;
        if (configured[inst] < smem) {
            BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            configured[inst] = smem;
        }
        const int grid = a.ntiles;
        kern<<<grid, SPMV_THREADS, smem, st>>>(a, x, y);
        BMSP_KERNEL_CHECK();
        return BMSP_OK;
    }
    spmv_blockpar_kernel<T, X><<<(unsigned)ceil_div(A->n_work, 8), 256, 0, st>>>(A->bmps, A->bcol, A->offsets, (const T*)A->values,
                                                                               (const int4*)A->work, A->n_work, A->rows, x, y,
                                                                               A->split_partial);
    BMSP_KERNEL_CHECK();
    if (A->n_split > 0) {
        spmv_fixup_kernel<<<(unsigned)ceil_div(A->rows, 256), 256, 0, st>>>(A->split_rows, A->split_partial, A->nbr, A->rows, y);
        BMSP_KERNEL_CHECK();
    }
    return BMSP_OK;
}

}  // namespace bmsp

using namespace bmsp;

extern "C" int bmsp_spmv(bmsp_matrix_t A, const void* x, int32_t x_dtype, float* y, void* stream) {
    if (!A || !x || !y || (x_dtype != BMSP_F16 && x_dtype != BMSP_F32)) { set_error("bmsp_spmv: invalid argument"); return BMSP_ERR_INVALID; }
    if (A->transposed) { set_error("bmsp_spmv: matrix is in transposed-operand form"); return BMSP_ERR_UNSUPPORTED; }
    if (A->rows == 0) return BMSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (A->dtype == BMSP_F16)
        return x_dtype == BMSP_F32 ? launch_spmv<__half, float>(A, (const float*)x, y, st) : launch_spmv<__half, __half>(A, (const __half*)x, y, st);
    return x_dtype == BMSP_F32 ? launch_spmv<float, float>(A, (const float*)x, y, st) : launch_spmv<float, __half>(A, (const __half*)x, y, st);
}

// nblk*(4+8) + nbr*(4+4) + nnz*sizeof(value) + ncols*sizeof(x) + nrows*4
extern "C" int bmsp_spmv_bytes(bmsp_matrix_t A, int32_t x_dtype, int64_t* bytes) {
    if (!A || !bytes) { set_error("bmsp_spmv_bytes: null argument"); return BMSP_ERR_INVALID; }
    *bytes = A->nblk * 12 + (int64_t)A->nbr * 8 + A->nnz * (A->dtype == BMSP_F16 ? 2 : 4) +
             (int64_t)A->cols * (x_dtype == BMSP_F16 ? 2 : 4) + (int64_t)A->rows * 4;
    return BMSP_OK;
}
