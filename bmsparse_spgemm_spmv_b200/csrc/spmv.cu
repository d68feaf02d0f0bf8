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
constexpr int SPMV_THREADS = 256;  // 2 matrix rows per thread
constexpr int STAGES = 4;
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
    return (size_t)(cap_blk + 4) * 8 + (size_t)(cap_blk + 4) * 4 + (size_t)(cap_val + 8) * vsize + 2 * ROWSLOT * 4 + 16;
}

template <typename X> __device__ __forceinline__ float ld_x(const X* x, int64_t i);
template <> __device__ __forceinline__ float ld_x<float>(const float* x, int64_t i) { return __ldg(x + i); }
template <> __device__ __forceinline__ float ld_x<__half>(const __half* x, int64_t i) { return __half2float(__ldg(x + i)); }

template <typename T, typename X>
__global__ void __launch_bounds__(SPMV_THREADS) spmv_rowtile_kernel(SpmvArgs<T> a, const X* __restrict__ x, float* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VA = 16 / sizeof(T);   // values per 16 bytes
    const size_t sb = stage_bytes(a.cap_blk, a.cap_val, sizeof(T));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sb * STAGES);
    const int tid = threadIdx.x;

    auto st_bm   = [&](int s) { return reinterpret_cast<uint64_t*>(smem + sb * s); };
    auto st_bc   = [&](int s) { return reinterpret_cast<int32_t*>(smem + sb * s + (size_t)(a.cap_blk + 4) * 8); };
    auto st_val  = [&](int s) { return reinterpret_cast<T*>(smem + sb * s + (size_t)(a.cap_blk + 4) * 12); };
    auto st_brp  = [&](int s) { return reinterpret_cast<int32_t*>(smem + sb * s + (size_t)(a.cap_blk + 4) * 12 + (size_t)(a.cap_val + 8) * sizeof(T)); };
    auto st_rvb  = [&](int s) { return reinterpret_cast<uint32_t*>(st_brp(s) + ROWSLOT); };
    auto st_meta = [&](int s) { return reinterpret_cast<TileMeta*>(st_brp(s) + 2 * ROWSLOT); };

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // producer (thread 0): stage tile t into ring slot s
    auto issue = [&](int t, int s) {
        const int r0 = t * RT, r1 = min(r0 + RT, a.nbr);
        const int p0 = a.brp[r0], p1 = a.brp[r1];
        const uint32_t v0 = a.rvb[r0], v1 = a.rvb[r1];
        TileMeta m;
        m.p0a = p0 & ~1; m.p0c = p0 & ~3; m.v0a = v0 & ~(uint32_t)(VA - 1);
        m.staged = (p1 - p0 <= a.cap_blk) && ((int64_t)v1 - v0 <= a.cap_val);
        *st_meta(s) = m;
        const uint32_t nrow = (uint32_t)(((r1 - r0 + 1) + 3) & ~3) * 4;
        uint32_t n8 = 0, n4 = 0, nv = 0;
        if (m.staged) {
            n8 = (uint32_t)((p1 - m.p0a + 1) & ~1) * 8;
            n4 = (uint32_t)((p1 - m.p0c + 3) & ~3) * 4;
            nv = (uint32_t)((v1 - m.v0a + VA - 1) & ~(uint32_t)(VA - 1)) * sizeof(T);
        }
        mbar_arrive_expect_tx(&bars[s], n8 + n4 + nv + 2 * nrow);
        bulk_g2s(st_brp(s), a.brp + r0, nrow, &bars[s]);
        bulk_g2s(st_rvb(s), a.rvb + r0, nrow, &bars[s]);
        if (n8) bulk_g2s(st_bm(s), a.bmps + m.p0a, n8, &bars[s]);
        if (n4) bulk_g2s(st_bc(s), a.bcol + m.p0c, n4, &bars[s]);
        if (nv) bulk_g2s(st_val(s), a.values + m.v0a, nv, &bars[s]);
    };

    const int first = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            int t = first + s * stride;
            if (t < a.ntiles) issue(t, s);
        }
    }

    // per-thread constants: row inside the 8x8 block
    const int ri = tid & 7;
    const int sh = 24 - 8 * (ri & 3);                       // shift of my row byte inside its 32-bit half
    const uint32_t premask = ~(0xFFFFFFFFu >> (8 * (ri & 3)));   // bits of earlier rows in the same half (0 for ri&3==0)
    const bool lowhalf = ri >= 4;

    int it = 0;
    for (int t = first; t < a.ntiles; t += stride, it++) {
        const int s = it % STAGES;
        mbar_wait(&bars[s], (uint32_t)((it / STAGES) & 1));
        const TileMeta m = *st_meta(s);
        const uint64_t* bm = m.staged ? st_bm(s) : a.bmps + m.p0a;
        const int32_t* bc = m.staged ? st_bc(s) : a.bcol + m.p0c;
        const T* vals = m.staged ? st_val(s) : a.values + m.v0a;
        const int32_t* s_brp = st_brp(s);
        const uint32_t* s_rvb = st_rvb(s);
        const int r0 = t * RT;
#pragma unroll
        for (int j = 0; j < RT * 8 / SPMV_THREADS; j++) {
            const int rowi = tid + j * SPMV_THREADS;
            const int lbr = rowi >> 3;
            const int64_t row = (int64_t)r0 * 8 + rowi;
            if (row >= a.rows) continue;
            const int pb = s_brp[lbr], pe = s_brp[lbr + 1];
            uint32_t k = s_rvb[lbr] - m.v0a;     // index of the block's first value in `vals`
            float acc = 0.f;
            for (int b = pb; b < pe; b++) {
                const uint64_t bmp = bm[b - m.p0a];
                const uint32_t hi = (uint32_t)(bmp >> 32), lo = (uint32_t)bmp;
                const uint32_t w = lowhalf ? lo : hi;
                uint32_t byte = (w >> sh) & 0xFFu;
                const uint32_t nhi = __popc(hi);
                if (byte) {
                    uint32_t kk = k + __popc(w & premask) + (lowhalf ? nhi : 0u);
                    const int64_t xb = (int64_t)bc[b - m.p0c] * 8;
                    do {
                        const int c = __clz(byte) - 24;          // MSB of the byte is column 0
                        byte &= ~(0x80u >> c);
                        acc = fmaf(val_to_f32(vals[kk]), ld_x<X>(x, xb + c), acc);
                        kk++;
                    } while (byte);
                }
                k += nhi + __popc(lo);
            }
            y[row] = acc;
        }
        __syncthreads();                      // every thread is done with slot s
        if (tid == 0) {
            const int tn = t + STAGES * stride;
            if (tn < a.ntiles) issue(tn, s);
        }
    }
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
        const int64_t xb = valid ? (int64_t)ld_stream_s32(bcol + b) * 8 : 0;
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
            acc[p >> 3][lane] += val_to_f32(values[k]) * ld_x<X>(x, xb + (p & 7));
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
        int cb = (int)(ab * 1.5) + 64, cv = (int)(av * 1.5) + 256;
        cb = (cb + 3) & ~3; cv = (cv + 7) & ~7;
        const size_t budget = 40 * 1024;
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
        const size_t smem = stage_bytes(a.cap_blk, a.cap_val, sizeof(T)) * STAGES + STAGES * 8;
        static int sms = 0;
        static size_t configured[4] = {0, 0, 0, 0};
        const int inst = (sizeof(T) == 2 ? 0 : 1) * 2 + (sizeof(X) == 2 ? 1 : 0);
        auto kern = spmv_rowtile_kernel<T, X>;
        if (configured[inst] < smem) {
            BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[inst] = smem;
        }
        if (!sms) { int dev; BMSP_CUDA(cudaGetDevice(&dev)); BMSP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); }
        int occ = 0;
        BMSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SPMV_THREADS, smem));
        if (occ < 1) { set_error("spmv: kernel does not fit (smem %zu)", smem); return BMSP_ERR_CUDA; }
        int grid = min(a.ntiles, sms * occ);
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
