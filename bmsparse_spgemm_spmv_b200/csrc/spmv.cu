// spmv.cu -- y = A x on the bmSparse form, HBM-streaming, no tensor cores.
//
// Replaces bmSparse_SpMV / spmv_kernel (src/bmSparse_SPMV.cu:153-230).  The reference rebuilds the
// block-row pointers with reduce_by_key + scan on every call (:196-206), runs one 64-thread CTA per
// block row with thread <-> cell and re-reads key/bitmap/offset (24 B) from all 64 threads per block.
// Here the matrix is streamed exactly once through the compact surface (8 B bitmap + 4 B block column
// per block, 2 B per value, 8 B per block row):
//
//   path 0 "row-tiled"  (dense-ish blocks: Poisson, block-clustered)
//       one CTA (128 threads) per tile of 64 block rows; the tile's bitmaps / block columns / values / row
//       pointers are bulk-copied (cp.async.bulk + mbarrier, SASS UBLKCP) into shared memory; 16 CTAs are
//       resident per SM, each in a different phase, which is what hides the copy latency.  Two threads per
//       block row, one per 32-bit bitmap half (4 matrix rows, 4 fp32 accumulators): each walks the blocks of
//       its block row, ranks with popc, and consumes set bits in "rounds" whose loads are predicated (no
//       branches) so several are in flight; x is read through the read-only path (L1/L2); y is stored as
//       one float4 per thread.
//   path 1 "block-parallel" (about one value per block: uniform random, R-MAT)
//       one warp per work item (a block row, or a 4096-block slice of a long one); lane <-> block,
//       coalesced 8 B + 4 B metadata loads, value offsets by a warp scan of popc, eight per-row partial
//       sums per lane kept in shared memory, shuffle reduction at the end; sliced rows are finished by
//       a deterministic fix-up kernel.
#include "common.cuh"
#include <cstdlib>

namespace bmsp {

constexpr int RT = 64;             // block rows per tile (path 0)
constexpr int SPMV_THREADS = 2 * RT;   // two threads per block row: one per 32-bit bitmap half (4 matrix rows each)
constexpr int SLICE = 4096;        // blocks per work item (path 1)
constexpr int ROWSLOT = RT + 4;    // staged row-pointer slice, padded to a 16-byte multiple

template <typename T>
struct SpmvArgs {
    const uint64_t* bmps; const int32_t* bcol; const T* values;
    const int32_t* brp; const uint32_t* rvb;
    int32_t rows, nbr, ntiles, cap_blk, cap_val;
};

__host__ __device__ inline size_t stage_bytes(int cap_blk, int cap_val, int vsize) {
    return (size_t)(cap_blk + 4) * 8 + (size_t)(cap_blk + 4) * 4 + (size_t)(cap_val + 8) * vsize + 2 * ROWSLOT * 4 + 16 + 16;   // + x slots, added by the launcher
}

template <typename X> __device__ __forceinline__ float ld_xp(const X* p);
template <> __device__ __forceinline__ float ld_xp<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_xp<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <typename X> __device__ __forceinline__ float ld_x(const X* x, uint32_t i);
template <> __device__ __forceinline__ float ld_x<float>(const float* x, uint32_t i) { return __ldg(x + i); }
template <> __device__ __forceinline__ float ld_x<__half>(const __half* x, uint32_t i) { return __half2float(__ldg(x + i)); }

// One 8-bit row mask, MSB-aligned in `t` (column 0 = bit 31): walk its set bits in column order; the
// values are consecutive from vp.  xq points 24 elements before the block's x segment so that the
// leading-zero count of the right-aligned form indexes it directly.
template <typename T, typename X>
__device__ __forceinline__ void row_bits(uint32_t t, const T*& vp, const X* __restrict__ xs, float& acc) {
    while (t) {
        uint32_t c;                                      // column = leading zeros (FLO.U32.SH)
        asm("bfind.shiftamt.u32 %0, %1;" : "=r"(c) : "r"(t));
        const X* px;                                     // xs + c in one IMAD.WIDE
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(px) : "r"(c), "n"(sizeof(X)), "l"(xs));
        acc = fmaf(val_to_f32(*vp), ld_xp<X>(px), acc);
        vp++;
        t ^= 0x80000000u >> c;
    }
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// The blocks [pb, pe) of one block row, seen by the thread that owns bitmap half `h` (rows 4h..4h+3).
// bm / bc / vals are indexed relative to the tile's aligned bases (shared memory for staged tiles).
template <typename T, typename X>
__device__ __forceinline__ void half_block_row(const uint64_t* __restrict__ bm, const int32_t* __restrict__ bc,
                                               const T* __restrict__ vals, int pb, int pe, uint32_t k, const int h,
                                               const X* __restrict__ x, float (&acc)[4]) {
    if (pb < pe) prefetch_l1(x + (uint32_t)bc[pb] * 8u + 4u * h);
    for (int b = pb; b < pe; b++) {
        const uint2 w2 = *reinterpret_cast<const uint2*>(bm + b);     // .y = rows 0-3, .x = rows 4-7
        const uint32_t nhi = __popc(w2.y), nlo = __popc(w2.x);
        const uint32_t w = h ? w2.x : w2.y;
        const T* vp = vals + (k + (h ? nhi : 0u));
        k += nhi + nlo;
        // the next block's x segment is touched one block ahead: its L2 latency hides behind this block
        if (b + 1 < pe) prefetch_l1(x + (uint32_t)bc[b + 1] * 8u + 4u * h);
        if (w) {
            const X* xs = x + (uint32_t)bc[b] * 8u;
            row_bits<T, X>(w & 0xFF000000u, vp, xs, acc[0]);
            row_bits<T, X>((w << 8) & 0xFF000000u, vp, xs, acc[1]);
            row_bits<T, X>((w << 16) & 0xFF000000u, vp, xs, acc[2]);
            row_bits<T, X>(w << 24, vp, xs, acc[3]);
        }
    }
}

// ---- shared-memory fast path: explicit 32-bit shared addresses (no generic-pointer arithmetic) ----
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_val(uint32_t a, __half) {
    unsigned short u;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u) : "r"(a));
    return __half2float(__ushort_as_half(u));
}
__device__ __forceinline__ float lds_val(uint32_t a, float) { return lds_f32(a); }
__device__ __forceinline__ float lds_x(uint32_t a, __half h) { return lds_val(a, h); }
__device__ __forceinline__ float lds_x(uint32_t a, float) { return lds_f32(a); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// One "round" = the next set bit (ascending column) of each of the four 8-bit row masks of a bitmap half.
// row_issue does the bit bookkeeping and issues the two loads of one row PREDICATED (no branches), so the
// eight loads of a round are in flight together; the FMAs follow.  b is the row mask MSB-aligned, va the
// shared address of the row's next value, xs31 = the block's x segment + 31 elements (the found bit
// position p indexes it as xs31 - p).  SASS: FLO, BMSK, IMAD.WIDE, @P LDG, @P LDS, LOP3, @P IADD.
template <typename T, typename X> struct RowLoad;
template <> struct RowLoad<__half, float> {
    float xv; unsigned short raw;
    __device__ __forceinline__ void issue(uint32_t& b, uint32_t& va, const float* xs31) {
        asm volatile("{\n\t.reg .pred pq;\n\t.reg .u32 pp, mm;\n\t.reg .u64 ad;\n\t"
                     "setp.ne.u32 pq, %2, 0;\n\tbfind.u32 pp, %2;\n\tbmsk.clamp.b32 mm, pp, 1;\n\tmad.wide.s32 ad, pp, -4, %4;\n\t"
                     "@pq ld.global.nc.f32 %0, [ad];\n\t@pq ld.shared.u16 %1, [%3];\n\t"
                     "not.b32 mm, mm;\n\tand.b32 %2, %2, mm;\n\t@pq add.u32 %3, %3, 2;\n\t}"
                     : "=f"(xv), "=h"(raw), "+r"(b), "+r"(va) : "l"(xs31));
    }
    __device__ __forceinline__ float prod_add(float acc) const { return fmaf(__half2float(__ushort_as_half(raw)), xv, acc); }
};
template <> struct RowLoad<__half, __half> {
    unsigned short xr, raw;
    __device__ __forceinline__ void issue(uint32_t& b, uint32_t& va, const __half* xs31) {
        asm volatile("{\n\t.reg .pred pq;\n\t.reg .u32 pp, mm;\n\t.reg .u64 ad;\n\t"
                     "setp.ne.u32 pq, %2, 0;\n\tbfind.u32 pp, %2;\n\tbmsk.clamp.b32 mm, pp, 1;\n\tmad.wide.s32 ad, pp, -2, %4;\n\t"
                     "@pq ld.global.nc.u16 %0, [ad];\n\t@pq ld.shared.u16 %1, [%3];\n\t"
                     "not.b32 mm, mm;\n\tand.b32 %2, %2, mm;\n\t@pq add.u32 %3, %3, 2;\n\t}"
                     : "=h"(xr), "=h"(raw), "+r"(b), "+r"(va) : "l"(xs31));
    }
    __device__ __forceinline__ float prod_add(float acc) const {
        return fmaf(__half2float(__ushort_as_half(raw)), __half2float(__ushort_as_half(xr)), acc);
    }
};
template <> struct RowLoad<float, float> {
    float xv, raw;
    __device__ __forceinline__ void issue(uint32_t& b, uint32_t& va, const float* xs31) {
        asm volatile("{\n\t.reg .pred pq;\n\t.reg .u32 pp, mm;\n\t.reg .u64 ad;\n\t"
                     "setp.ne.u32 pq, %2, 0;\n\tbfind.u32 pp, %2;\n\tbmsk.clamp.b32 mm, pp, 1;\n\tmad.wide.s32 ad, pp, -4, %4;\n\t"
                     "@pq ld.global.nc.f32 %0, [ad];\n\t@pq ld.shared.f32 %1, [%3];\n\t"
                     "not.b32 mm, mm;\n\tand.b32 %2, %2, mm;\n\t@pq add.u32 %3, %3, 4;\n\t}"
                     : "=f"(xv), "=f"(raw), "+r"(b), "+r"(va) : "l"(xs31));
    }
    __device__ __forceinline__ float prod_add(float acc) const { return fmaf(raw, xv, acc); }
};
template <> struct RowLoad<float, __half> {
    unsigned short xr; float raw;
    __device__ __forceinline__ void issue(uint32_t& b, uint32_t& va, const __half* xs31) {
        asm volatile("{\n\t.reg .pred pq;\n\t.reg .u32 pp, mm;\n\t.reg .u64 ad;\n\t"
                     "setp.ne.u32 pq, %2, 0;\n\tbfind.u32 pp, %2;\n\tbmsk.clamp.b32 mm, pp, 1;\n\tmad.wide.s32 ad, pp, -2, %4;\n\t"
                     "@pq ld.global.nc.u16 %0, [ad];\n\t@pq ld.shared.f32 %1, [%3];\n\t"
                     "not.b32 mm, mm;\n\tand.b32 %2, %2, mm;\n\t@pq add.u32 %3, %3, 4;\n\t}"
                     : "=h"(xr), "=f"(raw), "+r"(b), "+r"(va) : "l"(xs31));
    }
    __device__ __forceinline__ float prod_add(float acc) const { return fmaf(raw, __half2float(__ushort_as_half(xr)), acc); }
};

// The blocks of one block row seen by the thread that owns bitmap half h (rows 4h..4h+3); bitmaps and
// values in shared memory (32-bit shared addresses), x through the read-only path.
template <typename T, typename X, int RPR>
__device__ __forceinline__ void half_block_row_s(uint32_t a_bm, uint32_t a_bc, uint32_t a_val, int nb, uint32_t k, const int h,
                                                 const X* __restrict__ x, float (&acc)[4]) {
    const uint32_t hm = 0u - (uint32_t)h;
    for (int i = 0; i < nb; i++) {
        const uint2 w2 = lds_v2(a_bm);                  // .y = rows 0-3, .x = rows 4-7
        a_bm += 8;
        const uint32_t nhi = __popc(w2.y), nlo = __popc(w2.x);
        const uint32_t w = h ? w2.x : w2.y;
        const uint32_t va = a_val + (k + (nhi & hm)) * (uint32_t)sizeof(T);
        k += nhi + nlo;
        if (w) {
            const X* xs31 = x + (size_t)lds_u32(a_bc) * 8 + 31;
            uint32_t b0 = w & 0xFF000000u, b1 = (w << 8) & 0xFF000000u, b2 = (w << 16) & 0xFF000000u, b3 = w << 24;
            uint32_t va0 = va, va1 = va0 + __popc(b0) * (uint32_t)sizeof(T), va2 = va1 + __popc(b1) * (uint32_t)sizeof(T),
                     va3 = va2 + __popc(b2) * (uint32_t)sizeof(T);
            if (RPR == 4) {
                do {
                    const bool q0 = b0 != 0, q1 = b1 != 0, q2 = b2 != 0, q3 = b3 != 0;
                    RowLoad<T, X> l0, l1, l2, l3;
                    l0.issue(b0, va0, xs31);
                    l1.issue(b1, va1, xs31);
                    l2.issue(b2, va2, xs31);
                    l3.issue(b3, va3, xs31);
                    if (q0) acc[0] = l0.prod_add(acc[0]);
                    if (q1) acc[1] = l1.prod_add(acc[1]);
                    if (q2) acc[2] = l2.prod_add(acc[2]);
                    if (q3) acc[3] = l3.prod_add(acc[3]);
                } while (b0 | b1 | b2 | b3);
            } else {
                while (b0 | b1) {
                    const bool q0 = b0 != 0, q1 = b1 != 0;
                    RowLoad<T, X> l0, l1;
                    l0.issue(b0, va0, xs31);
                    l1.issue(b1, va1, xs31);
                    if (q0) acc[0] = l0.prod_add(acc[0]);
                    if (q1) acc[1] = l1.prod_add(acc[1]);
                }
                while (b2 | b3) {
                    const bool q2 = b2 != 0, q3 = b3 != 0;
                    RowLoad<T, X> l2, l3;
                    l2.issue(b2, va2, xs31);
                    l3.issue(b3, va3, xs31);
                    if (q2) acc[2] = l2.prod_add(acc[2]);
                    if (q3) acc[3] = l3.prod_add(acc[3]);
                }
            }
        }
        a_bc += 4;
    }
}

struct TileMeta { int32_t p0, p1; uint32_t v0a; int32_t staged; };

// One CTA per tile, a single staging buffer: latency is hidden by the CTAs resident per SM, each in a
// different phase (row pointers -> bulk copies in flight -> x gather in flight -> compute -> store).
template <typename T, typename X, int MINB, int RPR>
__global__ void __launch_bounds__(SPMV_THREADS, MINB) spmv_rowtile_kernel(SpmvArgs<T> a, const X* __restrict__ x, float* __restrict__ y, int ncols) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VA = 16 / sizeof(T);   // values per 16 bytes
    const int tid = threadIdx.x;
    const uint32_t off_bc = (uint32_t)(a.cap_blk + 4) * 8, off_val = (uint32_t)(a.cap_blk + 4) * 12;
    const uint32_t off_brp = off_val + (uint32_t)(a.cap_val + 8) * sizeof(T);
    const uint32_t off_rvb = off_brp + ROWSLOT * 4, off_meta = off_rvb + ROWSLOT * 4, off_bar = off_meta + 16;
    TileMeta* s_meta = reinterpret_cast<TileMeta*>(smem + off_meta);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + off_bar);
    const uint32_t sbase = smem_u32(smem);

    const int t = blockIdx.x;
    const int r0 = t * RT, r1 = min(r0 + RT, a.nbr);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        const int p0 = a.brp[r0], p1 = a.brp[r1];
        const uint32_t v0 = a.rvb[r0], v1 = a.rvb[r1];
        TileMeta m;
        m.p0 = p0; m.p1 = p1; m.v0a = v0 & ~(uint32_t)(VA - 1);
        m.staged = (p1 - p0 <= a.cap_blk) && ((int64_t)v1 - v0 <= a.cap_val);
        *s_meta = m;
        const int p0a = p0 & ~1, p0c = p0 & ~3;
        const uint32_t nrow = (uint32_t)(((r1 - r0 + 1) + 3) & ~3) * 4;
        uint32_t n8 = 0, n4 = 0, nv = 0;
        if (m.staged) {
            n8 = (uint32_t)((p1 - p0a + 1) & ~1) * 8;
            n4 = (uint32_t)((p1 - p0c + 3) & ~3) * 4;
            nv = (uint32_t)((v1 - m.v0a + VA - 1) & ~(uint32_t)(VA - 1)) * sizeof(T);
        }
        mbar_arrive_expect_tx(bar, n8 + n4 + nv + 2 * nrow);
        bulk_g2s(smem + off_brp, a.brp + r0, nrow, bar);
        bulk_g2s(smem + off_rvb, a.rvb + r0, nrow, bar);
        if (n8) bulk_g2s(smem, a.bmps + p0a, n8, bar);
        if (n4) bulk_g2s(smem + off_bc, a.bcol + p0c, n4, bar);
        if (nv) bulk_g2s(smem + off_val, a.values + m.v0a, nv, bar);
    }
    __syncthreads();                 // barrier initialised and armed before anyone polls it
    mbar_wait(bar, 0);

    const TileMeta m = *s_meta;
    const int lbr = tid >> 1, h = tid & 1;
    const int64_t row = ((int64_t)(r0 + lbr)) * 8 + h * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (m.staged) {
        if (row >= a.rows) return;
        const uint32_t pb = lds_u32(sbase + off_brp + 4u * lbr), pe = lds_u32(sbase + off_brp + 4u * lbr + 4u);
        const uint32_t k = lds_u32(sbase + off_rvb + 4u * lbr) - m.v0a;
        const uint32_t rel = pb - (uint32_t)m.p0;
        half_block_row_s<T, X, RPR>(sbase + ((uint32_t)(m.p0 & 1) + rel) * 8u, sbase + off_bc + ((uint32_t)(m.p0 & 3) + rel) * 4u, sbase + off_val,
                               (int)(pe - pb), k, h, x, acc);
    } else {
        if (row >= a.rows) return;
        const int32_t* s_brp = reinterpret_cast<const int32_t*>(smem + off_brp);
        const uint32_t* s_rvb = reinterpret_cast<const uint32_t*>(smem + off_rvb);
        half_block_row<T, X>(a.bmps, a.bcol, a.values + m.v0a, s_brp[lbr], s_brp[lbr + 1], s_rvb[lbr] - m.v0a, h, x, acc);
    }
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
        int cb = (int)(ab * 1.125) + 16, cv = (int)(av * 1.125) + 64;
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
        static int variant = -1;
        if (variant < 0) { const char* e = getenv("BMSP_SPMV_VARIANT"); variant = e ? atoi(e) : 0; }
        // default: 16 CTAs/SM (32 registers), two rows per round.  BMSP_SPMV_VARIANT=1 selects the four-rows-per-round
        // build (40 registers, 12 CTAs/SM) for experiments; measured on P4096: 106.7 us vs 112.8 us.
        void (*kern)(SpmvArgs<T>, const X*, float*, int) = spmv_rowtile_kernel<T, X, 16, 2>;
        if (variant == 1) kern = spmv_rowtile_kernel<T, X, 12, 4>;
        static size_t configured = 0;
        if (configured < smem) {
            BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            configured = smem;
        }
        const int grid = a.ntiles;
        kern<<<grid, SPMV_THREADS, smem, st>>>(a, x, y, A->cols);
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
    if (((uintptr_t)x & 15) || ((uintptr_t)y & 15)) { set_error("bmsp_spmv: x and y must be 16-byte aligned"); return BMSP_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    if (A->spmv_path < 0) BMSP_TRY(plan_spmv(A, st));
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
