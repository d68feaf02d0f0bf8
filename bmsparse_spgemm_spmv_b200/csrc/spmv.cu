// spmv.cu -- y = A x on the bmSparse form, HBM-streaming, no tensor cores.
//
// Replaces bmSparse_SpMV / spmv_kernel (src/bmSparse_SPMV.cu:153-230).  The reference rebuilds the
// block-row pointers with reduce_by_key + scan on every call (:196-206), runs one 64-thread CTA per
// block row with thread <-> cell and re-reads key/bitmap/offset (24 B) from all 64 threads per block.
// Here the matrix is streamed exactly once through the compact surface (8 B bitmap + 4 B block column
// per block, 2 B per value, 8 B per block row):
//
//   path 0 "row-tiled"  (dense-ish blocks: Poisson, block-clustered)
//       one CTA per tile of 64 (32, 16 for heavy rows) block rows.  A per-matrix tile plan lists, for every tile, the
//       distinct 32-column lines of x its blocks touch and gives every block a 16-bit offset into the tile's staged
//       copy of those lines.  The tile's bitmaps / x offsets / values / row pointers are bulk-copied (cp.async.bulk +
//       mbarrier, SASS UBLKCP) into shared memory while all threads gather the x lines with coalesced 16-byte loads;
//       12+ CTAs are resident per SM, each in a different phase, which is what hides the copy latency.  Two threads
//       per block row, one per 32-bit bitmap half (4 matrix rows, 4 fp32 accumulators): each walks the blocks of its
//       block row, ranks with popc, and consumes set bits two rows at a time in predicated rounds (one PTX block per
//       row pair, no branches except the early exit); every per-value read (fp16 value, fp32 x) is a shared-memory
//       load; y is stored as one float4 per thread.  Tiles that exceed the shared-memory capacities are read from
//       global memory by the same threads.
//   path 1 "block-parallel" (about one value per block: uniform random, R-MAT)
//       one warp per work item (a block row, or a 512-block slice of a long one); lane <-> block, 4 x 32 blocks per step,
//       coalesced 8 B + 4 B metadata loads, value offsets by a warp scan of popc, eight per-row partial
//       sums per lane kept in shared memory, shuffle reduction at the end; sliced rows are finished by
//       a deterministic fix-up kernel.
#include "common.cuh"
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cstring>
#include <utility>
#include <type_traits>

namespace bmsp {

constexpr int SLICE = 512;         // blocks per work item (path 1): 4 steps of 128 blocks; longer block rows are sliced and summed by the fix-up

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

// ---- explicit 32-bit shared addresses (no generic-pointer arithmetic) ----
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

// ===================================================================================== path 0, tile plan
// Built once per matrix (plan_spmv): for every tile of RT block rows, the sorted distinct x "lines" (32 columns
// = 4 block columns = 128 B of fp32 x) its blocks touch, and per block a 16-bit offset of its 8-column x segment
// inside the tile's staged copy of those lines.  The kernel then never reads block_col: it streams bitmaps
// (8 B), x offsets (2 B) and values per block, gathers each distinct x line ONCE per tile with coalesced
// 16-byte loads (8 threads per line) and serves every per-value x read from shared memory.  Staged lines
// have a pitch of 33 elements: when the 16 block rows of a warp read the same column of consecutive block
// columns (stencils, bands) the 16 addresses fall into 16 different banks.
constexpr int XL_STRIDE = 33;      // staged x line pitch, elements (32 + 1 skew)
constexpr int PLAN_MAXB = 2048;    // blocks per tile the planner sorts in shared memory
constexpr int XL_MAX = 1900;       // distinct lines per tile (16-bit element offsets)

// 64 bytes; [lmin, lmax] = x lines touched.  The distinct lines of a stencil / band / cluster tile form a few runs of consecutive
// lines: up to three runs (first line, line count) are kept in the descriptor so that the streaming kernel can fetch the tile's
// x with one bulk copy per run without reading the line list (nruns = 0: more than three runs, copy line by line).
struct __align__(16) TileDesc { int32_t p0, nb; uint32_t v0; int32_t nv, nl, flags, lmin, lmax; int32_t nruns; uint32_t run_line[3], run_count[3]; int32_t pad; };
static_assert(sizeof(TileDesc) == 64, "TileDesc is four 16-byte words");

__global__ void __launch_bounds__(256) tile_plan_kernel(const int32_t* __restrict__ brp, const int32_t* __restrict__ bcol,
                                                        const uint32_t* __restrict__ rvb, int nbr, int ncols, int rt, int pitch, TileDesc* __restrict__ desc,
                                                        uint32_t* __restrict__ lines, uint16_t* __restrict__ xoff,
                                                        unsigned long long* __restrict__ stats) {
    __shared__ uint32_t s_key[PLAN_MAXB];
    __shared__ uint32_t s_uniq[PLAN_MAXB];
    __shared__ uint32_t s_w[9];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r0 = t * rt, r1 = min(r0 + rt, nbr);
    const int p0 = brp[r0], nb = brp[r1] - p0;
    const uint32_t v0 = rvb[r0], v1 = rvb[r1];
    TileDesc d;
    d.p0 = p0; d.nb = nb; d.v0 = v0; d.nv = (int32_t)(v1 - v0); d.nl = 0; d.flags = 0; d.lmin = 0; d.lmax = nb > 0 ? 0x7FFFFFFF : -1;
    d.nruns = 0; d.pad = 0;
    for (int k = 0; k < 3; k++) { d.run_line[k] = 0; d.run_count[k] = 0; }
    if (nb > PLAN_MAXB) {          // too many blocks to plan: the kernel reads this tile straight from global memory
        if (tid == 0) { desc[t] = d; atomicAdd(stats + 4, 1ull); }
        return;
    }
    int n2 = 32;
    while (n2 < nb) n2 <<= 1;
    for (int i = tid; i < n2; i += 256) s_key[i] = i < nb ? ((uint32_t)bcol[p0 + i] >> 2) : 0xFFFFFFFFu;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += 256) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = s_key[i], b = s_key[ixj];
                    if ((a > b) == ((i & k) == 0)) { s_key[i] = b; s_key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // distinct keys, order kept: thread t owns `per` consecutive sorted slots
    const int per = n2 >= 256 ? n2 / 256 : 1;
    const int i0 = tid * per, i1 = min(i0 + per, nb);
    uint32_t cnt = 0;
    for (int i = i0; i < i1; i++) cnt += (i == 0 || s_key[i] != s_key[i - 1]) ? 1u : 0u;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int w = 0; w < 8; w++) { const uint32_t c = s_w[w]; s_w[w] = run; run += c; }
        s_w[8] = run;
    }
    __syncthreads();
    uint32_t run = s_w[wid] + inc - cnt;
    for (int i = i0; i < i1; i++)
        if (i == 0 || s_key[i] != s_key[i - 1]) s_uniq[run++] = s_key[i];
    const int nl = (int)s_w[8];
    __syncthreads();
    for (int j = tid; j < nl; j += 256) lines[p0 + j] = s_uniq[j];
    const bool ok = nl <= XL_MAX;
    if (ok)
        for (int i = tid; i < nb; i += 256) {
            const uint32_t c = (uint32_t)bcol[p0 + i], line = c >> 2;
            int lo = 0, hi = nl;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_uniq[mid] < line) lo = mid + 1; else hi = mid;
            }
            xoff[p0 + i] = (uint16_t)(lo * pitch + (int)(c & 3u) * 8);
        }
    if (tid == 0) {
        // flags: bit 0 = planned (x offsets valid), bit 1 = the last line reaches past the last column (guarded gather)
        if (nb > 0) { d.lmin = (int32_t)s_uniq[0]; d.lmax = (int32_t)s_uniq[nl - 1]; }
        d.nl = nl; d.flags = (ok ? 1 : 0) | ((nl > 0 && (uint64_t)s_uniq[nl - 1] * 32u + 32u > (uint64_t)ncols) ? 2 : 0);
        // runs of consecutive lines (at most three are kept; a fourth makes the tile a line-by-line one)
        int nruns = 0;
        for (int j = 0; j < nl && nruns <= 3; j++) {
            if (j == 0 || s_uniq[j] != s_uniq[j - 1] + 1u) { if (nruns < 3) { d.run_line[nruns] = s_uniq[j]; d.run_count[nruns] = 0; } nruns++; }
            if (nruns <= 3) d.run_count[nruns - 1]++;
        }
        d.nruns = nruns <= 3 ? nruns : 0;
        desc[t] = d;
        atomicMax(stats + 0, (unsigned long long)nb);
        atomicMax(stats + 1, (unsigned long long)(v1 - v0));
        atomicMax(stats + 2, (unsigned long long)nl);
        atomicAdd(stats + 3, (unsigned long long)nl);
        if (!ok) atomicAdd(stats + 4, 1ull);
    }
}

// shared-memory carve-up of spmv_tile_kernel (cap_blk and cap_val are multiples of 8); computed on the host
struct TileSmem { uint32_t xo, val, row, xs, bar, total; };
inline TileSmem tile_smem(int rt, int cap_blk, int cap_val, int cap_lines, int vsize, int xsize) {
    TileSmem s;
    s.xo = (uint32_t)(cap_blk + 2) * 8;
    s.val = s.xo + (uint32_t)(cap_blk + 16) * 2;
    s.row = s.val + (uint32_t)(cap_val + 16) * vsize;
    s.xs = s.row + (uint32_t)(rt + 2) * 8;
    s.bar = (s.xs + (uint32_t)cap_lines * XL_STRIDE * xsize + 15u) & ~15u;
    s.total = s.bar + 16;
    return s;
}

template <typename T>
struct TileArgs {
    const uint64_t* bmps; const int32_t* bcol; const T* values; const int2* rowpair;   // rowpair[r] = (block_row_ptr[r], first value of row r)
    const TileDesc* desc; const uint32_t* lines; const uint16_t* xoff;
    int32_t rows, nbr, cols, cap_blk, cap_val, cap_lines;
    int32_t tile0;       // first tile of this launch (row-range launches of the host-buffer pipeline)
    const int32_t* tile_list;   // streaming kernel: when set, the launch's q-th tile is tile_list[q] (boundary / interior launches of the multi-GPU product)
    TileSmem so;
};

__global__ void zip_rows_kernel(const int32_t* __restrict__ brp, const uint32_t* __restrict__ rvb, int n, int2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_int2(brp[i], (int)rvb[i]);
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}

// Two 8-bit row masks b0, b1 of one bitmap half (left in place: row r of the half occupies bits 31-8r..24-8r)
// advance together: every round takes the highest set bit p of each mask, multiplies the row's next value
// (shared address va + round * sizeof(T)) with x[column] (shared address xb - p * sizeof(X) - 8 r sizeof(X),
// the row term folded into the load's immediate) and clears the bit.  The whole walk is one PTX block so that
// the predicates stay in predicate registers from round to round: per value FLO, IMAD, 2 LDS, BMSK, LOP3,
// (cvt,) FFMA, and no branch other than the early exit once both masks are empty.  Value loads are not
// predicated (their addresses stay inside the staged tile), x loads and the FMAs are.
#define BMSP_RND(OFF, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB)                                        \
    "bfind.u32 p0, %2;\n\tbfind.u32 p1, %3;\n\t"                                                          \
    "mad.lo.s32 a0, p0, " XMUL ", %6;\n\tmad.lo.s32 a1, p1, " XMUL ", %6;\n\t"                            \
    VLD "0, [%4+" #OFF "];\n\t" VLD "1, [%5+" #OFF "];\n\t"                                               \
    "@q0 " XLD "0, [a0+" XA "];\n\t@q1 " XLD "1, [a1+" XB "];\n\t"                                        \
    "bmsk.clamp.b32 m0, p0, 1;\n\tbmsk.clamp.b32 m1, p1, 1;\n\t" VC0 VC1 XC0 XC1                          \
    "@q0 fma.rn.f32 %0, f0, x0, %0;\n\t@q1 fma.rn.f32 %1, f1, x1, %1;\n\t"                                \
    "not.b32 m0, m0;\n\tnot.b32 m1, m1;\n\tand.b32 %2, %2, m0;\n\tand.b32 %3, %3, m1;\n\t"                \
    "setp.ne.u32 q0, %2, 0;\n\tsetp.ne.u32 q1, %3, 0;\n\tor.pred qa, q0, q1;\n\t@!qa bra DONE;\n\t"
#define BMSP_PAIR(O0, O1, O2, O3, O4, O5, O6, O7, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB)            \
    asm("{\n\t.reg .pred q0, q1, qa;\n\t.reg .u32 p0, p1, m0, m1, a0, a1, h0, h1, g0, g1;\n\t"            \
        ".reg .f32 x0, x1, f0, f1;\n\t.reg .b16 lo, hi;\n\t"                                              \
        "setp.ne.u32 q0, %2, 0;\n\tsetp.ne.u32 q1, %3, 0;\n\tor.pred qa, q0, q1;\n\t@!qa bra DONE;\n\t"   \
        BMSP_RND(O0, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB) BMSP_RND(O1, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB)  \
        BMSP_RND(O2, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB) BMSP_RND(O3, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB)  \
        BMSP_RND(O4, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB) BMSP_RND(O5, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB)  \
        BMSP_RND(O6, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB) BMSP_RND(O7, VLD, VC0, VC1, XLD, XC0, XC1, XMUL, XA, XB)  \
        "DONE:\n\t}"                                                                                      \
        : "+f"(acc0), "+f"(acc1), "+r"(b0), "+r"(b1) : "r"(va0), "r"(va1), "r"(xb))
#define BMSP_CVT16(dst, src) "mov.b32 {lo, hi}, " src ";\n\tcvt.f32.f16 " dst ", lo;\n\t"
// PAIR 0: rows 0,1 of the half (x immediates 0 and -8 columns); PAIR 1: rows 2,3 (-16 and -24 columns)
template <typename T, typename X, int PAIR> struct RowPair;
#define BMSP_ROWPAIR(TT, XX, PP, ...)                                                                                             \
    template <> struct RowPair<TT, XX, PP> {                                                                                      \
        static __device__ __forceinline__ void run(float& acc0, float& acc1, uint32_t b0, uint32_t b1, uint32_t va0, uint32_t va1, \
                                                   uint32_t xb) {                                                                 \
            BMSP_PAIR(__VA_ARGS__);                                                                                               \
        }                                                                                                                         \
    };
#define BMSP_V16 0, 2, 4, 6, 8, 10, 12, 14, "ld.shared.u16 h", BMSP_CVT16("f0", "h0"), BMSP_CVT16("f1", "h1")
#define BMSP_V32 0, 4, 8, 12, 16, 20, 24, 28, "ld.shared.f32 f", "", ""
#define BMSP_X32 "ld.shared.f32 x", "", "", "-4"
#define BMSP_X16 "ld.shared.u16 g", BMSP_CVT16("x0", "g0"), BMSP_CVT16("x1", "g1"), "-2"
BMSP_ROWPAIR(__half, float, 0, BMSP_V16, BMSP_X32, "0", "-32")
BMSP_ROWPAIR(__half, float, 1, BMSP_V16, BMSP_X32, "-64", "-96")
BMSP_ROWPAIR(__half, __half, 0, BMSP_V16, BMSP_X16, "0", "-16")
BMSP_ROWPAIR(__half, __half, 1, BMSP_V16, BMSP_X16, "-32", "-48")
BMSP_ROWPAIR(float, float, 0, BMSP_V32, BMSP_X32, "0", "-32")
BMSP_ROWPAIR(float, float, 1, BMSP_V32, BMSP_X32, "-64", "-96")
BMSP_ROWPAIR(float, __half, 0, BMSP_V32, BMSP_X16, "0", "-16")
BMSP_ROWPAIR(float, __half, 1, BMSP_V32, BMSP_X16, "-32", "-48")
#undef BMSP_RND
#undef BMSP_PAIR
#undef BMSP_CVT16
#undef BMSP_ROWPAIR
#undef BMSP_V16
#undef BMSP_V32
#undef BMSP_X32
#undef BMSP_X16

template <typename V> __device__ __forceinline__ float lds_as_f32(uint32_t a);
template <> __device__ __forceinline__ float lds_as_f32<float>(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
template <> __device__ __forceinline__ float lds_as_f32<__half>(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return __half2float(__ushort_as_half(v));
}

// The blocks of one block row seen by the thread that owns bitmap half h (rows 4h..4h+3).  Everything lives in
// shared memory: bitmaps at a_bm, 16-bit x offsets at a_xo, the row's values from a_v on; xs31 = staged x + 31
// elements.
template <typename X> __device__ __forceinline__ void lds_x4(uint32_t a, float (&o)[4]);
template <> __device__ __forceinline__ void lds_x4<float>(uint32_t a, float (&o)[4]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(a));
}
template <> __device__ __forceinline__ void lds_x4<__half>(uint32_t a, float (&o)[4]) {
    uint32_t lo, hi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&lo)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    o[0] = f0.x; o[1] = f0.y; o[2] = f1.x; o[3] = f1.y;
}

// XV: the staged x lines are dense (pitch 32) and 16-byte aligned, so the four x elements of a bitmap half are one vector load
template <typename T, typename X, bool XV = false>
__device__ __forceinline__ void tile_half_row(uint32_t a_bm, uint32_t a_xo, uint32_t a_v, int nb, const int h, const uint32_t xs31,
                                              float (&acc)[4]) {
    constexpr uint32_t SV = sizeof(T), SX = sizeof(X);
    const uint32_t hm = 0u - (uint32_t)h;
#pragma unroll 1
    for (int i = 0; i < nb; i++) {
        const uint2 w2 = lds_v2(a_bm);                  // .y = rows 0-3, .x = rows 4-7
        const uint32_t xo = lds_u16(a_xo);
        a_bm += 8; a_xo += 2;
        const uint32_t nhi = __popc(w2.y), nlo = __popc(w2.x);
        const uint32_t w = (w2.x & hm) | (w2.y & ~hm);
        const uint32_t va0 = a_v + (nhi & hm) * SV;
        a_v += (nhi + nlo) * SV;
        if (w == (0x80402010u >> (4 * h))) {
            // diagonal block (the +-m neighbours of a stencil, any band at a multiple of 8): row i of the half holds exactly
            // column 4h + i -- four straight multiply-adds, no bit walk (20 instructions instead of ~50)
            const uint32_t xa = xs31 - 31u * SX + (xo + 4u * (uint32_t)h) * SX;
            float x4[4];
            if constexpr (XV) lds_x4<X>(xa, x4);
            else {
#pragma unroll
                for (int i = 0; i < 4; i++) x4[i] = lds_as_f32<X>(xa + i * SX);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) acc[i] = fmaf(lds_as_f32<T>(va0 + i * SV), x4[i], acc[i]);
        } else if (w == (h ? 0x1C0E0703u : 0xC0E07038u)) {
            // tridiagonal block (the diagonal block of a stencil, any 3-wide band): row i of the half holds columns
            // 4h+i-1 .. 4h+i+1, clipped to the block -- 11 values.  Written as 12 slots (3 per row) over the six x elements
            // 4h-1 .. 4h+4 so that the code does not depend on h: slot 0 of the upper half (column -1) and slot 11 of the
            // lower half (column 8) do not exist and are predicated off (their loads stay inside the staged tile).
            const uint32_t vb = va0 + ((uint32_t)h - 1u) * SV;                       // slot s at vb + s * SV
            const uint32_t xa = xs31 - 31u * SX + (xo + 4u * (uint32_t)h - 1u) * SX;   // x[4h-1+j] at xa + j * SX
            float xl[6], v[12];
            if constexpr (XV) {
                float x4[4];
                lds_x4<X>(xa + SX, x4);
                xl[0] = lds_as_f32<X>(xa); xl[1] = x4[0]; xl[2] = x4[1]; xl[3] = x4[2]; xl[4] = x4[3]; xl[5] = lds_as_f32<X>(xa + 5 * SX);
            } else {
#pragma unroll
                for (int j = 0; j < 6; j++) xl[j] = lds_as_f32<X>(xa + j * SX);
            }
#pragma unroll
            for (int s = 0; s < 12; s++) v[s] = lds_as_f32<T>(vb + s * SV);
            if (h) acc[0] = fmaf(v[0], xl[0], acc[0]);
            acc[0] = fmaf(v[1], xl[1], acc[0]); acc[0] = fmaf(v[2], xl[2], acc[0]);
            acc[1] = fmaf(v[3], xl[1], acc[1]); acc[1] = fmaf(v[4], xl[2], acc[1]); acc[1] = fmaf(v[5], xl[3], acc[1]);
            acc[2] = fmaf(v[6], xl[2], acc[2]); acc[2] = fmaf(v[7], xl[3], acc[2]); acc[2] = fmaf(v[8], xl[4], acc[2]);
            acc[3] = fmaf(v[9], xl[3], acc[3]); acc[3] = fmaf(v[10], xl[4], acc[3]);
            if (!h) acc[3] = fmaf(v[11], xl[5], acc[3]);
        } else if (w && !(w & (w - 1u))) {
            // one value in this half (the +-1 neighbours of a stencil: a single cell in the whole block): one product, added to
            // the row it belongs to
            const uint32_t t = 31u - (uint32_t)(31 - __clz((int)w));          // = 8 * row + column
            const float prod = lds_as_f32<T>(va0) * lds_as_f32<X>(xs31 - 31u * SX + (xo + (t & 7u)) * SX);
            const uint32_t rr = t >> 3;
            acc[0] += rr == 0u ? prod : 0.f; acc[1] += rr == 1u ? prod : 0.f;
            acc[2] += rr == 2u ? prod : 0.f; acc[3] += rr == 3u ? prod : 0.f;
        } else if (w) {
            const uint32_t xb = xs31 + xo * SX;
            const uint32_t b0 = w & 0xFF000000u, b1 = w & 0x00FF0000u, b2 = w & 0x0000FF00u, b3 = w & 0x000000FFu;
            const uint32_t va1 = va0 + __popc(b0) * SV, va2 = va1 + __popc(b1) * SV, va3 = va2 + __popc(b2) * SV;
            RowPair<T, X, 0>::run(acc[0], acc[1], b0, b1, va0, va1, xb);
            RowPair<T, X, 1>::run(acc[2], acc[3], b2, b3, va2, va3, xb);
        }
    }
}

// The blocks of one whole block row (all 8 matrix rows) walked by one thread: the per-block bookkeeping is paid
// once per block instead of once per bitmap half.
template <typename T, typename X>
__device__ __forceinline__ void tile_block_row(uint32_t a_bm, uint32_t a_xo, uint32_t a_v, int nb, const uint32_t xs31, float (&acc)[8]) {
    constexpr uint32_t SV = sizeof(T), SX = sizeof(X);
#pragma unroll 1
    for (int i = 0; i < nb; i++) {
        const uint2 w2 = lds_v2(a_bm);                  // .y = rows 0-3, .x = rows 4-7
        const uint32_t xo = lds_u16(a_xo);
        a_bm += 8; a_xo += 2;
        const uint32_t hi = w2.y, lo = w2.x;
        const uint32_t xb = xs31 + xo * SX;
        const uint32_t b0 = hi & 0xFF000000u, b1 = hi & 0x00FF0000u, b2 = hi & 0x0000FF00u, b3 = hi & 0x000000FFu;
        const uint32_t b4 = lo & 0xFF000000u, b5 = lo & 0x00FF0000u, b6 = lo & 0x0000FF00u, b7 = lo & 0x000000FFu;
        const uint32_t va0 = a_v, va1 = va0 + __popc(b0) * SV, va2 = va1 + __popc(b1) * SV, va3 = va2 + __popc(b2) * SV;
        const uint32_t va4 = va3 + __popc(b3) * SV, va5 = va4 + __popc(b4) * SV, va6 = va5 + __popc(b5) * SV, va7 = va6 + __popc(b6) * SV;
        a_v = va7 + __popc(b7) * SV;
        if (hi) {
            RowPair<T, X, 0>::run(acc[0], acc[1], b0, b1, va0, va1, xb);
            RowPair<T, X, 1>::run(acc[2], acc[3], b2, b3, va2, va3, xb);
        }
        if (lo) {
            RowPair<T, X, 0>::run(acc[4], acc[5], b4, b5, va4, va5, xb);
            RowPair<T, X, 1>::run(acc[6], acc[7], b6, b7, va6, va7, xb);
        }
    }
}

template <typename X> __device__ __forceinline__ void sts_x(uint32_t a, X v);
template <> __device__ __forceinline__ void sts_x<float>(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
template <> __device__ __forceinline__ void sts_x<__half>(uint32_t a, __half v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(__half_as_ushort(v)) : "memory");
}
template <typename X> struct alignas(4 * sizeof(X)) XQuad { X e[4]; };

// ---- multi-GPU: halo exchange over peer memory (NVLink) fused into the product, SURVEY.md section 8e -------------
// Every rank keeps x for its extended column range in three peer-mapped buffers (rotating) plus an inbox of epoch
// flags, one slot per peer.  A product reads x[cur] (own slice + halo slices the peers pushed), writes y into the
// own slice of x[nxt] and -- in the same kernel -- stores the rows a peer needs straight into that peer's x[nxt]
// (P2P stores); the last pushing CTA then publishes `signal_epoch` in the peers' inboxes (st.release.sys after a
// system fence).  Only tiles whose x lines leave the rank's own columns wait (spin on the inbox until every peer's
// slot reached `wait_epoch`), and the tile order is rotated by half a grid so that these boundary tiles -- which
// are also the ones that push -- run mid-kernel: the peers' rows of the previous product arrived long before, and
// this product's rows are on their way long before the peers need them, so the exchange costs no time at all.
// Peers signal each other in both directions even when one direction carries no data.  Write-after-read: the
// buffer a peer overwrites during its product e was last read by my product e-2, which finished before my product
// e-1 (whose epoch the peer waited for) started -- hence three buffers.
constexpr int HALO_MAX = 8;
struct HaloDev {
    int32_t n_push, n_peer;
    int32_t lo[HALO_MAX], hi[HALO_MAX];      // local row ranges [lo, hi) pushed to a peer (lo a multiple of 4)
    float* dst[HALO_MAX];                    // peer address of row lo[i] (16-byte aligned)
    uint32_t* peer_flag[HALO_MAX];           // my slot in each peer's inbox
    const uint32_t* my_flag[HALO_MAX];       // the peers' slots in my inbox
    uint32_t* scratch;                       // [0] pushing CTAs that finished, [1] error: a wait timed out
    uint32_t wait_epoch, signal_epoch, n_sig;
    int32_t solo_tile;                       // tile that signals when no tile pushes anything (-1: none)
    int32_t n_iv, iv_lo[HALO_MAX], iv_hi[HALO_MAX];   // the tiles that own pushed rows, as disjoint tile intervals [lo, hi] (host-computed:
                                             // the streaming kernel's producer tests a tile against these instead of the row ranges)
    int32_t own_c0, own_c1;                  // columns [own_c0, own_c1) of x_ext are this rank's own slice: tiles that stay inside never wait
    int32_t rot;                             // tile order rotation: CTA b runs tile (b + rot) mod ntiles, so that the boundary tiles
                                             // (which need the peers' rows and produce the rows the peers need) run mid-kernel
};
struct NoHalo {};

__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// one thread per peer: spin until the peer's slot reached wait_epoch; gives up after 4 s (error flag, no hang)
__device__ __forceinline__ void halo_wait(const HaloDev& h, int i, bool full_fence = true) {
    const uint32_t* f = h.my_flag[i];
    if ((int32_t)(ld_relaxed_sys(f) - h.wait_epoch) < 0) {
        const uint64_t t0 = globaltimer_ns();
        while ((int32_t)(ld_relaxed_sys(f) - h.wait_epoch) < 0) {
            if (*(volatile uint32_t*)(h.scratch + 1)) return;
            if (globaltimer_ns() - t0 > 4000000000ull) { atomicExch(h.scratch + 1, 1u); return; }
            __nanosleep(64);
        }
    }
    // The peer released its flag after fencing its rows.  Readers that go through L1 (the per-tile kernel's x gather, unstaged
    // tiles) need the full fence: it also drops the SM's L1 lines (SASS CCTL.IVALL), and a 32-column line that straddles the own /
    // halo boundary may have been cached before the peer's part of it arrived.  The streaming kernel reads x with bulk copies
    // (L2): an acquire load of the flag pairs with the peer's st.release.sys and does not stall the warp for microseconds.
    if (full_fence) fence_acq_rel_sys();
    else (void)ld_acquire_sys(f);
}
// one thread of every pushing CTA, after the CTA's remote stores were fenced (system scope) and barriered
__device__ __forceinline__ void halo_signal(const HaloDev& h) {
    const uint32_t old = atomicAdd(h.scratch, 1u);
    if (old + 1u == h.n_sig) {
        *(volatile uint32_t*)h.scratch = 0u;             // ready for the next launch
        __threadfence_system();
        for (int p = 0; p < h.n_peer; p++) st_release_sys(h.peer_flag[p], h.signal_epoch);
    }
}
// rows [row, row + NR) of y (NR a multiple of 4 or the ragged tail) to every peer range that contains them
template <int NR>
__device__ __forceinline__ void halo_store(const HaloDev& h, int row, int rows, const float (&acc)[NR]) {
    for (int i = 0; i < h.n_push; i++) {
        const int lo = h.lo[i], hi = h.hi[i];
        if (row >= hi || row + NR <= lo) continue;
        float* d = h.dst[i] + (row - lo);
        if (row >= lo && row + NR <= hi && row + NR <= rows) {
#pragma unroll
            for (int q = 0; q < NR; q += 4) *reinterpret_cast<float4*>(d + q) = make_float4(acc[q], acc[q + 1], acc[q + 2], acc[q + 3]);
        } else {
#pragma unroll
            for (int q = 0; q < NR; q++) if (row + q >= lo && row + q < hi && row + q < rows) d[q] = acc[q];
        }
    }
}

__global__ void halo_wait_kernel(const HaloDev h) {
    if ((int)threadIdx.x < h.n_peer) halo_wait(h, threadIdx.x);
}
// publishes the epoch after a product whose kernels stored the peers' rows without fences of their own: the kernel boundary
// orders those stores before this thread, its system fence and release stores carry them to the peers
__global__ void halo_flag_kernel(const HaloDev h) {
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int p = 0; p < h.n_peer; p++) st_release_sys(h.peer_flag[p], h.signal_epoch);
    }
}
// stand-alone push (first exchange after set_x, and products that run the block-parallel kernel): every thread reads four rows of y
// once and stores them to every peer range that contains them -- 16-byte stores, 512 contiguous bytes per warp and peer, all
// peers' links busy at the same time (for a scattered matrix every peer needs the whole slice: an all-gather written by the owner)
__global__ void __launch_bounds__(256) halo_push_kernel(const float* __restrict__ y, int rows, const HaloDev h) {
    int lo_all = 0x7FFFFFFF, hi_all = 0;
    for (int i = 0; i < h.n_push; i++) { lo_all = min(lo_all, h.lo[i]); hi_all = max(hi_all, h.hi[i]); }
    for (int r = lo_all + 4 * (int)(blockIdx.x * blockDim.x + threadIdx.x); r < hi_all; r += 4 * (int)(gridDim.x * blockDim.x)) {
        float v[4];
        if (r + 4 <= rows) { const float4 t = *reinterpret_cast<const float4*>(y + r); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else { for (int q = 0; q < 4; q++) v[q] = r + q < rows ? y[r + q] : 0.f; }
        for (int i = 0; i < h.n_push; i++) {
            const int lo = h.lo[i], hi = h.hi[i];
            if (r >= hi || r + 4 <= lo) continue;
            float* d = h.dst[i] + (r - lo);
            if (r >= lo && r + 4 <= hi) *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
            else { for (int q = 0; q < 4; q++) if (r + q >= lo && r + q < hi) d[q] = v[q]; }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) halo_signal(h);
}

// One CTA per tile of RTT block rows, TPR threads per block row (1: a thread owns all 8 matrix rows of its block
// row; 2: one thread per 32-bit bitmap half).  Every thread reads the 32-byte tile descriptor; thread 0 arms the
// mbarrier and issues the bulk copies (row pointers, bitmaps, x offsets, values) while all threads gather the
// tile's distinct x lines (coalesced 16-byte loads, stored at a 33-element pitch); latency is hidden by the CTAs
// resident per SM, each in a different phase.
// H = HaloDev: the multi-GPU variant (waits for the peers' halo slices before gathering x, pushes its boundary rows).
template <typename T, typename X, int RTT, int TPR, int MINB, typename H = NoHalo>
__global__ void __launch_bounds__(RTT * TPR, MINB) spmv_tile_kernel(const TileArgs<T> a, const X* __restrict__ x, float* __restrict__ y, const H hd) {
    constexpr bool DIST = !std::is_same<H, NoHalo>::value;
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VA = 16 / sizeof(T);   // values per 16 bytes
    constexpr int NT = RTT * TPR;
    constexpr int LPI = NT / 8;          // x lines per gather step (8 threads per line)
    constexpr int GB = TPR == 1 ? 8 : 4; // gather steps in flight
    static_assert(NT >= 8 && NT % 8 == 0, "8 threads per x line");
    constexpr uint32_t SX = sizeof(X);
    const int tid = threadIdx.x;
    int t = blockIdx.x + a.tile0;
    if constexpr (DIST) { t += hd.rot; if (t >= (int)gridDim.x) t -= (int)gridDim.x; }
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a.so.bar);

    const int4* dp = reinterpret_cast<const int4*>(a.desc + t);
    const int4 d0 = __ldg(dp);
    const int4 d1 = __ldg(dp + 1);       // nl, flags, lmin, lmax
    const int p0 = d0.x, nb = d0.y, nv = d0.w, nl = d1.x;
    const uint32_t v0 = (uint32_t)d0.z, v0a = v0 & ~(uint32_t)(VA - 1);
    const bool staged = (d1.y & 1) && nb <= a.cap_blk && nv <= a.cap_val && nl <= a.cap_lines;
    const int r0 = t * RTT, nrow = min(RTT, a.nbr - r0);

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        const uint32_t nrb = (uint32_t)((nrow + 1 + 1) & ~1) * 8;
        const int p0a = p0 & ~1, p0x = p0 & ~7;
        uint32_t n8 = 0, n2 = 0, nvb = 0;
        if (staged && nb > 0) {
            n8 = (uint32_t)((p0 + nb - p0a + 1) & ~1) * 8;
            n2 = (uint32_t)((p0 + nb - p0x + 7) & ~7) * 2;
            nvb = ((v0 + (uint32_t)nv - v0a + (uint32_t)(VA - 1)) & ~(uint32_t)(VA - 1)) * (uint32_t)sizeof(T);
        }
        mbar_arrive_expect_tx(bar, n8 + n2 + nvb + nrb);
        bulk_g2s(smem + a.so.row, a.rowpair + r0, nrb, bar);
        if (n8) bulk_g2s(smem, a.bmps + p0a, n8, bar);
        if (n2) bulk_g2s(smem + a.so.xo, a.xoff + p0x, n2, bar);
        if (nvb) bulk_g2s(smem + a.so.val, a.values + v0a, nvb, bar);
    }
    bool part = false;               // DIST: this tile owns rows a peer needs (CTA-uniform)
    if constexpr (DIST) {
        const int trow0 = r0 * 8, trow1 = min(a.rows, (r0 + nrow) * 8);
        part = t == hd.solo_tile;
        for (int i = 0; i < hd.n_push; i++) part |= hd.lo[i] < trow1 && hd.hi[i] > trow0;
        // A tile depends on the peers when its x lines leave this rank's own columns (it reads rows they pushed) or when it
        // pushes rows itself: the peers' epoch is the back-pressure that keeps this rank from overwriting a buffer a slower
        // peer still reads, also for one-directional patterns (triangular, directed graphs) where a pushing tile reads nothing
        // remote.  The matrix copies above are already in flight; x must not be touched before the halo is in.
        if (part || (nb > 0 && ((int64_t)d1.z * 32 < hd.own_c0 || ((int64_t)d1.w + 1) * 32 > hd.own_c1))) {
            if (tid < hd.n_peer) halo_wait(hd, tid);
            __syncthreads();
        }
    }
    if (staged) {
        // x lines -> shared: 8 threads per line (4 elements each), GB steps of LPI lines in flight before the first store
        const uint32_t j4 = (uint32_t)(tid & 7) * 4u;
        const int l0 = tid >> 3;
        const uint32_t* lp = a.lines + p0 + l0;
        uint32_t dst = sbase + a.so.xs + ((uint32_t)l0 * XL_STRIDE + j4) * SX;
        const X* xj = x + j4;
        if (!(d1.y & 2)) {
            for (int l = l0; l < nl; l += LPI * GB) {
                XQuad<X> buf[GB];
#pragma unroll
                for (int g = 0; g < GB; g++)
                    if (l + g * LPI < nl) buf[g] = *reinterpret_cast<const XQuad<X>*>(xj + (size_t)__ldg(lp + g * LPI) * 32u);
#pragma unroll
                for (int g = 0; g < GB; g++)
                    if (l + g * LPI < nl) {
#pragma unroll
                        for (int e = 0; e < 4; e++) sts_x<X>(dst + (uint32_t)(g * LPI * XL_STRIDE + e) * SX, buf[g].e[e]);
                    }
                lp += LPI * GB;
                dst += (uint32_t)(LPI * GB * XL_STRIDE) * SX;
            }
        } else {      // the tile's last line reaches past the last column: element-wise guarded loads
            for (int l = l0; l < nl; l += LPI) {
                const uint32_t col = __ldg(lp) * 32u + j4;
#pragma unroll
                for (int e = 0; e < 4; e++) sts_x<X>(dst + e * SX, col + e < (uint32_t)a.cols ? x[col + e] : X(0.f));
                lp += LPI;
                dst += (uint32_t)(LPI * XL_STRIDE) * SX;
            }
        }
    }
    __syncthreads();                 // barrier initialised and armed, x lines visible
    mbar_wait(bar, 0);

    const int lbr = tid / TPR, h = tid % TPR;
    constexpr int NR = 8 / TPR;      // matrix rows per thread
    const int row = (r0 + lbr) * 8 + h * NR;
    const bool active = lbr < nrow && row < a.rows;
    if (!DIST && !active) return;
    float acc[NR];
#pragma unroll
    for (int q = 0; q < NR; q++) acc[q] = 0.f;
    if (active) {
        const uint2 rp = lds_v2(sbase + a.so.row + 8u * lbr);          // (first block, first value) of this block row
        const uint32_t pe = lds_u32(sbase + a.so.row + 8u * lbr + 8u);
        const uint32_t pb = rp.x, kv = rp.y;
        if (staged) {
            const uint32_t rel = pb - (uint32_t)p0;
            const uint32_t a_bm = sbase + ((uint32_t)(p0 & 1) + rel) * 8u, a_xo = sbase + a.so.xo + ((uint32_t)(p0 & 7) + rel) * 2u;
            const uint32_t a_v = sbase + a.so.val + (kv - v0a) * (uint32_t)sizeof(T), xs31 = sbase + a.so.xs + 31u * SX;
            if constexpr (TPR == 2) tile_half_row<T, X>(a_bm, a_xo, a_v, (int)(pe - pb), h, xs31, acc);
            else tile_block_row<T, X>(a_bm, a_xo, a_v, (int)(pe - pb), xs31, acc);
        } else if constexpr (TPR == 2) {
            half_block_row<T, X>(a.bmps, a.bcol, a.values, (int)pb, (int)pe, kv, h, x, acc);
        } else {
            float lo4[4] = {0.f, 0.f, 0.f, 0.f}, hi4[4] = {0.f, 0.f, 0.f, 0.f};
            half_block_row<T, X>(a.bmps, a.bcol, a.values, (int)pb, (int)pe, kv, 0, x, lo4);
            half_block_row<T, X>(a.bmps, a.bcol, a.values, (int)pb, (int)pe, kv, 1, x, hi4);
#pragma unroll
            for (int q = 0; q < 4; q++) { acc[q] = lo4[q]; acc[4 + q] = hi4[q]; }
        }
        float* yr = y + row;
        if (row + NR <= a.rows) {
#pragma unroll
            for (int q = 0; q < NR; q += 4) *reinterpret_cast<float4*>(yr + q) = make_float4(acc[q], acc[q + 1], acc[q + 2], acc[q + 3]);
        } else {
#pragma unroll
            for (int q = 0; q < NR; q++) if (row + q < a.rows) yr[q] = acc[q];
        }
    }
    if constexpr (DIST) {
        if (part) {
            if (active) halo_store<NR>(hd, row, a.rows, acc);
            __threadfence_system();
            __syncthreads();
            if (tid == 0) halo_signal(hd);
        }
    }
}


// ===================================================================================== path 0, streaming kernel
// Persistent CTAs, warp-specialised.  The one-CTA-per-tile kernel above holds a tile's shared memory through three dependent
// global latencies (descriptor -> line list -> x lines) before it computes anything: with shared memory as the limiting
// resource that idle time is what caps the bytes in flight (ncu, round 1: 58 % of the stall samples sit in that prologue).
// Here a CTA is NG groups; a group is one producer warp plus 2*RTT consumer threads (two per block row, one per 32-bit bitmap
// half) and owns `spg` tile-sized stages of shared memory.  Group g of CTA b takes the tiles b + (g + k NG) gridDim.x, k = 0, 1, ..
// -- the SMs sweep the matrix together, so the x lines shared by neighbouring tiles are in L2 at the same time.
//   producer (one elected lane): keeps the 64-byte descriptors of its next tiles arriving by cp.async in a small ring, and as
//     soon as a stage is released issues the tile's bulk copies -- row pointers, bitmaps, x offsets, values, and the tile's x
//     lines as one copy per run of consecutive lines (three runs for a 5-point stencil; line by line, all lanes, for scattered
//     tiles) -- all completing on the stage's `full` mbarrier.  A stage waits for ONE memory latency, not three.
//   consumers: wait for `full`, multiply out of shared memory, store y, arrive on the stage's `empty` mbarrier.
// No CTA-wide barrier after start-up.  One producer warp serves one group: a warp issues its ~200 instructions per tile at one
// every few cycles, so a single producer for a whole CTA of seven groups caps the SM at one tile per 0.7 us (measured).
struct StageLayout { uint32_t bm, xo, val, row, xs, stride; };
inline StageLayout stage_layout(int rt, int cap_blk, int cap_val, int cap_lines, int vsize, int xsize) {
    auto up16 = [](uint32_t v) { return (v + 15u) & ~15u; };
    StageLayout s;
    s.bm = 16;                                                   // 16-byte header: {p0, v0, staged, -}
    s.xo = up16(s.bm + (uint32_t)(cap_blk + 2) * 8);
    s.val = up16(s.xo + (uint32_t)(cap_blk + 16) * 2);
    s.row = up16(s.val + (uint32_t)(cap_val + 16) * vsize);
    s.xs = up16(s.row + (uint32_t)(rt + 2) * 8);
    s.stride = (s.xs + (uint32_t)cap_lines * 32u * xsize + 16u + 127u) & ~127u;   // + 16: the tridiagonal path reads one element past
    return s;
}

template <typename T>
struct StreamArgs {
    TileArgs<T> t;
    StageLayout so;
    int32_t spg, ntiles;         // stages per group; tiles of this launch (tile0 .. tile0 + ntiles)
    int32_t n_static;            // a group's first n_static tiles are its static share (tile slot q0 + k * qstep) ...
    int32_t* sched;              // ... the rest it claims one at a time from sched[0]; sched[1] counts finished groups.  nullptr: all static
};

__device__ __forceinline__ int4 lds_v4(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

constexpr int SCHED_SLOTS = 16;
constexpr int STREAM_DR = 8, STREAM_PD = 4;      // descriptor ring slots / prefetch distance of a producer

// The rare multi-GPU actions of the streaming kernel, out of line: a tile in 2000 waits for the peers or pushes rows, and inlined
// these bodies (spin loops, system fences, a named barrier) cost the multiply loop its schedule.
__device__ __noinline__ void halo_wait_all(const HaloDev* hd, int lane, bool full_fence) {
    if (lane < hd->n_peer) halo_wait(*hd, lane, full_fence);
}
__device__ __noinline__ void halo_push_tile(const HaloDev* hd, int row, int rows, bool active, float a0, float a1, float a2, float a3, int gt, int bar_id, int nthreads) {
    if (active) { const float acc[4] = {a0, a1, a2, a3}; halo_store<4>(*hd, row, rows, acc); }
    // the group's remote stores happen-before its barrier, the barrier before thread 0's system fence (cumulative), the fence
    // before the counter / flag: one fence per pushing tile instead of one per thread -- a system fence stalls its warp for
    // microseconds, and a CTA that stalls finishes its static share of the tiles that much later than everybody else
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    if (gt == 0) { __threadfence_system(); halo_signal(*hd); }
}

template <typename T, typename X, int RTT, int NG, int MINB, typename H = NoHalo>
__global__ void __launch_bounds__(NG * (32 + RTT * 2), MINB) spmv_stream_kernel(const __grid_constant__ StreamArgs<T> sa, const X* __restrict__ x,
                                                                               float* __restrict__ y, const __grid_constant__ H hd) {
    constexpr bool DIST = !std::is_same<H, NoHalo>::value;
    constexpr int GT = RTT * 2;            // consumer threads of a group: two per block row
    constexpr int VA = 16 / sizeof(T);
    constexpr uint32_t SX = sizeof(X);
    constexpr int DR = STREAM_DR, PD = STREAM_PD;
    extern __shared__ __align__(128) unsigned char smem[];
    const TileArgs<T>& a = sa.t;
    const StageLayout so = sa.so;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int spg = sa.spg, nstage = spg * NG;
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nstage * so.stride);      // full[nstage], empty[nstage]
    if (tid == 0) {
        for (int s = 0; s < nstage; s++) { mbar_init(bars + s, 1); mbar_init(bars + nstage + s, GT / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    const bool is_producer = warp < NG;
    const int g = NG == 1 ? 0 : (is_producer ? warp : (warp - NG) / (GT / 32));
    const int grid = (int)gridDim.x;
    // this group's tiles: q_k = blockIdx.x + (g + k NG) * grid, k = 0 .. n_k - 1
    // Tile slots q = 0 .. ntiles - 1 in launch order.  A group's sequence of slots: q0 + k * qstep for its static share, then slots
    // claimed from a counter (sa.sched).  Static shares alone leave the kernel as slow as its slowest CTA: a CTA that stalled -- a
    // system fence after pushing rows to a peer costs microseconds, a tile of a clustered matrix can be several times the average --
    // finishes its share that much later than the rest.  With the last part of the tiles handed out on demand the others absorb it.
    const int q0 = (int)blockIdx.x + g * grid, qstep = NG * grid;
    auto tile_at = [&](int q) {
        if (a.tile_list) return a.tile_list[q];
        if constexpr (DIST) { q += hd.rot; if (q >= sa.ntiles) q -= sa.ntiles; }
        return a.tile0 + q;
    };
    // the group's stages and their barriers
    const uint32_t gbase = sbase + (uint32_t)(g * spg) * so.stride;
    uint64_t* full = bars + g * spg;
    uint64_t* empty = bars + nstage + g * spg;

    if (is_producer) {
        // ------------------------------------------------------------------------------------------ producer
        // The descriptors of the next PD tiles travel global -> shared with cp.async (no register is waited on), one commit group
        // per tile; group k is complete once at most PD newer groups are pending.
        const uint32_t ring = sbase + (uint32_t)nstage * (so.stride + 16u) + (uint32_t)g * (DR * 64u);
        static_assert(NG * DR * 4 <= 64, "tile ids of the ring live in the 64 spare bytes behind it");
        const uint32_t tq = sbase + (uint32_t)nstage * (so.stride + 16u) + (uint32_t)NG * (DR * 64u) + (uint32_t)g * (DR * 4u);
        const int n_static = sa.sched ? sa.n_static : 0x7fffffff;
        const int dyn_base = sa.sched ? sa.n_static * qstep : 0;
        int pending = 0;                  // lane 0: the slot claimed for the next dynamic sequence number, asked for one iteration ahead
        bool ended = false;               // lane 0: a slot past the last tile has been seen (every later one is past it too)
        // descriptor of sequence number k (slot q) -> ring, its tile id -> tq; a slot past the end leaves the end marker (-1)
        auto issue = [&](int k, int q) {
            int t = -1;
            if (q < sa.ntiles) {
                t = tile_at(q);
                const char* src = reinterpret_cast<const char*>(a.desc + t);
                const uint32_t dst = ring + (uint32_t)(k % DR) * 64u;
#pragma unroll
                for (int c = 0; c < 4; c++) cp_async16(dst + 16u * c, src + 16 * c);
            } else ended = true;
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(tq + (uint32_t)(k % DR) * 4u), "r"(t) : "memory");
            cp_async_commit();
        };
        if (lane == 0) {
            for (int k = 0; k < PD; k++) issue(k, k < n_static ? q0 + k * qstep : (ended ? sa.ntiles : dyn_base + atomicAdd(sa.sched, 1)));
            if (PD >= n_static && !ended) pending = dyn_base + atomicAdd(sa.sched, 1);
        }
        int s = 0; uint32_t round = 0;
        for (int k = 0;; k++) {
            int4 d0, d1, d2, d3;
            if (lane == 0) {
                // refill the ring, then make sure tile k's descriptor has landed.  The claim for the sequence number after this one
                // goes out now and is looked at in the next iteration: the atomic's round trip never stalls the warp.
                const int kk = k + PD;
                const int q = kk < n_static ? q0 + kk * qstep : (ended ? sa.ntiles : pending);
                issue(kk, q);
                if (kk + 1 >= n_static && !ended) pending = dyn_base + atomicAdd(sa.sched, 1);
                cp_async_wait<PD>();
            }
            __syncwarp();
            const int t = (int)lds_u32(tq + (uint32_t)(k % DR) * 4u);
            if (t < 0) {
                // no more tiles: tell the consumers through the next stage's header
                if (round) mbar_wait(empty + s, (round - 1u) & 1u);
                if (lane == 0) {
                    *reinterpret_cast<int4*>(smem + (size_t)(g * spg + s) * so.stride) = make_int4(0, 0, 0, -1);
                    mbar_arrive(full + s);
                }
                break;
            }
            const uint32_t dsl = ring + (uint32_t)(k % DR) * 64u;
            d0 = lds_v4(dsl); d1 = lds_v4(dsl + 16); d2 = lds_v4(dsl + 32); d3 = lds_v4(dsl + 48);
            const int p0 = d0.x, nb = d0.y, nv = d0.w, nl = d1.x;
            const uint32_t v0 = (uint32_t)d0.z, v0a = v0 & ~(uint32_t)(VA - 1);
            const bool staged = (d1.y & 1) && nb <= a.cap_blk && nv <= a.cap_val && nl <= a.cap_lines;
            const int r0 = t * RTT, nrow = min(RTT, a.nbr - r0);
            const uint32_t sp32 = gbase + (uint32_t)s * so.stride;
            unsigned char* sp = smem + (size_t)(g * spg + s) * so.stride;
            const uint32_t nrb = (uint32_t)((nrow + 1 + 1) & ~1) * 8;
            const int p0a = p0 & ~1, p0x = p0 & ~7;
            uint32_t n8 = 0, n2 = 0, nvb = 0;
            int nlc = 0;                                                  // x lines that arrive by bulk copy
            if (staged && nb > 0) {
                n8 = (uint32_t)((p0 + nb - p0a + 1) & ~1) * 8;
                n2 = (uint32_t)((p0 + nb - p0x + 7) & ~7) * 2;
                nvb = ((v0 + (uint32_t)nv - v0a + (uint32_t)(VA - 1)) & ~(uint32_t)(VA - 1)) * (uint32_t)sizeof(T);
                nlc = (d1.y & 2) ? nl - 1 : nl;                          // a last line that reaches past the last column is loaded by hand
            }
            bool part = false;                                            // DIST: this tile owns rows a peer needs
            if constexpr (DIST) {
                // A tile whose x lines leave this rank's own columns must not read x before the peers' rows are in, and a tile that
                // pushes rows waits too (back-pressure, see spmv_tile_kernel).  The producer waits on the tile's behalf before it
                // completes `full`; the rotated tile order puts these tiles mid-kernel, where the wait is over before it starts.
                // Whether the tile pushes goes to the consumers through the stage header: evaluated by them, the loop over the push
                // ranges cost every warp ~100 instructions per tile (ncu: 64.5 M warp instructions against 51.9 M).
                part = t == hd.solo_tile;
                for (int i = 0; i < hd.n_iv; i++) part |= t >= hd.iv_lo[i] && t <= hd.iv_hi[i];
                if (part || (nb > 0 && ((int64_t)d1.z * 32 < hd.own_c0 || ((int64_t)d1.w + 1) * 32 > hd.own_c1))) {
                    halo_wait_all(&hd, lane, !staged);
                    __syncwarp();
                    // the x lines are fetched by the async proxy (bulk copies): order them behind the acquire made above
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
            }
            // everything above is ready before the stage is: what follows the wait is the stage's critical path (a group has two
            // stages; a microsecond between a stage's release and its next copies is a microsecond without loads in flight)
            if (round) mbar_wait(empty + s, (round - 1u) & 1u);          // the stage's previous tile has been consumed
            if (staged && nlc < nl) {
                const uint32_t col = (uint32_t)d1.w * 32u + (uint32_t)lane;          // lmax is the tile's last line
                sts_x<X>(sp32 + so.xs + ((uint32_t)(nl - 1) * 32u + (uint32_t)lane) * SX, col < (uint32_t)a.cols ? x[col] : X(0.f));
                __syncwarp();
            }
            const bool by_runs = d2.x > 0;
            if (lane == 0) {
                *reinterpret_cast<int4*>(sp) = make_int4(p0, (int)v0, (staged ? 1 : 0) | (part ? 2 : 0), t);     // the consumers take the tile from here
                mbar_arrive_expect_tx(full + s, n8 + n2 + nvb + nrb + (uint32_t)nlc * 32u * SX);
                bulk_g2s(sp + so.row, a.rowpair + r0, nrb, full + s);
                if (n8) bulk_g2s(sp + so.bm, a.bmps + p0a, n8, full + s);
                if (n2) bulk_g2s(sp + so.xo, a.xoff + p0x, n2, full + s);
                if (nvb) bulk_g2s(sp + so.val, a.values + v0a, nvb, full + s);
                if (nlc > 0 && by_runs) {                                 // up to three runs of consecutive lines
                    unsigned char* xs = sp + so.xs;
                    const uint32_t rl[3] = {(uint32_t)d2.y, (uint32_t)d2.z, (uint32_t)d2.w}, rc[3] = {(uint32_t)d3.x, (uint32_t)d3.y, (uint32_t)d3.z};
                    uint32_t first = 0;
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        if (r < d2.x) {
                            uint32_t cnt = rc[r];
                            if (first + cnt > (uint32_t)nlc) cnt = (uint32_t)nlc - first;     // the hand-loaded last line
                            if (cnt) bulk_g2s(xs + (size_t)first * 32u * SX, x + (size_t)rl[r] * 32u, cnt * 32u * SX, full + s);
                            first += rc[r];
                        }
                    }
                }
            }
            if (nlc > 0 && !by_runs) {                                    // scattered lines: one copy per line, all lanes
                __syncwarp();                                             // after lane 0's expect_tx
                unsigned char* xs = sp + so.xs;
                for (int j = lane; j < nlc; j += 32)
                    bulk_g2s(xs + (size_t)j * 32u * SX, x + (size_t)__ldg(a.lines + p0 + j) * 32u, 32u * SX, full + s);
            }
            if (++s == spg) { s = 0; round++; }
        }
        if (sa.sched && lane == 0) {
            // the last group to finish claiming re-arms the counters for the next launch that takes this slot
            __threadfence();
            if (atomicAdd(sa.sched + 1, 1) == NG * grid - 1) { sa.sched[0] = 0; sa.sched[1] = 0; }
        }
    } else {
        // ------------------------------------------------------------------------------------------ consumers
        const int gt = tid - NG * 32 - g * GT;
        const int lbr = gt >> 1, h = gt & 1;
        int s = 0; uint32_t par = 0;
        for (;;) {
            mbar_wait(full + s, par);
            const uint32_t sb = gbase + (uint32_t)s * so.stride;
            const int4 hdr = lds_v4(sb);                                  // {first block, first value, staged | pushes << 1, tile}
            const int t = hdr.w;
            if (t < 0) break;                                             // the producer ran out of tiles
            const int r0 = t * RTT, nrow = min(RTT, a.nbr - r0);
            const int p0 = hdr.x; const uint32_t v0a = (uint32_t)hdr.y & ~(uint32_t)(VA - 1);
            const bool staged = (hdr.z & 1) != 0;
            const int row = (r0 + lbr) * 8 + h * 4;
            const bool active = lbr < nrow && row < a.rows;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (active) {
                const uint2 rp = lds_v2(sb + so.row + 8u * lbr);          // (first block, first value) of this block row
                const uint32_t pe = lds_u32(sb + so.row + 8u * lbr + 8u);
                const uint32_t pb = rp.x, kv = rp.y;
                if (staged) {
                    const uint32_t rel = pb - (uint32_t)p0;
                    const uint32_t a_bm = sb + so.bm + ((uint32_t)(p0 & 1) + rel) * 8u, a_xo = sb + so.xo + ((uint32_t)(p0 & 7) + rel) * 2u;
                    const uint32_t a_v = sb + so.val + (kv - v0a) * (uint32_t)sizeof(T), xs31 = sb + so.xs + 31u * SX;
                    tile_half_row<T, X, true>(a_bm, a_xo, a_v, (int)(pe - pb), h, xs31, acc);
                } else {
                    half_block_row<T, X>(a.bmps, a.bcol, a.values, (int)pb, (int)pe, kv, h, x, acc);
                }
                float* yr = y + row;
                if (row + 4 <= a.rows) *reinterpret_cast<float4*>(yr) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                else {
#pragma unroll
                    for (int q = 0; q < 4; q++) if (row + q < a.rows) yr[q] = acc[q];
                }
            }
            if constexpr (DIST) {
                if (hdr.z & 2)                                           // the producer found rows of this tile in a push range
                    halo_push_tile(&hd, row, a.rows, active, acc[0], acc[1], acc[2], acc[3], gt, 1 + g, GT);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
            if (++s == spg) { s = 0; par ^= 1u; }
        }
    }
}

// ------------------------------------------------------------------------------------ path 1
// work item: x = block row, y = first block, z = end block, w = 1 when the block row is sliced.
// One warp per item, UNR x 32 blocks per step: every lane issues the metadata loads of its UNR blocks together, the UNR
// warp scans of popc (value offsets) interleave, and the first value / x element of each of the UNR blocks -- with
// about one value per block that is nearly all of them -- are loaded back to back before any of them is used, so a
// step has 4 * UNR independent loads in flight per lane instead of a chain of four dependent ones.
// Multi-GPU: the product itself is unchanged; a wait kernel before it and halo_push_kernel after it do the exchange (bmsp_spmv_halo).
// Two in-kernel variants were measured and dropped.  Every warp storing its finished block row to the peers (32-byte pieces):
// 3.78x at 8 GPUs on R-MAT-22 -- NVLink wants 256-byte writes.  Collecting a CTA's eight block rows in shared memory and storing
// 256 bytes per peer: the CTA-wide barrier keeps seven warps waiting for the one that drew a hub-row slice (2 GPUs: 404 us against
// 368 us).  And a system fence per CTA invalidates the SM's L1 (SASS: MEMBAR.SYS + CCTL.IVALL), which this kernel lives on: with
// fences at the start and the end of every CTA the 2-GPU product took 1161 us against 600 us on one GPU.
template <typename T, typename X>
__global__ void __launch_bounds__(256) spmv_blockpar_kernel(const uint64_t* __restrict__ bmps, const int32_t* __restrict__ bcol,
                                                           const uint64_t* __restrict__ offsets, const T* __restrict__ values,
                                                           const int4* __restrict__ work, int n_work, int rows,
                                                           const X* __restrict__ x, float* __restrict__ y,
                                                           float* __restrict__ partial, const int32_t* __restrict__ brp) {
    constexpr int UNR = 4;
    __shared__ float s_acc[8][8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int item = blockIdx.x * 8 + wid;
    if (item >= n_work) return;
    const bool active = true;
    const int4 w = work[item];
    float (*acc)[32] = s_acc[wid];
#pragma unroll
    for (int r = 0; r < 8; r++) acc[r][lane] = 0.f;
    if (w.w == 2) {
        // bundle: lanes 8q .. 8q+7 take block row w.x + q, one block per lane (at most 8 per row); the eight per-row sums are
        // reduced over the 8 lanes of the group (3 shuffle steps each) and lane 8q + r stores row r: 128 contiguous bytes per warp
        const int sub = lane >> 3, l8 = lane & 7, br = w.x + sub;
        const int b = __ldg(brp + br) + l8;
        if (b < __ldg(brp + br + 1)) {
            uint64_t rem = ld_stream_u64(bmps + b);
            const uint32_t xb = (uint32_t)ld_stream_s32(bcol + b) * 8u;
            uint64_t k = offsets[b];
            while (rem) {
                const int p = __clzll((long long)rem);
                rem &= ~(0x8000000000000000ull >> p);
                acc[p >> 3][lane] += val_to_f32(values[k++]) * ld_x<X>(x, xb + (uint32_t)(p & 7));
            }
        }
        float res = 0.f;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float v = acc[r][lane];
            v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
            if (l8 == r) res = v;
        }
        const int64_t row = (int64_t)br * 8 + l8;
        if (row < rows) y[row] = res;
        return;
    }
    uint64_t vbase = w.y < w.z ? offsets[w.y] : 0;
    for (int b0 = w.y; b0 < w.z; b0 += 32 * UNR) {
        uint64_t bmp[UNR]; uint32_t xb[UNR], cnt[UNR], inc[UNR];
#pragma unroll
        for (int g = 0; g < UNR; g++) {
            const int b = b0 + g * 32 + lane;
            const bool valid = b < w.z;
            bmp[g] = valid ? ld_stream_u64(bmps + b) : 0ull;
            xb[g] = valid ? (uint32_t)ld_stream_s32(bcol + b) * 8u : 0u;
        }
#pragma unroll
        for (int g = 0; g < UNR; g++) { cnt[g] = __popcll(bmp[g]); inc[g] = cnt[g]; }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int g = 0; g < UNR; g++) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc[g], o);
                if (lane >= o) inc[g] += t;
            }
        }
        uint64_t k[UNR];
#pragma unroll
        for (int g = 0; g < UNR; g++) {
            k[g] = vbase + inc[g] - cnt[g];
            vbase += __shfl_sync(0xffffffffu, inc[g], 31);
        }
        // first value of every block: independent loads, issued together
        float v0[UNR], x0[UNR]; int p0[UNR];
#pragma unroll
        for (int g = 0; g < UNR; g++) {
            p0[g] = bmp[g] ? __clzll((long long)bmp[g]) : 0;
            v0[g] = bmp[g] ? val_to_f32(values[k[g]]) : 0.f;
            x0[g] = bmp[g] ? ld_x<X>(x, xb[g] + (uint32_t)(p0[g] & 7)) : 0.f;
        }
#pragma unroll
        for (int g = 0; g < UNR; g++) {
            if (bmp[g]) {
                acc[p0[g] >> 3][lane] += v0[g] * x0[g];
                uint64_t rem = bmp[g] & ~(0x8000000000000000ull >> p0[g]);
                uint64_t kk = k[g] + 1;
                while (rem) {
                    const int p = __clzll((long long)rem);
                    rem &= ~(0x8000000000000000ull >> p);
                    acc[p >> 3][lane] += val_to_f32(values[kk]) * ld_x<X>(x, xb[g] + (uint32_t)(p & 7));
                    kk++;
                }
            }
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
    if (lane < 8 && active) {
        if (w.w) partial[(int64_t)item * 8 + lane] = res;
        else {
            const int64_t row = (int64_t)w.x * 8 + lane;
            if (row < rows) y[row] = res;
        }
    }
}

// sliced block rows (more than one work item): listed once per plan; spmv_fixup_kernel sums their slices
__global__ void split_list_kernel(const int32_t* __restrict__ item_ofs, int nbr, int32_t* __restrict__ list, int32_t* __restrict__ count) {
    const int br = blockIdx.x * blockDim.x + threadIdx.x;
    if (br < nbr && item_ofs[br + 1] - item_ofs[br] > 1) list[atomicAdd(count, 1)] = br;
}

// One warp per sliced block row: lane l sums the slices l, l + 32, ... (two 16-byte loads per slice: its eight row partials), then
// a fixed shuffle tree adds the lanes -- the order of the additions depends on nothing but the slice count, so the result is
// deterministic.  (One thread per matrix row walking all slices in turn took 12 us on R-MAT-22, whose heaviest block row has ~800
// slices: 10 % of a product at 8 GPUs.)
__global__ void __launch_bounds__(256) spmv_fixup_kernel(const int32_t* __restrict__ item_ofs, const float* __restrict__ partial,
                                                         const int32_t* __restrict__ split_list, int n_split_rows, int rows, float* __restrict__ y) {
    const int w = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n_split_rows) return;
    const int br = split_list[w];
    const int i0 = item_ofs[br], i1 = item_ofs[br + 1];
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = i0 + lane; i < i1; i += 32) {
        const float4 a = *reinterpret_cast<const float4*>(partial + (int64_t)i * 8), b = *reinterpret_cast<const float4*>(partial + (int64_t)i * 8 + 4);
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w; s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int r = 0; r < 8; r++) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    }
    float mine = 0.f;
#pragma unroll
    for (int r = 0; r < 8; r++) if (lane == r) mine = s[r];
    const int64_t row = (int64_t)br * 8 + lane;
    if (lane < 8 && row < rows) y[row] = mine;
}

// Four consecutive block rows (an aligned group) with at most 8 blocks each form ONE work item, a "bundle" (w = 2): eight lanes per
// block row instead of a warp.  R-MAT-22 has 4.2 M matrix rows, 52 % of them empty: a warp per block row cost 41 ns per row -- 29 %
// of the product -- mostly its 40 reduction shuffles.
constexpr int BUNDLE_ROWS = 4, BUNDLE_MAXB = 8;
__device__ __forceinline__ bool is_bundle(const int32_t* __restrict__ brp, int nbr, int br) {
    const int g0 = br & ~(BUNDLE_ROWS - 1);
    if (g0 + BUNDLE_ROWS > nbr) return false;
    bool ok = true;
#pragma unroll
    for (int q = 0; q < BUNDLE_ROWS; q++) ok &= brp[g0 + q + 1] - brp[g0 + q] <= BUNDLE_MAXB;
    return ok;
}
__global__ void work_count_kernel(const int32_t* __restrict__ brp, int nbr, uint32_t* __restrict__ cnt, int bundles) {
    int br = blockIdx.x * blockDim.x + threadIdx.x;
    if (br >= nbr) return;
    if (bundles && is_bundle(brp, nbr, br)) { cnt[br] = (br & (BUNDLE_ROWS - 1)) == 0 ? 1u : 0u; return; }
    int nb = brp[br + 1] - brp[br];
    cnt[br] = nb <= SLICE ? 1u : (uint32_t)((nb + SLICE - 1) / SLICE);
}
__global__ void work_fill_kernel(const int32_t* __restrict__ brp, int nbr, const uint32_t* __restrict__ ofs, int4* __restrict__ work, int bundles) {
    int br = blockIdx.x * blockDim.x + threadIdx.x;
    if (br >= nbr) return;
    if (bundles && is_bundle(brp, nbr, br)) {
        if ((br & (BUNDLE_ROWS - 1)) == 0) work[ofs[br]] = make_int4(br, brp[br], brp[br + BUNDLE_ROWS], 2);
        return;
    }
    int b0 = brp[br], b1 = brp[br + 1];
    uint32_t o = ofs[br], n = ofs[br + 1] - o;
    for (uint32_t i = 0; i < n; i++) {
        int s = b0 + (int)i * SLICE;
        work[o + i] = make_int4(br, s, min(s + SLICE, b1), n > 1 ? 1 : 0);
    }
}

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// path 0 plan: tile descriptors, x lines and per-block x offsets; shared-memory capacities from the tile maxima
// when they fit the per-CTA budget, else from the averages (larger tiles are then read from global memory).
static int plan_tiles(bmsp_matrix_s* m, cudaStream_t st) {
    const int vsize = m->dtype == BMSP_F16 ? 2 : 4;
    // tile height: 64 block rows unless an average tile (bitmaps, x offsets, values, ~x lines) would not fit ~20 KB
    int rt = 64;
    const double row_bytes = ((double)m->nblk * (8 + 2 + 16) + (double)m->nnz * vsize) / m->nbr + 8;
    while (rt > 16 && row_bytes * rt > 20.0 * 1024) rt >>= 1;
    if (const char* e = getenv("BMSP_SPMV_RT")) { const int v = atoi(e); if (v == 16 || v == 32 || v == 64 || v == 128) rt = v; }   // experiments
    m->tile_rows = rt;
    // Light rows (64-block-row tiles: stencils, bands) stream through the persistent kernel.  Heavy rows (dense blocks: 16- or
    // 32-row tiles of 17 KB) are bound by the bit walk, not by memory: a stage then feeds only one or two consumer warps and
    // the one-CTA-per-tile kernel, which keeps 13 such tiles computing per SM, is faster (BC4M: 280 us vs 479 us).
    // BMSP_SPMV_KERNEL = 1 / 2 forces the per-tile / the streaming kernel (A/B runs).
    const int force = env_int("BMSP_SPMV_KERNEL", 0);
    m->spmv_kernel = force == 1 ? 1 : (force == 2 ? 0 : (rt == 64 ? 0 : 1));
    m->xl_pitch = m->spmv_kernel == 1 ? XL_STRIDE : 32;
    const int ntiles = (int)ceil_div(m->nbr, rt);
    unsigned long long* stats = nullptr;
    BMSP_TRY(dev_alloc((void**)&m->tile_desc, sizeof(TileDesc) * (size_t)ntiles, st));
    BMSP_TRY(dev_alloc_t(&m->tile_lines, (size_t)m->nblk + 8, st));
    BMSP_TRY(dev_alloc_t(&m->tile_xoff, (size_t)m->nblk + 16, st));
    BMSP_TRY(dev_alloc((void**)&m->tile_rowpair, sizeof(int2) * ((size_t)m->nbr + 1 + 8), st));
    zip_rows_kernel<<<(unsigned)ceil_div(m->nbr + 1, 256), 256, 0, st>>>(m->brp, m->rvb, m->nbr + 1, (int2*)m->tile_rowpair);
    BMSP_KERNEL_CHECK();
    BMSP_TRY(dev_alloc_t(&stats, 8, st));
    BMSP_CUDA(cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), st));
    if (m->spmv_kernel == 0 && !m->spmv_sched) {
        // tile-claim counters of the streaming kernel (zero before the plan's synchronisation below, so that no later launch -- on
        // whatever stream, captured in a graph or not -- has to wait for them)
        BMSP_TRY(dev_alloc_t(&m->spmv_sched, (size_t)SCHED_SLOTS * 2, st));
        BMSP_CUDA(cudaMemsetAsync(m->spmv_sched, 0, sizeof(int32_t) * SCHED_SLOTS * 2, st));
    }
    tile_plan_kernel<<<ntiles, 256, 0, st>>>(m->brp, m->bcol, m->rvb, m->nbr, m->cols, rt, m->xl_pitch, (TileDesc*)m->tile_desc, m->tile_lines, m->tile_xoff, stats);
    BMSP_KERNEL_CHECK();
    unsigned long long h[8];
    BMSP_CUDA(cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    BMSP_CUDA(cudaStreamSynchronize(st));
    dev_free(stats, st);
    // capacities: the tile maxima when they fit the per-CTA budget, else 1.25 x the averages (larger tiles are then
    // read from global memory by the kernel)
    auto up8 = [](long long v) { return (int)((v + 7) & ~7ll); };
    int cb = up8((long long)h[0]), cv = up8((long long)h[1]), cl = (int)h[2];
    const size_t budget = 40 * 1024;
    if (tile_smem(rt, cb, cv, cl, vsize, 4).total > budget) {
        const double ab = (double)m->nblk / ntiles, av = (double)m->nnz / ntiles, al = (double)h[3] / ntiles;
        cb = std::min(cb, up8((long long)(ab * 1.25) + 16)); cv = std::min(cv, up8((long long)(av * 1.25) + 64));
        cl = std::min(cl, (int)(al * 1.25) + 4);
        while (tile_smem(rt, cb, cv, cl, vsize, 4).total > budget && (cb > 64 || cv > 256 || cl > 16)) {
            cb = std::max(64, up8(cb * 3 / 4)); cv = std::max(256, up8(cv * 3 / 4)); cl = std::max(16, cl * 3 / 4);
        }
    }
    m->cap_blk = std::max(cb, 8); m->cap_val = std::max(cv, 8); m->cap_lines = std::max(cl, 1);
    return BMSP_OK;
}

int plan_spmv(bmsp_matrix_s* m, cudaStream_t st) {
    if (m->transposed || m->nbr == 0) { m->spmv_path = -1; return BMSP_OK; }
    const double per_blk = m->nblk ? (double)m->nnz / (double)m->nblk : 0.0;
    m->spmv_path = per_blk >= 2.5 ? 0 : 1;
    if (m->spmv_path == 0) return plan_tiles(m, st);
    uint32_t* cnt = nullptr;
    BMSP_TRY(dev_alloc_t(&cnt, (size_t)m->nbr + 1, st));
    static const int bundles = env_int("BMSP_SPMV_BUNDLES", 1);
    work_count_kernel<<<(unsigned)ceil_div(m->nbr, 256), 256, 0, st>>>(m->brp, m->nbr, cnt, bundles);
    BMSP_KERNEL_CHECK();
    BMSP_TRY(exclusive_scan_u32(cnt, cnt, m->nbr, st));
    uint32_t total = 0;
    BMSP_CUDA(cudaMemcpyAsync(&total, cnt + m->nbr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    BMSP_CUDA(cudaStreamSynchronize(st));
    m->n_work = (int32_t)total;
    BMSP_TRY(dev_alloc((void**)&m->work, sizeof(int4) * (size_t)total, st));
    work_fill_kernel<<<(unsigned)ceil_div(m->nbr, 256), 256, 0, st>>>(m->brp, m->nbr, cnt, (int4*)m->work, bundles);
    BMSP_KERNEL_CHECK();
    m->split_rows = (int32_t*)cnt;           // item offsets per block row, kept for the fix-up
    // sliced block rows (more than one item): listed on the device.  (Items minus block rows no longer says whether there are any:
    // a bundle is one item for four block rows.)  A sliced row has at least two items, so there are at most total / 2 of them.
    {
        const size_t cap = (size_t)total / 2 + 1;
        BMSP_TRY(dev_alloc_t(&m->split_list, cap + 1, st));
        int32_t* cntr = m->split_list + cap;
        BMSP_CUDA(cudaMemsetAsync(cntr, 0, sizeof(int32_t), st));
        split_list_kernel<<<(unsigned)ceil_div(m->nbr, 256), 256, 0, st>>>(m->split_rows, m->nbr, m->split_list, cntr);
        BMSP_KERNEL_CHECK();
        BMSP_CUDA(cudaMemcpyAsync(&m->n_split_rows, cntr, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        BMSP_CUDA(cudaStreamSynchronize(st));
        m->n_split = m->n_split_rows;         // > 0 iff some block row is sliced
        if (m->n_split > 0) BMSP_TRY(dev_alloc_t(&m->split_partial, (size_t)total * 8, st));
    }
    return BMSP_OK;
}

template <typename T, typename X, int RTT, int TPR, int MINB, typename H>
static int launch_tile_kernel(const TileArgs<T>& a, const X* x, float* y, const H& hd, int grid, size_t smem, cudaStream_t st) {
    auto kern = spmv_tile_kernel<T, X, RTT, TPR, MINB, H>;
    static size_t configured = 0;            // per instantiation
    if (configured < smem || configured == 0) {
        BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = std::max<size_t>(smem, 1);
    }
    kern<<<(unsigned)grid, RTT * TPR, smem, st>>>(a, x, y, hd);
    BMSP_KERNEL_CHECK();
    return BMSP_OK;
}

// Streaming kernel launch.  One group per CTA (NG = 1); MAXB bounds the CTAs per SM the registers allow.  Shared memory decides the
// rest: as many CTAs per SM as fit with two stages each (a group needs a second stage to have its next tile arriving while it
// multiplies the current one; P4096: 7 CTAs x 2 stages of 15.4 KB -- the measured optimum).
template <typename T, typename X, int RTT, int NG, int MAXB, typename H>
static int launch_stream_kernel(const TileArgs<T>& a, const X* x, float* y, const H& hd, int ntiles, cudaStream_t st, int32_t* sched = nullptr) {
    auto kern = spmv_stream_kernel<T, X, RTT, NG, MAXB, H>;
    static int sms = 0;
    static size_t smem_max = 0;
    if (!sms) {
        int dev = 0;
        BMSP_CUDA(cudaGetDevice(&dev));
        int v = 0;
        BMSP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        smem_max = (size_t)v;
        BMSP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    StreamArgs<T> sa;
    sa.t = a;
    sa.so = stage_layout(RTT, a.cap_blk, a.cap_val, a.cap_lines, sizeof(T), sizeof(X));
    sa.ntiles = ntiles;
    const size_t fixed = (size_t)NG * STREAM_DR * 64 + 64;                               // descriptor rings
    const size_t per_stage = (size_t)NG * (sa.so.stride + 16);
    static const int want_ctas = env_int("BMSP_SPMV_CTAS", 0), want_spg = env_int("BMSP_SPMV_STAGES", 0);   // experiments
    int ctas = 0, spg = 0;
    size_t per_cta = 0;
    for (int c = want_ctas > 0 ? std::min(want_ctas, MAXB) : MAXB; c >= 1; c--) {
        per_cta = std::min<size_t>(smem_max / c - 1024 - 256, 227 * 1024);               // 1 KB per resident CTA is the system's; 256 B allocation granule
        const int fit = per_cta > fixed ? (int)((per_cta - fixed) / per_stage) : 0;
        if (fit >= 2 || (c == 1 && fit >= 1)) { ctas = c; spg = std::min(fit, 8); break; }
    }
    if (want_spg > 0) spg = std::min(spg, want_spg);
    if (ctas < 1 || spg < 1) { set_error("spmv: a tile stage of %u bytes does not fit shared memory", sa.so.stride); return BMSP_ERR_CUDA; }
    sa.spg = spg;
    const size_t smem = (size_t)spg * per_stage + fixed;
    static size_t configured = 0;            // per instantiation
    if (configured < smem || configured == 0) {
        BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = std::max<size_t>(smem, 1);
    }
    const int grid = std::max(1, std::min((int)ceil_div(ntiles, NG), sms * ctas));
    // the last BMSP_SPMV_DYN percent of every group's share are claimed at run time (0: static shares, 100: every tile is claimed)
    static const int dyn_pct = std::min(100, std::max(0, env_int("BMSP_SPMV_DYN", 25)));
    const int share = ntiles / (grid * NG);
    sa.sched = (sched && dyn_pct > 0 && share >= 2) ? sched : nullptr;
    sa.n_static = sa.sched ? (int)((int64_t)share * (100 - dyn_pct) / 100) : 0;
    kern<<<(unsigned)grid, NG * (32 + RTT * 2), smem, st>>>(sa, x, y, hd);
    BMSP_KERNEL_CHECK();
    return BMSP_OK;
}

// tile0 / ntiles: path 0 only -- launch the tiles [tile0, tile0 + ntiles) (ntiles < 0: all of them).
// hd: NoHalo, or HaloDev for the fused multi-GPU product (path 0, fp32 x, all tiles).
template <typename T, typename X, typename H = NoHalo>
static int launch_spmv(bmsp_matrix_s* A, const X* x, float* y, cudaStream_t st, int tile0 = 0, int ntiles = -1, const H& hd = H(),
                       const int32_t* tile_list = nullptr) {
    if (A->spmv_path == 0) {
        TileArgs<T> a;
        a.tile0 = tile0; a.tile_list = nullptr;
        a.bmps = A->bmps; a.bcol = A->bcol; a.values = (const T*)A->values; a.rowpair = (const int2*)A->tile_rowpair;
        a.desc = (const TileDesc*)A->tile_desc; a.lines = A->tile_lines; a.xoff = A->tile_xoff;
        a.rows = A->rows; a.nbr = A->nbr; a.cols = A->cols; a.cap_blk = A->cap_blk; a.cap_val = A->cap_val; a.cap_lines = A->cap_lines;
        const int rt = A->tile_rows;
        a.so = tile_smem(rt, a.cap_blk, a.cap_val, a.cap_lines, sizeof(T), sizeof(X));
        const size_t smem = a.so.total;
        const int grid = ntiles < 0 ? (int)ceil_div(A->nbr, rt) : ntiles;
        if (grid <= 0) return BMSP_OK;
        a.tile_list = tile_list;
        if (A->spmv_kernel == 0) {
            // streaming kernel: one group (a producer warp + 2 * rt consumer threads) per CTA; CTAs per SM bounded by registers.
            // A launch over all tiles takes one of the matrix's scheduler slots in turn (launches on different streams may overlap:
            // they must not share a counter); launches over a tile range (the host-buffer pipeline's chunks) keep static shares.
            int32_t* sched = nullptr;
            if (ntiles < 0 && tile_list == nullptr) {
                if (A->spmv_sched) sched = A->spmv_sched + 2 * (A->spmv_sched_next++ % SCHED_SLOTS);      // allocated and zeroed by the plan
            }
            if (rt == 128) return launch_stream_kernel<T, X, 128, 1, 3, H>(a, x, y, hd, grid, st, sched);
            if (rt == 32) return launch_stream_kernel<T, X, 32, 1, 8, H>(a, x, y, hd, grid, st, sched);
            if (rt == 16) return launch_stream_kernel<T, X, 16, 1, 8, H>(a, x, y, hd, grid, st, sched);
            return launch_stream_kernel<T, X, 64, 1, 7, H>(a, x, y, hd, grid, st, sched);
        }
        // default: one thread per bitmap half; BMSP_SPMV_VARIANT=1: one thread per block row (64-row tiles only).
        // Measured on P4096: 92.7 us vs 103.5 us (fewer instructions, but too few warps to hide the staging latency).
        if (rt == 128) return launch_tile_kernel<T, X, 128, 2, 6, H>(a, x, y, hd, grid, smem, st);
        if (rt == 32) return launch_tile_kernel<T, X, 32, 2, 24, H>(a, x, y, hd, grid, smem, st);
        if (rt == 16) return launch_tile_kernel<T, X, 16, 2, 32, H>(a, x, y, hd, grid, smem, st);
        return launch_tile_kernel<T, X, 64, 2, 12, H>(a, x, y, hd, grid, smem, st);
    }
    const unsigned grid1 = (unsigned)ceil_div(A->n_work, 8), grid2 = (unsigned)ceil_div((int64_t)A->n_split_rows * 32, 256);
    spmv_blockpar_kernel<T, X><<<grid1, 256, 0, st>>>(A->bmps, A->bcol, A->offsets, (const T*)A->values, (const int4*)A->work, A->n_work,
                                                    A->rows, x, y, A->split_partial, A->brp);
    BMSP_KERNEL_CHECK();
    if (A->n_split > 0) {
        spmv_fixup_kernel<<<grid2, 256, 0, st>>>(A->split_rows, A->split_partial, A->split_list, A->n_split_rows, A->rows, y);
        BMSP_KERNEL_CHECK();
    }
    return BMSP_OK;
}

// ------------------------------------------------------------------------------------ host-buffer pipeline
// y_host = A x_host with both vectors in (pinned) host memory: the block rows are cut into chunks of whole tiles;
// chunk c needs the columns [0, x_need[c]) of x (running maximum of the tile plan's line ranges -- for a banded matrix
// the frontier advances with the rows, for a scattered one the first chunk needs everything and only the y copies
// overlap).  Three streams: s_in feeds x slices (H2D), the caller's stream runs the row-range launches, s_out drains
// y slices (D2H); PCIe is full duplex, so the step costs max(x, y) bytes over the link instead of their sum.
struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_start = nullptr, ev_done = nullptr;
    std::vector<cudaEvent_t> ev_x, ev_k;
    void* x_dev = nullptr;
    float* y_dev = nullptr;
    // the whole pipeline of one (x_host, y_host) pair captured as a CUDA graph: one launch instead of ~7 API calls per chunk
    cudaStream_t s_cap = nullptr;
    cudaGraphExec_t gexec = nullptr;
    const void* g_x = nullptr; const void* g_y = nullptr; int32_t g_xdt = -1;      // what gexec was captured for
    const void* seen_x = nullptr; const void* seen_y = nullptr; int32_t seen_xdt = -1;   // the previous call's buffers
    int nchunks = 0;
    std::vector<int> tile_lo;        // [nchunks+1] first tile of every chunk (path 0); {0, 0} for path 1
    std::vector<int64_t> x_need;     // [nchunks]  columns that must be resident before the chunk runs
};

static int host_pipe_get(bmsp_matrix_s* A, cudaStream_t st, HostPipe** out) {
    if (A->host_pipe) { *out = (HostPipe*)A->host_pipe; return BMSP_OK; }
    HostPipe* hp = new HostPipe();
    A->host_pipe = hp;        // released by spmv_host_release (bmsp_destroy) whatever happens below
    BMSP_CUDA(cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking));
    BMSP_CUDA(cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking));
    BMSP_CUDA(cudaEventCreateWithFlags(&hp->ev_start, cudaEventDisableTiming));
    BMSP_CUDA(cudaEventCreateWithFlags(&hp->ev_done, cudaEventDisableTiming));
    BMSP_TRY(dev_alloc(&hp->x_dev, (size_t)A->cols * 4 + 256, st));
    BMSP_TRY(dev_alloc_t(&hp->y_dev, (size_t)A->rows + 64, st));
    BMSP_CUDA(cudaStreamCreateWithFlags(&hp->s_cap, cudaStreamNonBlocking));
    int want = 8;      // P4096 on PCIe 5 x16: 8 chunks 1.67 ms, 16: 1.68, 32: 1.84 (per-chunk dependencies cost more than the shorter fill saves)
    if (const char* e = getenv("BMSP_HOST_CHUNKS")) want = std::max(1, atoi(e));
    if (A->spmv_path == 0 && (int64_t)A->rows * 4 >= (1 << 20)) {
        const int rt = A->tile_rows, ntiles = (int)ceil_div(A->nbr, rt);
        std::vector<TileDesc> desc((size_t)ntiles);
        BMSP_CUDA(cudaMemcpyAsync(desc.data(), A->tile_desc, sizeof(TileDesc) * (size_t)ntiles, cudaMemcpyDeviceToHost, st));
        BMSP_CUDA(cudaStreamSynchronize(st));
        const int nch = std::min(want, ntiles), per = (int)ceil_div(ntiles, nch);
        int64_t need = 0;
        for (int t0 = 0; t0 < ntiles; t0 += per) {
            const int t1 = std::min(ntiles, t0 + per);
            for (int t = t0; t < t1; t++)
                if (desc[t].nb > 0) need = std::max(need, std::min<int64_t>(A->cols, ((int64_t)desc[t].lmax + 1) * 32));
            hp->tile_lo.push_back(t0);
            hp->x_need.push_back(need);
        }
        hp->tile_lo.push_back(ntiles);
    } else {
        // one chunk: all tiles of a row-tiled matrix ({0, 0} launched nothing and left y_host untouched); the
        // block-parallel path ignores tile_lo
        hp->tile_lo = {0, A->spmv_path == 0 ? (int)ceil_div(A->nbr, A->tile_rows) : 0};
        hp->x_need = {(int64_t)A->cols};
    }
    hp->nchunks = (int)hp->x_need.size();
    hp->ev_x.resize(hp->nchunks); hp->ev_k.resize(hp->nchunks);
    for (auto& e : hp->ev_x) BMSP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : hp->ev_k) BMSP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *out = hp;
    return BMSP_OK;
}

void spmv_host_release(bmsp_matrix_s* m) {
    HostPipe* hp = (HostPipe*)m->host_pipe;
    if (!hp) return;
    if (hp->s_in) { cudaStreamSynchronize(hp->s_in); cudaStreamDestroy(hp->s_in); }
    if (hp->s_out) { cudaStreamSynchronize(hp->s_out); cudaStreamDestroy(hp->s_out); }
    if (hp->s_cap) { cudaStreamSynchronize(hp->s_cap); cudaStreamDestroy(hp->s_cap); }
    if (hp->gexec) cudaGraphExecDestroy(hp->gexec);
    if (hp->ev_start) cudaEventDestroy(hp->ev_start);
    if (hp->ev_done) cudaEventDestroy(hp->ev_done);
    for (auto e : hp->ev_x) if (e) cudaEventDestroy(e);
    for (auto e : hp->ev_k) if (e) cudaEventDestroy(e);
    dev_free(hp->x_dev, m->last_stream); dev_free(hp->y_dev, m->last_stream);     // the product kernels ran on the caller's stream
    delete hp;
    m->host_pipe = nullptr;
}

// y can leave without a copy engine: when y_host is pinned (and therefore mapped into the device's address space) the
// row-range launches store their float4 results straight into it over PCIe (posted writes), so the pipeline is
// H2D slices of x on s_in against kernels that write host memory -- no D2H copies, no events per chunk on the way out.
static float* mapped_host_ptr(float* y_host) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("BMSP_HOST_ZEROCOPY"); enabled = e ? atoi(e) : 1; }
    if (!enabled) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, y_host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer || ((uintptr_t)at.devicePointer & 15)) return nullptr;
    return (float*)at.devicePointer;
}

template <typename T>
static int spmv_host_run(bmsp_matrix_s* A, HostPipe* hp, const void* x_host, int32_t x_dtype, float* y_host, cudaStream_t st) {
    const size_t xs = x_dtype == BMSP_F16 ? 2 : 4;
    float* y_map = A->spmv_path == 0 ? mapped_host_ptr(y_host) : nullptr;     // row-tiled kernel only: it stores 512 contiguous bytes per warp
    float* y_out = y_map ? y_map : hp->y_dev;
    BMSP_CUDA(cudaEventRecord(hp->ev_start, st));         // earlier work on st (and the previous call) is done with x_dev / y_dev
    BMSP_CUDA(cudaStreamWaitEvent(hp->s_in, hp->ev_start, 0));
    BMSP_CUDA(cudaStreamWaitEvent(hp->s_out, hp->ev_start, 0));
    int64_t sent = 0;
    const int64_t rows_per_tile = (int64_t)A->tile_rows * 8;
    cudaEvent_t last_x = nullptr;
    for (int c = 0; c < hp->nchunks; c++) {
        const int64_t need = hp->x_need[c];
        if (need > sent) {
            BMSP_CUDA(cudaMemcpyAsync((char*)hp->x_dev + sent * xs, (const char*)x_host + sent * xs, (size_t)(need - sent) * xs,
                                      cudaMemcpyHostToDevice, hp->s_in));
            BMSP_CUDA(cudaEventRecord(hp->ev_x[c], hp->s_in));
            BMSP_CUDA(cudaStreamWaitEvent(st, hp->ev_x[c], 0));
            sent = need; last_x = hp->ev_x[c];
        }
        int64_t row0 = 0, row1 = A->rows;
        if (A->spmv_path == 0) {
            const int t0 = hp->tile_lo[c], t1 = hp->tile_lo[c + 1];
            row0 = std::min<int64_t>(A->rows, t0 * rows_per_tile); row1 = std::min<int64_t>(A->rows, t1 * rows_per_tile);
            if (x_dtype == BMSP_F32) BMSP_TRY((launch_spmv<T, float>(A, (const float*)hp->x_dev, y_out, st, t0, t1 - t0)));
            else BMSP_TRY((launch_spmv<T, __half>(A, (const __half*)hp->x_dev, y_out, st, t0, t1 - t0)));
        } else {
            if (x_dtype == BMSP_F32) BMSP_TRY((launch_spmv<T, float>(A, (const float*)hp->x_dev, y_out, st)));
            else BMSP_TRY((launch_spmv<T, __half>(A, (const __half*)hp->x_dev, y_out, st)));
        }
        if (row1 > row0 && !y_map) {
            BMSP_CUDA(cudaEventRecord(hp->ev_k[c], st));
            BMSP_CUDA(cudaStreamWaitEvent(hp->s_out, hp->ev_k[c], 0));
            BMSP_CUDA(cudaMemcpyAsync(y_host + row0, hp->y_dev + row0, (size_t)(row1 - row0) * 4, cudaMemcpyDeviceToHost, hp->s_out));
        }
    }
    BMSP_CUDA(cudaEventRecord(hp->ev_done, hp->s_out));
    BMSP_CUDA(cudaStreamWaitEvent(st, hp->ev_done, 0));   // the caller synchronises `st` as for bmsp_spmv
    (void)last_x;                                          // s_in is already joined: st waited on its last event above
    return BMSP_OK;
}

// Second call with the same pinned buffers: capture the pipeline once (origin stream s_cap, forks into s_in / s_out, all
// joined before the capture ends) and replay it with one cudaGraphLaunch on the caller's stream from then on.
template <typename T>
static int spmv_host_graph(bmsp_matrix_s* A, HostPipe* hp, const void* x_host, int32_t x_dtype, float* y_host, cudaStream_t st, bool* used) {
    *used = false;
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("BMSP_HOST_GRAPH"); enabled = e ? atoi(e) : 1; }
    if (!enabled) return BMSP_OK;
    if (!(hp->gexec && hp->g_x == x_host && hp->g_y == y_host && hp->g_xdt == x_dtype)) {
        const bool again = hp->seen_x == x_host && hp->seen_y == y_host && hp->seen_xdt == x_dtype;
        hp->seen_x = x_host; hp->seen_y = y_host; hp->seen_xdt = x_dtype;
        if (!again) return BMSP_OK;                        // first sight of these buffers: run eagerly
        cudaPointerAttributes ax, ay;
        if (cudaPointerGetAttributes(&ax, x_host) != cudaSuccess || cudaPointerGetAttributes(&ay, y_host) != cudaSuccess ||
            ax.type != cudaMemoryTypeHost || ay.type != cudaMemoryTypeHost) { cudaGetLastError(); return BMSP_OK; }   // pageable: eager only
        if (hp->gexec) { cudaGraphExecDestroy(hp->gexec); hp->gexec = nullptr; }
        cudaGraph_t graph = nullptr;
        BMSP_CUDA(cudaStreamBeginCapture(hp->s_cap, cudaStreamCaptureModeThreadLocal));
        const int rc = spmv_host_run<T>(A, hp, x_host, x_dtype, y_host, hp->s_cap);
        const cudaError_t ce = cudaStreamEndCapture(hp->s_cap, &graph);
        if (rc != BMSP_OK || ce != cudaSuccess || !graph) { cudaGetLastError(); if (graph) cudaGraphDestroy(graph); return BMSP_OK; }
        const cudaError_t ie = cudaGraphInstantiate(&hp->gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { cudaGetLastError(); hp->gexec = nullptr; return BMSP_OK; }
        hp->g_x = x_host; hp->g_y = y_host; hp->g_xdt = x_dtype;
    }
    BMSP_CUDA(cudaGraphLaunch(hp->gexec, st));
    *used = true;
    return BMSP_OK;
}

}  // namespace bmsp

using namespace bmsp;

extern "C" int bmsp_spmv_host(bmsp_matrix_t A, const void* x_host, int32_t x_dtype, float* y_host, void* stream) {
    if (!A || !x_host || !y_host || (x_dtype != BMSP_F16 && x_dtype != BMSP_F32)) { set_error("bmsp_spmv_host: invalid argument"); return BMSP_ERR_INVALID; }
    if (A->transposed) { set_error("bmsp_spmv_host: matrix is in transposed-operand form"); return BMSP_ERR_UNSUPPORTED; }
    if (A->rows == 0) return BMSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    touch(A, st);
    if (A->spmv_path < 0) BMSP_TRY(plan_spmv(A, st));
    HostPipe* hp = nullptr;
    BMSP_TRY(host_pipe_get(A, st, &hp));
    bool replayed = false;
    if (A->dtype == BMSP_F16) BMSP_TRY(spmv_host_graph<__half>(A, hp, x_host, x_dtype, y_host, st, &replayed));
    else BMSP_TRY(spmv_host_graph<float>(A, hp, x_host, x_dtype, y_host, st, &replayed));
    if (replayed) return BMSP_OK;
    return A->dtype == BMSP_F16 ? spmv_host_run<__half>(A, hp, x_host, x_dtype, y_host, st) : spmv_host_run<float>(A, hp, x_host, x_dtype, y_host, st);
}

namespace bmsp {

// ---- multi-GPU product with the halo exchange over peer memory ---------------------------------------------------
static int halo_to_dev(const bmsp_halo_desc* d, int rows, HaloDev* h) {
    if (!d || d->n_push < 0 || d->n_push > HALO_MAX || d->n_peer < 0 || d->n_peer > HALO_MAX || (d->n_peer > 0 && !d->scratch)) {
        set_error("halo descriptor: bad counts or missing scratch"); return BMSP_ERR_INVALID;
    }
    memset(h, 0, sizeof(*h));
    h->n_push = 0; h->n_peer = d->n_peer; h->scratch = (uint32_t*)d->scratch; h->solo_tile = -1;
    h->own_c0 = d->own_col_lo; h->own_c1 = d->own_col_hi;
    for (int i = 0; i < d->n_push; i++) {
        if (d->push_lo[i] < 0 || d->push_hi[i] > rows || d->push_lo[i] > d->push_hi[i] || (d->push_lo[i] & 3) || ((uintptr_t)d->push_dst[i] & 15)) {
            set_error("halo descriptor: push range %d [%d,%d) invalid or misaligned", i, d->push_lo[i], d->push_hi[i]); return BMSP_ERR_INVALID;
        }
        if (d->push_lo[i] == d->push_hi[i]) continue;      // an empty range pushes nothing and must not count as a signalling tile
        const int k = h->n_push++;
        h->lo[k] = d->push_lo[i]; h->hi[k] = d->push_hi[i]; h->dst[k] = (float*)d->push_dst[i];
    }
    for (int i = 0; i < d->n_peer; i++) { h->peer_flag[i] = (uint32_t*)d->peer_flag[i]; h->my_flag[i] = (const uint32_t*)d->my_flag[i]; }
    return BMSP_OK;
}

}  // namespace bmsp
using namespace bmsp;

extern "C" int bmsp_halo_push(const float* y_own, int32_t rows, const bmsp_halo_desc* halo, uint32_t signal_epoch, void* stream) {
    if (!y_own) { set_error("bmsp_halo_push: null vector"); return BMSP_ERR_INVALID; }
    HaloDev h;
    BMSP_TRY(halo_to_dev(halo, rows, &h));
    if (h.n_peer == 0) return BMSP_OK;
    int64_t total = 0;
    for (int i = 0; i < h.n_push; i++) total += h.hi[i] - h.lo[i];
    int lo_all = 0x7FFFFFFF, hi_all = 0;
    for (int i = 0; i < h.n_push; i++) { lo_all = std::min(lo_all, h.lo[i]); hi_all = std::max(hi_all, h.hi[i]); }
    (void)total;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)std::max(0, hi_all - lo_all), 1024), 592));
    h.signal_epoch = signal_epoch; h.n_sig = (uint32_t)grid;
    halo_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y_own, rows, h);
    BMSP_KERNEL_CHECK();
    return BMSP_OK;
}

extern "C" int bmsp_halo_status(const bmsp_halo_desc* halo, void* stream, int32_t* timed_out) {
    if (!halo || !timed_out) { set_error("bmsp_halo_status: null argument"); return BMSP_ERR_INVALID; }
    uint32_t v[2] = {0, 0};
    if (halo->scratch) {
        BMSP_CUDA(cudaMemcpyAsync(v, halo->scratch, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        BMSP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    }
    *timed_out = (int32_t)v[1];
    return BMSP_OK;
}

// Multi-GPU product, row-tiled path: which tiles depend on the peers (their x lines leave the rank's own columns) or push rows
// to them.  These "boundary" tiles run first, in a small launch of the halo variant of the streaming kernel (it waits for the
// peers' epoch, multiplies, stores the peers' rows over NVLink and publishes the next epoch -- the peers get their rows a whole
// kernel early); the interior tiles follow in a launch of the plain kernel.  One launch of the halo variant over all tiles cost
// 75.7 us against 67.9 us for the plain kernel on P4096 (one GPU as its own peer) although 32 of 32768 tiles do anything special.
__global__ void halo_classify_kernel(const TileDesc* __restrict__ desc, int ntiles, int rt, int rows, int own_c0, int own_c1, int n_push,
                                     const int* __restrict__ lo, const int* __restrict__ hi, uint32_t* __restrict__ flag) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const TileDesc d = desc[t];
    bool b = d.nb > 0 && ((int64_t)d.lmin * 32 < own_c0 || ((int64_t)d.lmax + 1) * 32 > own_c1);
    const int trow0 = t * rt * 8, trow1 = min(rows, (t + 1) * rt * 8);
    for (int i = 0; i < n_push; i++) b |= lo[i] < trow1 && hi[i] > trow0;
    flag[t] = b ? 1u : 0u;
}
__global__ void halo_lists_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pre, int ntiles, int32_t* __restrict__ list) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint32_t nb = pre[ntiles];
    list[flag[t] ? pre[t] : nb + ((uint32_t)t - pre[t])] = t;
}
static int halo_classify(bmsp_matrix_s* A, const HaloDev& h, cudaStream_t st) {
    int32_t key[18] = {h.own_c0, h.own_c1};
    for (int i = 0; i < h.n_push; i++) { key[2 + 2 * i] = h.lo[i]; key[3 + 2 * i] = h.hi[i]; }
    if (A->halo_n_boundary >= 0 && !memcmp(key, A->halo_key, sizeof(key))) return BMSP_OK;
    const int ntiles = (int)ceil_div(A->nbr, A->tile_rows);
    if (!A->halo_tiles) BMSP_TRY(dev_alloc_t(&A->halo_tiles, (size_t)ntiles + 1, st));
    uint32_t *flag = nullptr, *pre = nullptr; int* rng = nullptr;
    BMSP_TRY(dev_alloc_t(&flag, (size_t)ntiles + 1, st));
    BMSP_TRY(dev_alloc_t(&pre, (size_t)ntiles + 2, st));
    BMSP_TRY(dev_alloc_t(&rng, 2 * HALO_MAX, st));
    BMSP_CUDA(cudaMemcpyAsync(rng, h.lo, sizeof(int) * HALO_MAX, cudaMemcpyHostToDevice, st));
    BMSP_CUDA(cudaMemcpyAsync(rng + HALO_MAX, h.hi, sizeof(int) * HALO_MAX, cudaMemcpyHostToDevice, st));
    halo_classify_kernel<<<(unsigned)ceil_div(ntiles, 256), 256, 0, st>>>((const TileDesc*)A->tile_desc, ntiles, A->tile_rows, A->rows, h.own_c0, h.own_c1,
                                                                        h.n_push, rng, rng + HALO_MAX, flag);
    BMSP_KERNEL_CHECK();
    BMSP_TRY(exclusive_scan_u32(flag, pre, ntiles, st));
    halo_lists_kernel<<<(unsigned)ceil_div(ntiles, 256), 256, 0, st>>>(flag, pre, ntiles, A->halo_tiles);
    BMSP_KERNEL_CHECK();
    uint32_t nb = 0;
    BMSP_CUDA(cudaMemcpyAsync(&nb, pre + ntiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    BMSP_CUDA(cudaMemcpyAsync(&A->halo_first, A->halo_tiles, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    BMSP_CUDA(cudaStreamSynchronize(st));          // also: the pageable h.lo / h.hi copies above are done
    dev_free(flag, st); dev_free(pre, st); dev_free(rng, st);
    A->halo_n_boundary = (int32_t)nb;
    memcpy(A->halo_key, key, sizeof(key));
    return BMSP_OK;
}

extern "C" int bmsp_spmv_halo(bmsp_matrix_t A, const float* x_ext, float* y_own, const bmsp_halo_desc* halo, uint32_t wait_epoch,
                              uint32_t signal_epoch, void* stream) {
    if (!A || !x_ext || !y_own) { set_error("bmsp_spmv_halo: null argument"); return BMSP_ERR_INVALID; }
    if (A->transposed) { set_error("bmsp_spmv_halo: matrix is in transposed-operand form"); return BMSP_ERR_UNSUPPORTED; }
    if (((uintptr_t)x_ext & 15) || ((uintptr_t)y_own & 15)) { set_error("bmsp_spmv_halo: x and y must be 16-byte aligned"); return BMSP_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    touch(A, st);
    HaloDev h;
    BMSP_TRY(halo_to_dev(halo, A->rows, &h));
    h.wait_epoch = wait_epoch; h.signal_epoch = signal_epoch;
    if (A->rows == 0 || h.n_peer == 0) return A->rows == 0 ? BMSP_OK : bmsp_spmv(A, x_ext, BMSP_F32, y_own, stream);
    if (A->spmv_path < 0) BMSP_TRY(plan_spmv(A, st));
    static int fused = -1;
    if (fused < 0) { const char* e = getenv("BMSP_HALO_FUSED"); fused = e ? atoi(e) : 1; }
    if (A->spmv_path == 0 && fused) {
        // tiles that own pushed rows (union of the ranges' tile intervals); they count down to the signal
        const int64_t rpt = (int64_t)A->tile_rows * 8;
        std::vector<std::pair<int64_t, int64_t>> iv;
        for (int i = 0; i < h.n_push; i++) if (h.hi[i] > h.lo[i]) iv.push_back({h.lo[i] / rpt, (h.hi[i] - 1) / rpt});
        std::sort(iv.begin(), iv.end());
        int64_t n = 0, reach = -1;
        h.n_iv = 0;
        for (auto& p : iv) {
            const int64_t a0 = std::max(p.first, reach + 1);
            if (p.second >= a0) { n += p.second - a0 + 1; reach = p.second; h.iv_lo[h.n_iv] = (int32_t)a0; h.iv_hi[h.n_iv] = (int32_t)p.second; h.n_iv++; }
        }
        if (n == 0) { h.solo_tile = 0; n = 1; }
        h.n_sig = (uint32_t)n;
        static int rotate = -1;
        if (rotate < 0) { const char* e = getenv("BMSP_HALO_ROTATE"); rotate = e ? atoi(e) : 1; }
        h.rot = rotate ? (int)(ceil_div(A->nbr, A->tile_rows) / 2) : 0;
        static int split = -1;
        // 0: one launch of the halo variant over all tiles (76 us on P4096, one GPU as its own peer; the plain kernel takes 68 us);
        // 1: boundary launch, then interior launch on the same stream (85.7 us: the small launch is all latency);
        // 2: the boundary launch on a side stream NEXT TO the interior launch
        // (measured 79.5 us: the side-stream launch buys nothing either -- the default stays one launch)
        if (split < 0) { const char* e = getenv("BMSP_HALO_SPLIT"); split = e ? atoi(e) : 0; }
        const int ntiles = (int)ceil_div(A->nbr, A->tile_rows);
        if (split && A->spmv_kernel == 0) {
            BMSP_TRY(halo_classify(A, h, st));
            const int nb = A->halo_n_boundary;
            if (nb * 4 <= ntiles) {
                // boundary tiles (the solo tile, if any, is tile 0: make sure it is among them) with the halo variant, the rest plain
                if (h.solo_tile >= 0 && nb == 0) h.solo_tile = -2;           // nothing pushes and nothing waits: signal from a kernel of its own
                cudaStream_t sb = st;
                if (split == 2 && nb > 0 && ntiles > nb) {
                    if (!A->halo_side) {
                        BMSP_CUDA(cudaStreamCreateWithFlags(&A->halo_side, cudaStreamNonBlocking));
                        BMSP_CUDA(cudaEventCreateWithFlags(&A->halo_fork, cudaEventDisableTiming));
                        BMSP_CUDA(cudaEventCreateWithFlags(&A->halo_join, cudaEventDisableTiming));
                    }
                    sb = A->halo_side;
                    BMSP_CUDA(cudaEventRecord(A->halo_fork, st));              // both launches read what the previous step wrote
                    BMSP_CUDA(cudaStreamWaitEvent(sb, A->halo_fork, 0));
                }
                if (nb > 0) {
                    if (h.solo_tile >= 0) h.solo_tile = A->halo_first;        // the first boundary tile signals on behalf of a rank that pushes nothing
                    BMSP_TRY((A->dtype == BMSP_F16 ? launch_spmv<__half, float, HaloDev>(A, x_ext, y_own, sb, 0, nb, h, A->halo_tiles)
                                                  : launch_spmv<float, float, HaloDev>(A, x_ext, y_own, sb, 0, nb, h, A->halo_tiles)));
                    if (sb != st) BMSP_CUDA(cudaEventRecord(A->halo_join, sb));
                } else {
                    halo_wait_kernel<<<1, 32, 0, st>>>(h);
                    BMSP_KERNEL_CHECK();
                    halo_flag_kernel<<<1, 32, 0, st>>>(h);
                    BMSP_KERNEL_CHECK();
                }
                if (ntiles > nb)
                    BMSP_TRY((A->dtype == BMSP_F16 ? launch_spmv<__half, float, NoHalo>(A, x_ext, y_own, st, 0, ntiles - nb, NoHalo(), A->halo_tiles + nb)
                                                  : launch_spmv<float, float, NoHalo>(A, x_ext, y_own, st, 0, ntiles - nb, NoHalo(), A->halo_tiles + nb)));
                if (sb != st) BMSP_CUDA(cudaStreamWaitEvent(st, A->halo_join, 0));      // the step is complete on the caller's stream
                return BMSP_OK;
            }
        }
        return A->dtype == BMSP_F16 ? launch_spmv<__half, float, HaloDev>(A, x_ext, y_own, st, 0, -1, h)
                                    : launch_spmv<float, float, HaloDev>(A, x_ext, y_own, st, 0, -1, h);
    }
    // block-parallel path (and BMSP_HALO_FUSED=0): wait kernel, product, push kernel
    halo_wait_kernel<<<1, 32, 0, st>>>(h);
    BMSP_KERNEL_CHECK();
    BMSP_TRY(bmsp_spmv(A, x_ext, BMSP_F32, y_own, stream));
    return bmsp_halo_push(y_own, A->rows, halo, signal_epoch, stream);
}

// ---- peer-mapped device memory for the halo buffers (one process per GPU: CUDA IPC handles travel through the caller) ----
extern "C" int bmsp_peer_alloc(int64_t bytes, void** ptr, void* handle64) {
    if (!ptr || !handle64 || bytes <= 0) { set_error("bmsp_peer_alloc: invalid argument"); return BMSP_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    *ptr = nullptr;
    BMSP_CUDA(cudaMalloc(ptr, (size_t)bytes));          // IPC needs a plain allocation, not the stream-ordered pool
    BMSP_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t hd;
    BMSP_CUDA(cudaIpcGetMemHandle(&hd, *ptr));
    memcpy(handle64, &hd, 64);
    return BMSP_OK;
}
extern "C" int bmsp_peer_open(const void* handle64, void** ptr) {
    if (!ptr || !handle64) { set_error("bmsp_peer_open: invalid argument"); return BMSP_ERR_INVALID; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    *ptr = nullptr;
    BMSP_CUDA(cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return BMSP_OK;
}
extern "C" int bmsp_peer_close(void* ptr) { if (ptr) BMSP_CUDA(cudaIpcCloseMemHandle(ptr)); return BMSP_OK; }
extern "C" int bmsp_peer_free(void* ptr) { if (ptr) BMSP_CUDA(cudaFree(ptr)); return BMSP_OK; }

extern "C" int bmsp_spmv(bmsp_matrix_t A, const void* x, int32_t x_dtype, float* y, void* stream) {
    if (!A || !x || !y || (x_dtype != BMSP_F16 && x_dtype != BMSP_F32)) { set_error("bmsp_spmv: invalid argument"); return BMSP_ERR_INVALID; }
    if (A->transposed) { set_error("bmsp_spmv: matrix is in transposed-operand form"); return BMSP_ERR_UNSUPPORTED; }
    if (A->rows == 0) return BMSP_OK;
    if (((uintptr_t)x & 15) || ((uintptr_t)y & 15)) { set_error("bmsp_spmv: x and y must be 16-byte aligned"); return BMSP_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    touch(A, st);
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
