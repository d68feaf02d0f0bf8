// spgemm.cu -- C = A * B on the bmSparse form (A plain fp16, B in transposed-operand form fp16, C plain fp32).
//
// Replaces bmSparse_mult (src/bmSparse_SPGEMM.cu:827-1223).  The reference materialises every candidate
// (A-block, B-block) pair as a 16-byte task (T_3), filters them with 64 AND-tests each (T_4), sorts the
// survivors by C key with an indirect comparator or bb_segsort (T_5), and runs ~20 thrust passes over
// the task list before its numeric kernel.  Here nothing is materialised and nothing is sorted:
//
//   row-wise (Gustavson) over A's block rows, one CTA per block row (dynamic queue), three passes
//     COUNT   candidates are tested four at a time (one 32-bit load of B's inner-dimension masks against the
//             replicated mask of the A block); survivors set a bit in a per-row bit set over C's block
//             columns                                   -> C blocks and survivors per block row  (T_1..T_4)
//     FILL    same walk; survivors are also appended to a global pair list (8 B each -- the only thing that
//             is materialised, after the filter, never sorted).  The bit set is ranked by a popc prefix: the
//             rank of column j IS its sorted position, so C.keys come out ascending without a sort
//             (replaces T_5/T_6 and bb_segsort); the boolean 8x8 product of the operand bitmaps is OR-ed
//             into the block's bitmap; value offsets are scanned per row                              (T_9)
//     NUMERIC bit set rebuilt from C.keys; the pair list is multiplied (scalar lanes for sparse blocks,
//             mma.sync m16n8k8 fp16->fp32 for dense blocks, two B blocks per MMA) and accumulated in
//             shared memory, then stored coalesced                                                    (T_7)
//   block rows whose bit set / block list / value list exceed the shared-memory caps run the same
//   code on global scratch.
#include "common.cuh"
#include <vector>
#include <algorithm>

namespace bmsp {

enum { PASS_COUNT = 0, PASS_FILL = 1, PASS_NUMERIC = 2, PASS_NUMERIC_MMA = 3 };
enum { MODE_SETBITS = 0, MODE_FILL = 1, MODE_NUMERIC = 2, MODE_MMA = 3 };
constexpr int QSLOTS = 192;   // per-warp queue: 64 pairs (uint2) + 32 staged PairMeta (32 B each) for the MMA drain

struct GemmArgs {
    const int32_t* a_brp; const int32_t* a_bcol; const uint64_t* a_bmps; const uint8_t* a_kmask; const uint64_t* a_off; const __half* a_val;
    const int32_t* b_brp; const int32_t* b_bcol; const uint64_t* b_bmps; const uint8_t* b_kmask; const uint64_t* b_off; const __half* b_val;
    const uint4* b_pm;         // packed per-B-block records, two uint4 each: {bitmap lo, bitmap hi, block column, value offset}, {first 8 values}
    // fine index of B (sparse-block operands; null: candidate scan): bucket (k, t) = the blocks of B's block row k whose kmask has
    // bit t, i.e. that have something in row 8k + (inner index of bit t).  An A block with kmask `am` meets exactly the blocks of
    // the buckets (bcol(a), t), t in am -- read as contiguous runs instead of found by testing every block of the B block row.
    const uint32_t* f_ptr; const uint32_t* f_bcol; const uint8_t* f_kmask; const uint4* f_rec;
    const int2* rowinfo;       // per A block row: x = first C block column of the bit set (multiple of 32), y = words
    int32_t row_begin, row_end;
    int32_t G;                 // lanes cooperating on one A block (power of two <= 32)
    int32_t cap_words, cap_c, cap_nnz;
    uint32_t* g_bitset; uint32_t* g_wrank; int32_t max_words;   // per-CTA global scratch for over-cap rows
    int32_t* work_counter;
    const int32_t* row_list;   // when set: the rows to process, in this order (heavy-row / light-row launches); else row_begin + i
    int32_t n_list;
    int32_t batch;             // work items a CTA takes from the queue per atomic (tiny rows: the queue head would serialise the launch)
    int32_t sort_max;          // FILL: leave a work item's segment of the pair list sorted by pair class (see pair_class) when it has at
                               // most this many pairs (0: never; SORT_K x the smallest CTA of the pass: a thread keeps its pairs in registers)
    int32_t split8;            // NUMERIC: 8 lanes per surviving pair, one per row of the A block (blocks with >= 4 values: a lane
                               // that multiplies a whole pair alone walks up to 64 x 8 dependent loads)
    int32_t group;             // consecutive block rows per work item (1..32): tiny rows are processed a group at a time -- one bit set, one
                               // pair-list segment, one set of accumulators for the group -- so that a CTA has hundreds of pairs in flight
    uint32_t* row_count;       // COUNT out: C blocks per row (indexed row - row_begin)
    uint32_t* row_surv;        // COUNT out: surviving pairs per row; FILL/NUMERIC in: exclusive scan of it
    uint2* surv_list;          // FILL out / NUMERIC in: (A block, B block) of every surviving pair, row-segmented
    int32_t* maxes;            // [0] max words, [1] max C blocks per row, [2] max C values per row
    const int32_t* c_brp;      // FILL/NUMERIC in: C block-row pointers (indexed row - row_begin)
    uint64_t* c_keys; uint64_t* c_bmps; int32_t* c_bcol; uint8_t* c_kmask;
    uint64_t* c_off;           // FILL out: value offsets relative to the row; NUMERIC: rebased to absolute
    uint64_t* row_nnz;         // FILL out: values per row; NUMERIC in: exclusive scan of it (row value base)
    float* c_val;
    unsigned long long* stats; // [0] candidate pairs, [1] surviving pairs (COUNT pass)
};

// boolean 8x8 product of an A bitmap (row-major) and a B bitmap in transposed-operand form:
// bit (i,j) = (row i of A) & (byte j of Bt) != 0   -- bmp_calculator, SPGEMM.cu:787-810
// Two evaluations: blocks with few cells (uniform-random and R-MAT inputs average ~1.1 per block) walk A's set bits --
// cell (i,k) contributes column k of B, i.e. bit k of every byte of bt, gathered into one byte by a multiply -- about 8
// instructions per cell; denser blocks take the row-wise SWAR form (8 rows x ~14 instructions whatever the fill).
__device__ __forceinline__ uint64_t pair_bitmap_dense(uint64_t a, uint64_t bt);
__device__ __forceinline__ uint64_t pair_bitmap(uint64_t a, uint64_t bt) {
    if (__popcll(a) > 8) return pair_bitmap_dense(a, bt);
    uint64_t res = 0;
    while (a) {
        const int p = __clzll((long long)a);
        a &= ~(0x8000000000000000ull >> p);
        const int i = p >> 3, k = p & 7;
        const uint64_t col = (bt >> (7 - k)) & 0x0101010101010101ull;           // B(k, j) in bit 0 of byte j (from the MSB)
        res |= ((col * 0x0102040810204080ull) >> 56) << (56 - 8 * i);           // gathered: j = 0 lands in the byte's MSB
    }
    return res;
}
__device__ __forceinline__ uint64_t pair_bitmap_dense(uint64_t a, uint64_t bt) {
    const uint32_t bh = (uint32_t)(bt >> 32), bl = (uint32_t)bt;
    uint64_t res = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t ai = (uint32_t)(a >> (56 - 8 * i)) & 0xFFu;
        const uint32_t rep = ai * 0x01010101u;
        const uint32_t th = rep & bh, tl = rep & bl;
        // 1 in the low bit of every non-zero byte
        const uint32_t nh = ((((th & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | th) & 0x80808080u) >> 7;
        const uint32_t nl = ((((tl & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | tl) & 0x80808080u) >> 7;
        // gather bits 24,16,8,0 into a nibble (bit 24 -> 3 ... bit 0 -> 0): partial products never collide
        const uint32_t byte = ((((nh * 0x01020408u) >> 24) & 0xFu) << 4) | (((nl * 0x01020408u) >> 24) & 0xFu);
        res |= (uint64_t)byte << (56 - 8 * i);
    }
    return res;
}

__device__ __forceinline__ int rank64(uint64_t bmp, int p) { return p == 0 ? 0 : __popcll(bmp >> (64 - p)); }

__global__ void pair_bitmap_test_kernel(const uint64_t* a, const uint64_t* bt, uint64_t* out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pair_bitmap(a[i], bt[i]);
}

// One 32-byte record (= one DRAM sector) per B block: {bitmap lo, bitmap hi, block column, value offset} + its first eight values.
// A surviving pair of a sparse-block product (uniform random, R-MAT: ~1 value per block) then costs ONE random sector for
// everything it needs from B -- column, bitmap and values -- instead of three (b_bcol, the packed metadata, b_val).
__global__ void pack_meta_kernel(const uint64_t* __restrict__ bmps, const int32_t* __restrict__ bcol, const uint64_t* __restrict__ off,
                                 const __half* __restrict__ val, int64_t nnz, uint4* __restrict__ pm, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t b = bmps[i], o = off[i];
    pm[2 * i] = make_uint4((uint32_t)b, (uint32_t)(b >> 32), (uint32_t)bcol[i], (uint32_t)o);
    unsigned short h[8];
    const int cnt = __popcll(b);
#pragma unroll
    for (int k = 0; k < 8; k++) h[k] = (k < cnt && (int64_t)o + k < nnz) ? __half_as_ushort(val[o + k]) : (unsigned short)0;
    pm[2 * i + 1] = make_uint4(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16), h[4] | ((uint32_t)h[5] << 16), h[6] | ((uint32_t)h[7] << 16));
}

// P0: warp per A block row (grid-stride): candidate pairs, and the span [jmin, jmax] of C block columns.  The statistics are
// reduced per warp and per CTA before they touch global memory: with one global atomic -- or even one volatile load -- per row,
// the 2.1 M rows of P4096 serialise on three L2 addresses (3.7 ms for a kernel that otherwise takes 0.3 ms).
__global__ void __launch_bounds__(256) rowinfo_kernel(const int32_t* __restrict__ a_brp, const int32_t* __restrict__ a_bcol,
                                                      const int32_t* __restrict__ b_brp, const int32_t* __restrict__ b_bcol,
                                                      int row_begin, int row_end, int2* __restrict__ rowinfo,
                                                      unsigned long long* __restrict__ cand, int* __restrict__ maxes,
                                                      unsigned long long* __restrict__ sum_words, unsigned long long* __restrict__ max_cand,
                                                      unsigned long long* __restrict__ sum_cand) {
    __shared__ unsigned long long s_words, s_maxc, s_cand;
    __shared__ int s_maxw;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_words = 0; s_maxc = 0; s_maxw = 0; s_cand = 0; }
    __syncthreads();
    unsigned long long w_sum = 0, w_maxc = 0, w_cand = 0;
    int w_maxw = 0;
    for (int row = row_begin + blockIdx.x * 8 + (threadIdx.x >> 5); row < row_end; row += gridDim.x * 8) {
        int jmin = 0x7FFFFFFF, jmax = -1;
        unsigned long long c = 0;
        for (int a = a_brp[row] + lane; a < a_brp[row + 1]; a += 32) {
            const int k = a_bcol[a];
            const int b0 = b_brp[k], b1 = b_brp[k + 1];
            if (b1 > b0) {
                c += (unsigned long long)(b1 - b0);
                jmin = min(jmin, b_bcol[b0]);
                jmax = max(jmax, b_bcol[b1 - 1]);
            }
        }
        for (int o = 16; o; o >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, o);
            jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, o));
            jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, o));
        }
        if (lane == 0) {
            int jbase = 0, words = 0;
            if (jmax >= 0) { jbase = jmin & ~31; words = ((jmax - jbase) >> 5) + 1; }
            rowinfo[row - row_begin] = make_int2(jbase, words);
            cand[row - row_begin] = c;
            w_maxw = max(w_maxw, words); w_sum += (unsigned long long)words; w_maxc = max(w_maxc, c); w_cand += c;
        }
    }
    if (lane == 0) { atomicMax(&s_maxw, w_maxw); atomicAdd(&s_words, w_sum); atomicMax(&s_maxc, w_maxc); atomicAdd(&s_cand, w_cand); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_maxw) atomicMax(maxes, s_maxw);
        if (s_words) atomicAdd(sum_words, s_words);
        if (s_maxc) atomicMax(max_cand, s_maxc);
        if (s_cand) atomicAdd(sum_cand, s_cand);
    }
}

// CTA-wide exclusive scan over popc(bitset words) -> wrank (may be null); returns the total to every thread.
__device__ __forceinline__ uint32_t rank_words(const uint32_t* bitset, uint32_t* wrank, int nwords, uint32_t* s_tmp) {
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (nwords + T - 1) / T;
    const int w0 = min(tid * per, nwords), w1 = min(w0 + per, nwords);
    uint32_t s = 0;
    for (int w = w0; w < w1; w++) s += __popc(bitset[w]);
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_tmp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t v = lane < (T >> 5) ? s_tmp[lane] : 0, vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        s_tmp[lane] = vi - v;
        if (lane == 31) s_tmp[32] = vi;
    }
    __syncthreads();
    uint32_t run = s_tmp[wid] + inc - s;
    if (wrank) for (int w = w0; w < w1; w++) { wrank[w] = run; run += __popc(bitset[w]); }
    const uint32_t total = s_tmp[32];
    __syncthreads();
    return total;
}

// Per-row state shared by the enumeration modes.
struct RowCtx {
    int row, nr;         // first block row of the work item and the number of rows in it
    int a0, a1, jbase, c0;
    const int* abr;      // [nr+1] a_brp of the rows                        (shared)
    const int* jb;       // [nr]   first C block column of each row's bit set
    const int* wo;       // [nr+1] word offset of each row's bit set inside the work item's
    uint32_t* rsurv;     // [nr]   COUNT: surviving pairs per row
    uint32_t* bitset; uint32_t* wrank;
    uint64_t* cbmp;      // FILL: staging (shared or global C.bmps+c0); NUMERIC: shared copy or null
    uint32_t* coff;      // NUMERIC: value offsets relative to the row, or null (global mode)
    float* acc;          // NUMERIC: shared accumulators or null (global atomics)
    float* dense;        // NUMERIC_MMA: [ccount][64] fp32, lane-private slots (cell (r,c) at (c*4 + r/2) + 32*(r&1))
};

// index of the row of a work item that owns v, given the rows' ascending start offsets (A blocks, bit-set words or C blocks)
// (nr <= 32.  Five fixed halving steps: the linear scan this replaces was 15 % of the NUMERIC pass's instructions on P4096 -- every
// lane of every surviving pair walks the 16 rows of its group, and the lanes of a warp stop at different rows.)
__device__ __forceinline__ int local_row(const int* bounds, int nr, int v) {
    int rl = 0;
    if (nr > 1) {
#pragma unroll
        for (int s = 16; s; s >>= 1) {
            const int t = rl + s;
            if (t < nr && v >= bounds[t]) rl = t;
        }
    }
    return rl;
}

// the global loads a surviving pair starts with, kept apart from the work on them so that the scalar NUMERIC pass can have two
// pairs' loads in flight per thread before it touches the first result (A/B on one box: U1M numeric 10.65 -> 9.55 ms, R-MAT-18
// 146 -> 109 ms; the same batching in FILL and in the 8-lanes-per-pair walk cost more in registers than it hid in latency)
struct PairIn { uint4 pm; uint4 bv8; uint64_t abmp; uint32_t aoff; };
template <int MODE>
__device__ __forceinline__ PairIn load_pair(const GemmArgs& g, int a, int b) {
    PairIn in;
    in.pm = __ldg(g.b_pm + 2 * (int64_t)b);
    in.bv8 = MODE == MODE_FILL ? make_uint4(0, 0, 0, 0) : __ldg(g.b_pm + 2 * (int64_t)b + 1);     // same sector
    in.abmp = g.a_bmps[a];
    in.aoff = MODE == MODE_FILL ? 0u : (uint32_t)g.a_off[a];
    return in;
}
// value k (< 8) of a B block out of its record's inline copy
__device__ __forceinline__ float inline_val(const uint4& v, int k) {
    const uint32_t w = (k & 4) ? ((k & 2) ? v.w : v.z) : ((k & 2) ? v.y : v.x);
    return __half2float(__ushort_as_half((unsigned short)((k & 1) ? (w >> 16) : (w & 0xFFFFu))));
}

template <int MODE>
__device__ __forceinline__ void apply_pair(const GemmArgs& g, const RowCtx& r, int a, const PairIn& in, int arow = -1) {
    const uint4 pm = in.pm;
    const uint64_t abmp = in.abmp, bbmp = ((uint64_t)pm.y << 32) | pm.x;
    const int rl = local_row(r.abr, r.nr, a);
    const int j = (int)pm.z - r.jb[rl];
    const int wi = r.wo[rl] + (j >> 5);
    const uint32_t word = r.bitset[wi];
    const int c = (int)r.wrank[wi] + __popc(word & ((1u << (j & 31)) - 1u));
    if (MODE == MODE_FILL) {
        const uint64_t pb = pair_bitmap(abmp, bbmp);
        unsigned int* w = reinterpret_cast<unsigned int*>(&r.cbmp[c]);      // native 32-bit ATOMS.OR, not a 64-bit CAS loop
        if ((uint32_t)pb) atomicOr(w, (unsigned int)pb);
        if ((uint32_t)(pb >> 32)) atomicOr(w + 1, (unsigned int)(pb >> 32));
    } else {
        const __half* av = g.a_val + in.aoff;
        const __half* bv = g.b_val + pm.w;
        const bool b_inline = __popcll(bbmp) <= 8;               // the record carries the block's values
        uint64_t cb; float* dst;
        if (r.acc) { cb = r.cbmp[c]; dst = r.acc + r.coff[c]; }
        else { cb = g.c_bmps[r.c0 + c]; dst = g.c_val + g.c_off[r.c0 + c]; }
        uint64_t rem = abmp;
        int ka = 0;
        if (arow >= 0) {                                    // this lane's row of the A block only
            rem = abmp & (0xFF00000000000000ull >> (8 * arow));
            ka = arow ? __popcll(abmp >> (64 - 8 * arow)) : 0;
            if (!rem) return;
        }
        while (rem) {
            const int p = __clzll((long long)rem);
            rem &= ~(0x8000000000000000ull >> p);
            const int rr = p >> 3, k = p & 7;
            const float aval = __half2float(av[ka++]);
            uint64_t hits = bbmp & (0x8080808080808080ull >> k);    // B(k, c) for c = 0..7 (Bt cell = c*8+k)
            while (hits) {
                const int q = __clzll((long long)hits);
                hits &= ~(0x8000000000000000ull >> q);
                const int kb = rank64(bbmp, q);
                const float bval = b_inline ? inline_val(in.bv8, kb) : __half2float(bv[kb]);
                atomicAdd(dst + rank64(cb, rr * 8 + (q >> 3)), aval * bval);
            }
        }
    }
}

// ---- pair classes (scalar NUMERIC pass with several lanes per pair) ---------------------------------------------------------------
// The lanes of a warp walk their pairs in lockstep, so a warp takes as long as its slowest lane: on a stencil product the four pairs
// of a warp mix tridiagonal x tridiagonal blocks (9 products per lane) with diagonal x diagonal ones (1) and single-value A blocks
// (7 of 8 lanes idle) -- ncu showed ~4 active threads per instruction in the product loops.  FILL, which touches every pair's two
// bitmaps anyway, therefore leaves a work item's segment of the pair list sorted by class: classes 0..8 = A blocks with more than two
// values, ordered by (values per row of A) x (values per row of B) descending; classes 9..11 = A blocks with one or two values (top bit
// of the entry's A index set): NUMERIC gives those ONE lane per pair instead of eight.
constexpr int PAIR_CLASSES = 12, PAIR_LIGHT = 9, SORT_K = 4;
constexpr uint32_t PAIR_LIGHT_FLAG = 0x80000000u;
__device__ __forceinline__ int pair_class(uint64_t abmp, uint64_t bbmp) {
    const int pa = __popcll(abmp), pb = __popcll(bbmp);
    const int cb = min(3, (pb + 7) >> 3);                       // 1..3
    if (pa <= 2) return PAIR_LIGHT + (3 - cb);
    const int ca = min(3, (pa + 7) >> 3);
    return 8 - ((ca - 1) * 3 + (cb - 1));
}

template <int MODE>
__device__ __forceinline__ void process_pair(const GemmArgs& g, const RowCtx& r, int a, int b, int arow = -1) {
    apply_pair<MODE>(g, r, a, load_pair<MODE>(g, a, b), arow);
}

// ---- tensor-core path (dense blocks) -----------------------------------------------------------------
// Two halves of an 8x8 fp16 operand row for mma.m16n8k8: the cells p0 = g*8 + 2t and p0+1 of a block, packed
// as .f16x2 (absent cells are 0).  Works for the A block (cell = row*8 + k) and for the transposed-operand B
// block (cell = col*8 + k) alike -- which is exactly why the reference stores B transposed (SPGEMM.cu:309-312).
__device__ __forceinline__ uint32_t frag_pair(uint64_t bmp, const __half* __restrict__ vals, int p0) {
    const uint32_t two = (uint32_t)(bmp >> (62 - p0)) & 3u;        // bit1 = cell p0, bit0 = cell p0+1
    if (!two) return 0u;
    const int r = rank64(bmp, p0);
    const unsigned short lo = (two & 2u) ? __half_as_ushort(vals[r]) : (unsigned short)0;
    const unsigned short hi = (two & 1u) ? __half_as_ushort(vals[r + (int)(two >> 1)]) : (unsigned short)0;
    return (uint32_t)lo | ((uint32_t)hi << 16);
}

// Drain n (<= 32) queued (A block, B block) pairs with the whole warp.  Stage 1: lane l fetches the metadata of
// pair l (bitmaps, value offsets, dense C slot) -- one round of global latency for the whole batch -- into a per-warp
// shared staging area.  Stage 2: walk the batch; consecutive pairs that share the A block are multiplied two at a
// time by one mma.sync.m16n8k8 (A-operand = the two B^t blocks stacked, B-operand = the A block, D = the two 8x8
// fp32 products transposed); the fragments of the next step are loaded before the current MMA is issued.  Each
// lane owns two fixed slots of every dense C block, so the accumulation needs neither atomics nor warp syncs.
struct PairMeta { uint64_t abmp, bbmp; uint32_t aoff, boff; int32_t cidx; uint32_t a; };   // 32 bytes

__device__ __forceinline__ void drain_mma(const GemmArgs& g, const RowCtx& r, const uint2* q, int n, PairMeta* sm) {
    const int lane = threadIdx.x & 31;
    const int p0 = (lane >> 2) * 8 + (lane & 3) * 2;
    if (lane < n) {
        const uint2 e = q[lane];
        PairMeta m;
        m.a = e.x;
        m.abmp = g.a_bmps[e.x]; m.aoff = (uint32_t)g.a_off[e.x];
        m.bbmp = g.b_bmps[e.y]; m.boff = (uint32_t)g.b_off[e.y];
        const int j = g.b_bcol[e.y] - r.jbase;
        m.cidx = (int)r.wrank[j >> 5] + __popc(r.bitset[j >> 5] & ((1u << (j & 31)) - 1u));
        sm[lane] = m;
    }
    __syncwarp();
    uint32_t fb = 0, fa0 = 0, fa1 = 0, cur_a = 0xFFFFFFFFu;
    int c0 = 0, c1 = 0;
    bool two = false;
    auto load = [&](int i, uint32_t& fb_, uint32_t& fa0_, uint32_t& fa1_, int& c0_, int& c1_, bool& two_, uint32_t& a_) {
        const PairMeta m0 = sm[i];
        two_ = (i + 1 < n) && (sm[i + 1].a == m0.a);
        if (m0.a != a_) { fb_ = frag_pair(m0.abmp, g.a_val + m0.aoff, p0); a_ = m0.a; }
        fa0_ = frag_pair(m0.bbmp, g.b_val + m0.boff, p0);
        c0_ = m0.cidx;
        if (two_) { const PairMeta m1 = sm[i + 1]; fa1_ = frag_pair(m1.bbmp, g.b_val + m1.boff, p0); c1_ = m1.cidx; }
        else fa1_ = 0u;
    };
    int i = 0;
    if (n > 0) load(0, fb, fa0, fa1, c0, c1, two, cur_a);
    while (i < n) {
        const int nxt = i + (two ? 2 : 1);
        uint32_t nfb = fb, nfa0 = 0, nfa1 = 0, na = cur_a; int nc0 = 0, nc1 = 0; bool ntwo = false;
        if (nxt < n) load(nxt, nfb, nfa0, nfa1, nc0, nc1, ntwo, na);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(fa0), "r"(fa1), "r"(fb));
        // several warps of the CTA work on different A blocks of the row and may meet in a C block: shared-memory float atomics
        // (one instruction each); a lone warp keeps its lane-private read-modify-write
        float* a0 = r.dense + c0 * 64 + lane;
        if (blockDim.x > 32) {
            atomicAdd(a0, d0); atomicAdd(a0 + 32, d1);
            if (two) { float* a1 = r.dense + c1 * 64 + lane; atomicAdd(a1, d2); atomicAdd(a1 + 32, d3); }
        } else {
            a0[0] += d0; a0[32] += d1;
            if (two) { float* a1 = r.dense + c1 * 64 + lane; a1[0] += d2; a1[32] += d3; }
        }
        i = nxt; fb = nfb; fa0 = nfa0; fa1 = nfa1; c0 = nc0; c1 = nc1; two = ntwo; cur_a = na;
    }
    __syncwarp();
}

template <int MODE>
__device__ __forceinline__ void drain(const GemmArgs& g, const RowCtx& r, const uint2* q, int n) {
    if constexpr (MODE == MODE_MMA) {
        drain_mma(g, r, q, n, reinterpret_cast<PairMeta*>(const_cast<uint2*>(q) + 64));
    } else {
        const int lane = threadIdx.x & 31;
        if (lane < n) { const uint2 e = q[lane]; process_pair<MODE>(g, r, (int)e.x, (int)e.y); }
    }
}

// Candidate walk, four B blocks per lane and step: `slots` A blocks at a time, G lanes per A block, each lane
// tests one aligned 32-bit word of B's inner-dimension masks (kmask) against the replicated mask of its A block.
// Survivors set their C block column in the row's bit set; with LIST they are also appended to the row's segment
// of the global pair list (one shared cursor bump per warp and step).
template <bool LIST, bool STATS, bool BITS = true>
__device__ __forceinline__ void enumerate_vec(const GemmArgs& g, const RowCtx& r, uint2* list, uint32_t* s_cursor,
                                              unsigned long long& n_cand, uint32_t& n_surv) {
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31;
    const int Gs = g.G, slots = T / Gs, slot = tid / Gs, gl = tid % Gs;
    const int niter = (r.a1 - r.a0 + slots - 1) / slots;
    for (int itA = 0; itA < niter; itA++) {
        const int a = r.a0 + itA * slots + slot;
        int b0 = 0, b1 = 0, rl = 0, jb = 0, wb = 0; uint32_t am4 = 0;
        if (a < r.a1) {
            const int k = g.a_bcol[a];
            b0 = g.b_brp[k]; b1 = g.b_brp[k + 1];
            am4 = (uint32_t)g.a_kmask[a] * 0x01010101u;
            rl = local_row(r.abr, r.nr, a); jb = r.jb[rl]; wb = r.wo[rl];
        }
        const int bs = b0 & ~3;
        int ngrp = (b1 - bs + 3) >> 2;           // 4-block groups of this lane's B row (0 when the row is empty)
        if (b1 <= b0) ngrp = 0;
        int maxgrp = ngrp;
#pragma unroll
        for (int o = 16; o; o >>= 1) maxgrp = max(maxgrp, __shfl_xor_sync(0xffffffffu, maxgrp, o));
        for (int grp0 = 0; grp0 < maxgrp; grp0 += Gs) {
            const int grp = grp0 + gl;
            const int bb = bs + 4 * grp;
            uint32_t nz = 0;
            if (grp < ngrp) {
                const uint32_t km4 = *reinterpret_cast<const uint32_t*>(g.b_kmask + bb);
                const int lo = max(b0 - bb, 0), hi = min(b1 - bb, 4);          // valid bytes [lo, hi)
                const uint32_t vm = (0xFFFFFFFFu << (8 * lo)) & (hi >= 4 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu << (8 * hi)));
                nz = __vcmpne4(km4 & am4, 0u) & vm;                              // 0xFF in every surviving byte
                if (STATS) { n_cand += (unsigned)(hi - lo); }
            }
            const int cnt = __popc(nz) >> 3;
            if (STATS) { n_surv += cnt; if (r.nr > 1 && cnt) atomicAdd(&r.rsurv[rl], (uint32_t)cnt); }   // single row: reduced by the caller
            uint32_t rem = nz;
            if (LIST) {
                // recompute per-byte positions without packing tricks (at most 4 survivors per lane)
                uint32_t total = 0, base = 0;
                uint32_t p4[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const bool s = (nz >> (8 * i)) & 1u;
                    const uint32_t m = __ballot_sync(0xffffffffu, s);
                    p4[i] = total + __popc(m & ((1u << lane) - 1u));
                    total += __popc(m);
                }
                if (total) {
                    if (lane == 0) base = atomicAdd(s_cursor, total);
                    base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if ((nz >> (8 * i)) & 1u) list[base + p4[i]] = make_uint2((uint32_t)a, (uint32_t)(bb + i));
                }
            }
            while (BITS && rem) {
                const int i = (__ffs(rem) - 1) >> 3;
                rem &= ~(0xFFu << (8 * i));
                const int j = g.b_bcol[bb + i] - jb;
                atomicOr(&r.bitset[wb + (j >> 5)], 1u << (j & 31));
            }
        }
    }
}

// ---- fine index of B ---------------------------------------------------------------------------------------------
// Uniform-random and R-MAT operands have about one value per 8x8 block.  The candidate scan then tests every block of B's block
// row bcol(a) against the A block (U1M: 128 tests for 16 survivors) and fetches each survivor's record with a random access: one
// 64-byte DRAM burst per pair and pass (ncu, U1M: 31 GB read by FILL, 31 GB by NUMERIC, for 12 GB of algorithmic bytes).  With
// B's blocks bucketed by inner index the survivors of an A block are the contiguous bucket(s) of its own inner index(es).
__global__ void fine_max_kernel(const uint32_t* __restrict__ cnt, int64_t n, uint32_t* __restrict__ mx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = i < n ? cnt[i] : 0u;
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v) atomicMax(mx, v);
}
__global__ void fine_count_kernel(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ kmask, int64_t nblk, uint32_t* __restrict__ cnt) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint32_t br = (uint32_t)(keys[b] >> 32);
    uint32_t m = kmask[b];
    while (m) { const int t = __ffs(m) - 1; m &= m - 1; atomicAdd(cnt + (size_t)br * 8 + t, 1u); }
}
// one thread per bucket walks its block row in order: entries of a bucket keep the block order (deterministic)
__global__ void fine_fill_kernel(const int32_t* __restrict__ brp, const uint8_t* __restrict__ kmask, const uint4* __restrict__ pm, int32_t nbr,
                                 const uint32_t* __restrict__ fptr, uint32_t* __restrict__ f_bcol, uint8_t* __restrict__ f_kmask, uint4* __restrict__ f_rec) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (int64_t)nbr * 8) return;
    const int br = (int)(u >> 3), t = (int)(u & 7);
    uint32_t w = fptr[u];
    for (int b = brp[br]; b < brp[br + 1]; b++) {
        const uint32_t km = kmask[b];
        if (!((km >> t) & 1u)) continue;
        const uint4 r0 = pm[2 * (int64_t)b], r1 = pm[2 * (int64_t)b + 1];
        f_bcol[w] = r0.z; f_kmask[w] = (uint8_t)km; f_rec[2 * (int64_t)w] = r0; f_rec[2 * (int64_t)w + 1] = r1;
        w++;
    }
}

// CTA-wide exclusive scan of popc(bmp[c]), c < n: writes the 64-bit prefix to off[c] and returns the total.
__device__ __forceinline__ uint32_t scan_popc64(const uint64_t* bmp, uint64_t* off, int n, uint32_t* s_tmp) {
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + T - 1) / T;
    const int c0 = min(tid * per, n), c1 = min(c0 + per, n);
    uint32_t s = 0;
    for (int c = c0; c < c1; c++) s += __popcll(bmp[c]);
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_tmp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t v = lane < (T >> 5) ? s_tmp[lane] : 0, vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        s_tmp[lane] = vi - v;
        if (lane == 31) s_tmp[32] = vi;
    }
    __syncthreads();
    uint32_t run = s_tmp[wid] + inc - s;
    for (int c = c0; c < c1; c++) { off[c] = run; run += __popcll(bmp[c]); }
    const uint32_t total = s_tmp[32];
    __syncthreads();
    return total;
}

// Walk the row's candidate pairs: `slots` A blocks at a time, G lanes per A block striding over the
// B block row.  MODE_SETBITS marks C block columns; the other modes compact survivors through a
// per-warp queue so that all 32 lanes work on surviving pairs.
template <int MODE, bool STATS>
__device__ __forceinline__ void enumerate_row(const GemmArgs& g, const RowCtx& r, uint2* q, unsigned long long& n_cand,
                                              unsigned long long& n_surv) {
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31;
    const int Gs = g.G, slots = T / Gs, slot = tid / Gs, gl = tid % Gs;
    int qn = 0;
    const int niter = (r.a1 - r.a0 + slots - 1) / slots;
    for (int itA = 0; itA < niter; itA++) {
        const int a = r.a0 + itA * slots + slot;
        int b0 = 0, b1 = 0; uint32_t am = 0;
        if (a < r.a1) {
            const int k = g.a_bcol[a];
            b0 = g.b_brp[k]; b1 = g.b_brp[k + 1];
            am = g.a_kmask[a];
        }
        int maxlen = b1 - b0;    // warp-uniform trip count (ballot below)
#pragma unroll
        for (int o = 16; o; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
        for (int off = 0; off < maxlen; off += Gs) {
            const int b = b0 + off + gl;
            const bool valid = b < b1;
            const bool surv = valid && (am & g.b_kmask[b]) != 0;
            if (STATS) { n_cand += valid; n_surv += surv; }
            if (MODE == MODE_SETBITS) {
                if (surv) { const int j = g.b_bcol[b] - r.jbase; atomicOr(&r.bitset[j >> 5], 1u << (j & 31)); }
            } else {
                const uint32_t m = __ballot_sync(0xffffffffu, surv);
                if (surv) q[qn + __popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)a, (uint32_t)b);
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    drain<MODE>(g, r, q, 32);
                    __syncwarp();
                    uint2 t = make_uint2(0, 0);
                    const bool mv = lane + 32 < qn;
                    if (mv) t = q[32 + lane];
                    __syncwarp();
                    if (mv) q[lane] = t;
                    qn -= 32;
                    __syncwarp();
                }
            }
        }
    }
    if (MODE != MODE_SETBITS) {
        drain<MODE>(g, r, q, qn);
        __syncwarp();
    }
}

// Walk the surviving pairs of a work item through B's fine index: 8 lanes per A block, one bucket per bit of its kmask, the
// lanes striding over the bucket, two entries in flight per lane.  A B block that shares several inner indices with the A block
// sits in several of those buckets; it is taken from the lowest one.  Buckets longer than FINE_LONG entries (the hub rows of a
// power-law B) are left to a second walk with a whole warp per bucket -- eight lanes on a 10^4-entry bucket would be the tail of
// the row.
//   FMODE 0: set the C block column's bit (f_bcol / f_kmask only: 5 bytes per pair), count survivors when STATS
//   FMODE 1: OR the pair's boolean block product into the C block's bitmap          (32-byte record)
//   FMODE 2: multiply                                                                 (32-byte record, values inline)
// CTA-wide: every thread of the CTA must call it (it synchronises).
constexpr int FINE_LONG = 256;
template <int FMODE, bool STATS>
__device__ __forceinline__ void fine_entry(const GemmArgs& g, const RowCtx& r, int a, int rl, int jb, int wb, uint64_t abmp, uint32_t aoff, uint32_t f,
                                           uint32_t& cnt) {
    if (FMODE == 0) {
        const int j = (int)g.f_bcol[f] - jb;
        atomicOr(&r.bitset[wb + (j >> 5)], 1u << (j & 31));
        cnt++;
    } else {
        PairIn in;
        in.pm = __ldg(g.f_rec + 2 * (int64_t)f);
        in.bv8 = FMODE == 2 ? __ldg(g.f_rec + 2 * (int64_t)f + 1) : make_uint4(0, 0, 0, 0);
        in.abmp = abmp; in.aoff = aoff;
        apply_pair<FMODE == 1 ? MODE_FILL : MODE_NUMERIC>(g, r, a, in);
    }
}
template <int FMODE, bool STATS>
__device__ __forceinline__ void fine_bucket(const GemmArgs& g, const RowCtx& r, int a, int rl, uint32_t am, uint32_t below, uint64_t abmp, uint32_t aoff,
                                            uint32_t f0, uint32_t f1, uint32_t first, uint32_t step, uint32_t& cnt) {
    const int jb = r.jb[rl], wb = r.wo[rl];
    uint32_t f = f0 + first;
    if constexpr (FMODE == 0) {
        for (; f < f1; f += step) {
            if (below && (g.f_kmask[f] & below)) continue;                 // already met in a lower bucket
            fine_entry<FMODE, STATS>(g, r, a, rl, jb, wb, abmp, aoff, f, cnt);
        }
    } else {
    // two records in flight per lane: the loads of entries f and f + step are issued before either is used
    for (; f + step < f1; f += 2 * step) {
        const bool s0 = !(below && (g.f_kmask[f] & below)), s1 = !(below && (g.f_kmask[f + step] & below));
        PairIn i0, i1;
        i0.abmp = i1.abmp = abmp; i0.aoff = i1.aoff = aoff;
        i0.pm = i1.pm = make_uint4(0, 0, 0, 0); i0.bv8 = i1.bv8 = make_uint4(0, 0, 0, 0);
        if (s0) { i0.pm = __ldg(g.f_rec + 2 * (int64_t)f); if (FMODE == 2) i0.bv8 = __ldg(g.f_rec + 2 * (int64_t)f + 1); }
        if (s1) { i1.pm = __ldg(g.f_rec + 2 * (int64_t)(f + step)); if (FMODE == 2) i1.bv8 = __ldg(g.f_rec + 2 * (int64_t)(f + step) + 1); }
        if (s0) apply_pair<FMODE == 1 ? MODE_FILL : MODE_NUMERIC>(g, r, a, i0);
        if (s1) apply_pair<FMODE == 1 ? MODE_FILL : MODE_NUMERIC>(g, r, a, i1);
    }
    if (f < f1 && !(below && (g.f_kmask[f] & below))) fine_entry<FMODE, STATS>(g, r, a, rl, jb, wb, abmp, aoff, f, cnt);
    }
}
template <int FMODE, bool STATS>
__device__ __forceinline__ void enumerate_fine(const GemmArgs& g, const RowCtx& r, uint32_t& n_surv) {
    constexpr int GL = 8;
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = T >> 5;
    const int slots = T / GL, slot = tid / GL, gl = tid % GL;
    bool any_long = false;
    for (int a = r.a0 + slot; a < r.a1; a += slots) {
        const int k = g.a_bcol[a];
        const uint32_t am = g.a_kmask[a];
        const int rl = local_row(r.abr, r.nr, a);
        uint64_t abmp = 0; uint32_t aoff = 0;
        if (FMODE >= 1) abmp = g.a_bmps[a];
        if (FMODE == 2) aoff = (uint32_t)g.a_off[a];
        uint32_t m = am, cnt = 0;
        while (m) {
            const int t = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t f0 = g.f_ptr[(size_t)k * 8 + t], f1 = g.f_ptr[(size_t)k * 8 + t + 1];
            if (f1 - f0 > FINE_LONG) { any_long = true; continue; }       // second walk below
            fine_bucket<FMODE, STATS>(g, r, a, rl, am, am & ((1u << t) - 1u), abmp, aoff, f0, f1, gl, GL, cnt);
        }
        if (STATS && cnt) { n_surv += cnt; if (r.nr > 1) atomicAdd(&r.rsurv[rl], cnt); }
    }
    // long buckets: warp w takes the A blocks a0 + w, a0 + w + nwarps, ...; its 32 lanes stride over the bucket.  The walk is
    // warp-uniform (every lane sees the same A block and bit), so nothing has to be handed over between lanes.
    if (!__syncthreads_or(any_long)) return;
    for (int a = r.a0 + wid; a < r.a1; a += nwarps) {
        const int k = g.a_bcol[a];
        const uint32_t am = g.a_kmask[a];
        uint32_t m = am, cnt = 0;
        int rl = -1; uint64_t abmp = 0; uint32_t aoff = 0;
        while (m) {
            const int t = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t f0 = g.f_ptr[(size_t)k * 8 + t], f1 = g.f_ptr[(size_t)k * 8 + t + 1];
            if (f1 - f0 <= FINE_LONG) continue;
            if (rl < 0) {
                rl = local_row(r.abr, r.nr, a);
                if (FMODE >= 1) abmp = g.a_bmps[a];
                if (FMODE == 2) aoff = (uint32_t)g.a_off[a];
            }
            fine_bucket<FMODE, STATS>(g, r, a, rl, am, am & ((1u << t) - 1u), abmp, aoff, f0, f1, (uint32_t)lane, 32u, cnt);
        }
        if (STATS && cnt) { n_surv += cnt; if (r.nr > 1) atomicAdd(&r.rsurv[rl], cnt); }
    }
}

template <int PASS, int MAXT>
__global__ void __launch_bounds__(MAXT) spgemm_pass_kernel(GemmArgs g) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = T >> 5;
    // shared-memory carve-up (sizes mirrored by pass_smem_bytes on the host)
    constexpr bool HAS_CBMP = PASS == PASS_FILL || PASS == PASS_NUMERIC;
    constexpr bool HAS_QUEUE = PASS == PASS_NUMERIC_MMA;
    uint64_t* s_cbmp = reinterpret_cast<uint64_t*>(smem);                                        // [cap_c]     FILL, NUMERIC
    uint2* s_queue = reinterpret_cast<uint2*>(s_cbmp + (HAS_CBMP ? g.cap_c : 0));                 // [nwarps*QSLOTS] NUMERIC_MMA
    uint32_t* s_bitset = reinterpret_cast<uint32_t*>(s_queue + (HAS_QUEUE ? nwarps * QSLOTS : 0)); // [cap_words]
    uint32_t* s_wrank = s_bitset + g.cap_words;                                                   // [cap_words] (COUNT: only for row groups)
    uint32_t* s_coff = s_wrank + ((PASS == PASS_COUNT && g.group == 1) ? 0 : g.cap_words);        // [cap_c]     NUMERIC
    float* s_acc = reinterpret_cast<float*>(s_coff + (PASS == PASS_NUMERIC ? g.cap_c : 0));       // [cap_nnz]   NUMERIC
    float* s_dense = s_acc + (PASS == PASS_NUMERIC ? g.cap_nnz : 0);                              // [cap_c*64]  NUMERIC_MMA
    uint32_t* s_tmp = reinterpret_cast<uint32_t*>(s_dense + (PASS == PASS_NUMERIC_MMA ? g.cap_c * 64 : 0)); // [34]
    int* s_row = reinterpret_cast<int*>(s_tmp + 34);
    uint32_t* s_cursor = reinterpret_cast<uint32_t*>(s_row + 1);
    int* s_batch = s_row + 2;                          // [0] next queue index of this CTA's batch, [1] its end
    // per-row tables of the work item (<= 32 rows): A block offsets, bit-set origin and word offsets, C block offsets, surviving
    // pairs, value offsets
    int* s_abr = s_row + 4;                            // [33]
    int* s_jb = s_abr + 33;                            // [32]
    int* s_wo = s_jb + 32;                             // [33]
    int* s_cbr = s_wo + 33;                            // [33]
    uint32_t* s_rsurv = reinterpret_cast<uint32_t*>(s_cbr + 33);   // [32]
    uint32_t* s_rowoff = s_rsurv + 32;                 // [33]
    if (tid == 0) { s_batch[0] = 0; s_batch[1] = 0; }
    __shared__ uint32_t s_cls[PAIR_CLASSES];           // FILL: pairs per class of the work item (zero between work items)
    if (tid < PAIR_CLASSES) s_cls[tid] = 0;

    unsigned long long n_cand = 0, n_surv_total = 0;
    int my_max1 = 0, my_max2 = 0, my_max3 = 0;         // per-thread maxima, published once when the CTA retires
    uint2* q = s_queue + wid * QSLOTS;
    const int R = g.row_list ? 1 : g.group;
    const int n_items = g.row_list ? g.n_list : (g.row_end - g.row_begin + R - 1) / R;

    while (true) {
        if (tid == 0) {
            if (s_batch[0] == s_batch[1]) { s_batch[0] = atomicAdd(g.work_counter, g.batch); s_batch[1] = s_batch[0] + g.batch; }
            const int i = s_batch[0]++;
            *s_row = i < n_items ? (g.row_list ? g.row_list[i] : g.row_begin + i * R) : g.row_end;
            *s_cursor = 0;
        }
        __syncthreads();
        RowCtx r;
        r.row = *s_row;
        __syncthreads();
        if (r.row >= g.row_end) break;
        r.nr = min(R, g.row_end - r.row);
        const int nr = r.nr;
        const int lrow = r.row - g.row_begin;
        // ---- the rows of this work item
        if (tid < 32) {
            int nw = 0;
            if (tid < nr) { const int2 ri = g.rowinfo[lrow + tid]; s_jb[tid] = ri.x; nw = ri.y; s_rsurv[tid] = 0; }
            int inc = nw;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (tid < nr) s_wo[tid + 1] = inc;
            if (tid == 0) s_wo[0] = 0;
            if (tid <= nr) {
                s_abr[tid] = g.a_brp[r.row + tid];
                if (PASS != PASS_COUNT) s_cbr[tid] = g.c_brp[lrow + tid];
                if (PASS == PASS_NUMERIC || PASS == PASS_NUMERIC_MMA) s_rowoff[tid] = (uint32_t)(g.row_nnz[lrow + tid] - g.row_nnz[lrow]);
            }
        }
        __syncthreads();
        const int nwords = s_wo[nr];
        r.a0 = s_abr[0]; r.a1 = s_abr[nr]; r.jbase = s_jb[0];
        r.abr = s_abr; r.jb = s_jb; r.wo = s_wo; r.rsurv = s_rsurv;
        if (nwords == 0) {
            if (PASS == PASS_COUNT && tid < nr) { g.row_count[lrow + tid] = 0; g.row_surv[lrow + tid] = 0; }
            if (PASS == PASS_FILL && tid < nr) g.row_nnz[lrow + tid] = 0;
            continue;
        }
        const bool wfit = nwords <= g.cap_words;
        r.bitset = wfit ? s_bitset : g.g_bitset + (size_t)blockIdx.x * g.max_words;
        r.wrank = wfit ? s_wrank : g.g_wrank + (size_t)blockIdx.x * g.max_words;
        r.cbmp = nullptr; r.coff = nullptr; r.acc = nullptr; r.dense = nullptr; r.c0 = 0;
        for (int w = tid; w < nwords; w += T) r.bitset[w] = 0;
        int ccount = 0;
        uint32_t s0 = 0, nsurv = 0;
        if (PASS != PASS_COUNT) {
            r.c0 = s_cbr[0]; ccount = s_cbr[nr] - r.c0;
            s0 = g.row_surv[lrow]; nsurv = g.row_surv[lrow + nr] - s0;
        }
        if (PASS != PASS_COUNT && ccount == 0) {
            if (PASS == PASS_FILL && tid < nr) g.row_nnz[lrow + tid] = 0;
            __syncthreads();
            continue;
        }
        uint2* list = g.surv_list + s0;
        const int c0 = r.c0;                           // s_cbr holds absolute C block offsets; c - c0 is the index inside the work item

        if (PASS == PASS_COUNT) {
            __syncthreads();
            uint32_t ns = 0;
            if (g.f_ptr) enumerate_fine<0, true>(g, r, ns);
            else enumerate_vec<false, true>(g, r, nullptr, nullptr, n_cand, ns);      // bits; survivors per row into s_rsurv
            n_surv_total += ns;
            if (nr == 1) {
#pragma unroll
                for (int o = 16; o; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
                if (lane == 0 && ns) atomicAdd(&s_rsurv[0], ns);
            }
            __syncthreads();
            const uint32_t total = rank_words(r.bitset, nr > 1 ? r.wrank : nullptr, nwords, s_tmp);
            if (tid < nr) {
                uint32_t cnt = total;
                if (nr > 1) {
                    const int w0 = s_wo[tid], w1 = s_wo[tid + 1];
                    cnt = (w1 < nwords ? r.wrank[w1] : total) - (w0 < nwords ? r.wrank[w0] : total);
                }
                g.row_count[lrow + tid] = cnt; g.row_surv[lrow + tid] = s_rsurv[tid];
                my_max3 = max(my_max3, (int)cnt);                                // largest single row (dense-block NUMERIC runs row by row)
            }
            if (tid == 0) my_max1 = max(my_max1, (int)total);
        }
        if (PASS == PASS_FILL) {
            const bool cfit = ccount <= g.cap_c;
            r.cbmp = cfit ? s_cbmp : g.c_bmps + c0;            // global C.bmps is pre-zeroed
            if (cfit) for (int c = tid; c < ccount; c += T) r.cbmp[c] = 0;
            __syncthreads();
            uint32_t ns = 0;
            if (g.f_ptr) {
                // fine index: no pair list at all -- bits from the 5-byte bucket entries, then the records once
                enumerate_fine<0, false>(g, r, ns);
                __syncthreads();
                rank_words(r.bitset, r.wrank, nwords, s_tmp);
                enumerate_fine<1, false>(g, r, ns);
            } else {
            // pair list only: a survivor's C block column comes with its packed B record (one sector, fetched here for the first time
            // and again -- from L2 -- by the pair pass below) instead of a separate random read of b_bcol per survivor
            enumerate_vec<true, false, false>(g, r, list, s_cursor, n_cand, ns);
            __syncthreads();
            for (uint32_t e = tid; e < nsurv; e += T) {
                const uint2 pr = list[e];
                const int rl = local_row(r.abr, r.nr, (int)pr.x);
                const int j = (int)__ldg(&g.b_pm[2 * (int64_t)pr.y].z) - r.jb[rl];
                atomicOr(&r.bitset[r.wo[rl] + (j >> 5)], 1u << (j & 31));
            }
            __syncthreads();
            rank_words(r.bitset, r.wrank, nwords, s_tmp);
            if (nsurv <= (uint32_t)g.sort_max) {
                // apply the pairs and leave the segment sorted by class for NUMERIC (counting sort through shared counters; a thread
                // keeps its <= SORT_K pairs in registers between reading and rewriting the segment)
                uint2 mine[SORT_K]; uint32_t slot[SORT_K];
#pragma unroll
                for (int k = 0; k < SORT_K; k++) {
                    const uint32_t e = tid + (uint32_t)k * T;
                    slot[k] = 0xFFFFFFFFu;
                    if (e < nsurv) {
                        const uint2 pr = list[e];
                        const PairIn in = load_pair<MODE_FILL>(g, (int)pr.x, (int)pr.y);
                        apply_pair<MODE_FILL>(g, r, (int)pr.x, in);
                        const int cls = pair_class(in.abmp, ((uint64_t)in.pm.y << 32) | in.pm.x);
                        mine[k] = pr;
                        slot[k] = ((uint32_t)cls << 24) | atomicAdd(&s_cls[cls], 1u);
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    uint32_t run = 0;
                    for (int c = 0; c < PAIR_CLASSES; c++) { const uint32_t t = s_cls[c]; s_cls[c] = run; run += t; }
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < SORT_K; k++) {
                    if (slot[k] != 0xFFFFFFFFu) {
                        const uint32_t cls = slot[k] >> 24;
                        uint2 pr = mine[k];
                        if (cls >= (uint32_t)PAIR_LIGHT) pr.x |= PAIR_LIGHT_FLAG;
                        list[s_cls[cls] + (slot[k] & 0xFFFFFFu)] = pr;
                    }
                }
                __syncthreads();
                if (tid < PAIR_CLASSES) s_cls[tid] = 0;
            } else {
                for (uint32_t e = tid; e < nsurv; e += T) { const uint2 pr = list[e]; process_pair<MODE_FILL>(g, r, (int)pr.x, (int)pr.y); }
            }
            }
            __syncthreads();
            // keys, bitmaps and the derived per-block arrays out: ascending (row, bit index) = ascending key
            for (int w = tid; w < nwords; w += T) {
                uint32_t word = r.bitset[w];
                if (!word) continue;
                int c = (int)r.wrank[w];
                const int rl = local_row(s_wo, nr, w);
                const int jw = s_jb[rl] + (w - s_wo[rl]) * 32;
                while (word) {
                    const int bit = __ffs(word) - 1;
                    word &= word - 1;
                    const int j = jw + bit;
                    const uint64_t bm = r.cbmp[c];
                    g.c_keys[c0 + c] = ((uint64_t)(uint32_t)(r.row + rl) << 32) | (uint32_t)j;
                    g.c_bcol[c0 + c] = j;
                    g.c_kmask[c0 + c] = (uint8_t)kmask_of(bm);
                    if (cfit) g.c_bmps[c0 + c] = bm;
                    c++;
                }
            }
            __syncthreads();
            // value offsets: scanned over the work item, then made relative to each block's own row (NUMERIC rebases them once the
            // row totals are scanned)
            const uint32_t rn = scan_popc64(r.cbmp, g.c_off + c0, ccount, s_tmp);
            if (tid <= nr) { const int cs = s_cbr[tid] - c0; s_rowoff[tid] = cs < ccount ? (uint32_t)g.c_off[c0 + cs] : rn; }
            __syncthreads();
            if (nr > 1)
                for (int c = tid; c < ccount; c += T) g.c_off[c0 + c] -= s_rowoff[local_row(s_cbr, nr, c0 + c)];
            if (tid < nr) g.row_nnz[lrow + tid] = s_rowoff[tid + 1] - s_rowoff[tid];
            if (tid == 0) my_max2 = max(my_max2, (int)min(rn, 0x7FFFFFFFu));
        }
        if (PASS == PASS_NUMERIC || PASS == PASS_NUMERIC_MMA) {
            const uint64_t vbase = g.row_nnz[lrow];
            const int64_t rownnz = (int64_t)(g.row_nnz[lrow + nr] - vbase);
            const bool fit = PASS == PASS_NUMERIC ? (ccount <= g.cap_c && rownnz <= g.cap_nnz) : (ccount <= g.cap_c);
            for (int c = tid; c < ccount; c += T) {
                const int rl = local_row(s_cbr, nr, c0 + c);
                const uint64_t local = s_rowoff[rl] + g.c_off[c0 + c];      // relative to the work item's first value
                if (PASS == PASS_NUMERIC && fit) { s_cbmp[c] = g.c_bmps[c0 + c]; s_coff[c] = (uint32_t)local; }
                g.c_off[c0 + c] = vbase + local;                             // absolute from here on
                const int j = g.c_bcol[c0 + c] - s_jb[rl];                   // bit set from C's own block columns
                atomicOr(&r.bitset[s_wo[rl] + (j >> 5)], 1u << (j & 31));
            }
            if (PASS == PASS_NUMERIC && fit) {
                r.cbmp = s_cbmp; r.coff = s_coff; r.acc = s_acc;
                for (int v = tid; v < rownnz; v += T) s_acc[v] = 0.f;
            }
            if (PASS == PASS_NUMERIC_MMA && fit) {
                r.dense = s_dense;
                for (int v = tid; v < ccount * 64; v += T) s_dense[v] = 0.f;
            }
            __syncthreads();
            rank_words(r.bitset, r.wrank, nwords, s_tmp);
            if (PASS == PASS_NUMERIC && g.f_ptr) {
                uint32_t ns = 0;
                enumerate_fine<2, false>(g, r, ns);
            } else if (PASS == PASS_NUMERIC) {
                if (g.split8) {
                    // the segment is sorted by class (FILL): h pairs with eight lanes each -- one per row of the A block --, then the
                    // pairs whose A block has one or two values, one lane each.  An unsorted segment (too many pairs for FILL's
                    // registers) carries no flags: h = nsurv.
                    uint32_t h = nsurv;
                    if (nsurv <= (uint32_t)g.sort_max) {
                        h = 0;
                        for (uint32_t e0 = 0; e0 < nsurv; e0 += T) h += (uint32_t)__syncthreads_count(e0 + tid < nsurv && !(list[e0 + tid].x & PAIR_LIGHT_FLAG));
                    }
                    for (uint64_t e = tid; e < (uint64_t)h * 8u; e += T) { const uint2 pr = list[e >> 3]; process_pair<MODE_NUMERIC>(g, r, (int)(pr.x & ~PAIR_LIGHT_FLAG), (int)pr.y, (int)(e & 7u)); }
                    for (uint32_t e = h + tid; e < nsurv; e += T) { const uint2 pr = list[e]; process_pair<MODE_NUMERIC>(g, r, (int)(pr.x & ~PAIR_LIGHT_FLAG), (int)pr.y); }
                } else if (MAXT == 1024) {            // 32 registers per thread: one pair at a time
                    for (uint32_t e = tid; e < nsurv; e += T) { const uint2 pr = list[e]; process_pair<MODE_NUMERIC>(g, r, (int)pr.x, (int)pr.y); }
                } else {
                    for (uint32_t e = tid; e < nsurv; e += 2 * T) {       // two pairs' loads in flight
                        const bool two = e + T < nsurv;
                        const uint2 p0 = list[e], p1 = two ? list[e + T] : make_uint2(0, 0);
                        const PairIn i0 = load_pair<MODE_NUMERIC>(g, (int)p0.x, (int)p0.y);
                        PairIn i1 = i0;
                        if (two) i1 = load_pair<MODE_NUMERIC>(g, (int)p1.x, (int)p1.y);
                        apply_pair<MODE_NUMERIC>(g, r, (int)p0.x, i0);
                        if (two) apply_pair<MODE_NUMERIC>(g, r, (int)p1.x, i1);
                    }
                }
            } else if (!fit) {
                // too many C blocks for the dense accumulators: scalar products into global memory.  The row's pairs are enumerated
                // again rather than read from the pair list: FILL may have written the list a group of rows at a time
                unsigned long long dc = 0, ds = 0;
                enumerate_row<MODE_NUMERIC, false>(g, r, q, dc, ds);
            } else {
                unsigned long long dc = 0, ds = 0;
                // every warp enumerates its own A blocks, so the pairs of an A block stay adjacent in its queue (two B blocks per MMA);
                // walking FILL's pair list instead of enumerating a third time measured no faster (BC4M 19.2 vs 18.6 ms)
                enumerate_row<MODE_MMA, false>(g, r, q, dc, ds);
            }
            __syncthreads();
            if (PASS == PASS_NUMERIC && fit) for (int v = tid; v < rownnz; v += T) g.c_val[vbase + v] = s_acc[v];
            if (PASS == PASS_NUMERIC_MMA && fit) {
                // compact the dense accumulators through C's bitmaps: slot L <-> cell (r = 2t, c = g), slot L+32 <-> (2t+1, g)
                const int P0 = (lane & 3) * 16 + (lane >> 2);
                for (int c = wid; c < ccount; c += nwarps) {
                    const uint64_t bmp = g.c_bmps[c0 + c];
                    float* dst = g.c_val + g.c_off[c0 + c];
                    if ((bmp >> (63 - P0)) & 1ull) dst[rank64(bmp, P0)] = s_dense[c * 64 + lane];
                    if ((bmp >> (55 - P0)) & 1ull) dst[rank64(bmp, P0 + 8)] = s_dense[c * 64 + 32 + lane];
                }
            }
        }
        __syncthreads();
    }
    if (my_max1) atomicMax(g.maxes + 1, my_max1);
    if (my_max2) atomicMax(g.maxes + 2, my_max2);
    if (my_max3) atomicMax(g.maxes + 3, my_max3);
    if (PASS == PASS_COUNT) {
#pragma unroll
        for (int o = 16; o; o >>= 1) { n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o); n_surv_total += __shfl_xor_sync(0xffffffffu, n_surv_total, o); }
        if (lane == 0) { atomicAdd(g.stats, n_cand); atomicAdd(g.stats + 1, n_surv_total); }
    }
}

// ---- dense-block numeric pass, narrow rows (block-clustered / banded operands) -------------------------------------------------
// The generic mma.sync pass above fetches every fragment element with its own rank + 2-byte global load (two dependent latencies
// per fragment, 32 scattered loads) and keeps 12 two-warp CTAs per SM busy at 19 % warps active (ncu, round 1: BC4M 18.5 ms).
// Here one warp owns a block row.  A block's compact values are read ONCE, coalesced (lane l takes values l and l + 32 of the
// block, packed in one register); a fragment element is then a rank (shift + popc) and a warp shuffle -- no dependent load.  Two
// B^t blocks are stacked as the 16x8 A-operand of mma.m16n8k8, the A block is the B-operand, D = the two 8x8 products transposed;
// each lane owns two fixed slots of every dense C block in shared memory, so the accumulation needs no atomics.  The loads of the
// next pair of B blocks are issued before the current MMA.  C block index of a pair: a per-row table over the row's column span.
struct DenseArgs {
    const int32_t* c_brp;      // [nrows+1] C block-row pointers (row - row_begin)
    const uint64_t* row_nnz;   // [nrows+1] value base of every row
    int32_t wcols;             // C block columns per window: a warp's accumulators cover wcols dense blocks (wcols * 256 bytes)
};

// values l and l + 32 of a block (cnt values from vals + off) in one register: low half = value l, high half = value l + 32
__device__ __forceinline__ uint32_t load_block_vals(const __half* __restrict__ vals, uint32_t off, int cnt, int lane) {
    const unsigned short lo = lane < cnt ? __half_as_ushort(__ldg(vals + off + lane)) : (unsigned short)0;
    const unsigned short hi = lane + 32 < cnt ? __half_as_ushort(__ldg(vals + off + 32 + lane)) : (unsigned short)0;
    return (uint32_t)lo | ((uint32_t)hi << 16);
}
// fragment register for cells p0, p0 + 1 of a block whose values sit in the warp as packed by load_block_vals
__device__ __forceinline__ uint32_t frag_shfl(uint64_t bmp, uint32_t packed, int p0) {
    const uint32_t two = (uint32_t)(bmp >> (62 - p0)) & 3u;        // bit1 = cell p0, bit0 = cell p0+1
    const int r0 = rank64(bmp, p0), r1 = r0 + (int)(two >> 1);
    const uint32_t w0 = __shfl_sync(0xffffffffu, packed, r0 & 31), w1 = __shfl_sync(0xffffffffu, packed, r1 & 31);
    const uint32_t v0 = (r0 & 32) ? (w0 >> 16) : (w0 & 0xFFFFu), v1 = (r1 & 32) ? (w1 >> 16) : (w1 & 0xFFFFu);
    return ((two & 2u) ? v0 : 0u) | (((two & 1u) ? v1 : 0u) << 16);
}

// One warp per block row; the row's C block columns are covered in windows of `wcols` columns so that a warp's dense accumulators
// stay small (wcols = 16: 4 KB per warp, 48 warps per SM; with all of a 63-column row resident the SM held 12 warps and the pass
// waited on its own loads: 23.7 ms).  Per A block the lanes test the B block row in parallel (block column inside the window,
// inner-dimension masks overlap) and the survivors are taken two at a time off the ballot.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spgemm_dense_kernel(GemmArgs g, DenseArgs da) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int wcols = da.wcols;
    float* s_dense = reinterpret_cast<float*>(smem) + (size_t)wid * wcols * 64;
    const int p0 = (lane >> 2) * 8 + (lane & 3) * 2;
    const int P0 = (lane & 3) * 16 + (lane >> 2);          // slot L <-> cell (r = 2t, c = g), slot L + 32 <-> (2t + 1, g)
    const int nrows = g.row_end - g.row_begin;
    for (;;) {
        int lrow = 0;
        if (lane == 0) lrow = atomicAdd(g.work_counter, 1);
        lrow = __shfl_sync(0xffffffffu, lrow, 0);
        if (lrow >= nrows) break;
        const int row = g.row_begin + lrow;
        const int c0 = da.c_brp[lrow], ccount = da.c_brp[lrow + 1] - c0;
        if (ccount == 0) continue;
        const uint64_t vbase = da.row_nnz[lrow];
        for (int c = lane; c < ccount; c += 32) g.c_off[c0 + c] += vbase;       // FILL left row-relative offsets: absolute from here on
        const int jfirst = g.c_bcol[c0], jlast = g.c_bcol[c0 + ccount - 1];
        const int a0 = g.a_brp[row], a1 = g.a_brp[row + 1];
        int cdone = 0;                                                           // C blocks of the windows already stored
        for (int jw = jfirst; jw <= jlast; jw += wcols) {
            for (int v = lane; v < wcols * 64; v += 32) s_dense[v] = 0.f;
            __syncwarp();
            for (int a = a0; a < a1; a++) {
                const int k = g.a_bcol[a];
                const uint32_t am = g.a_kmask[a];
                const int b0 = g.b_brp[k], b1 = g.b_brp[k + 1];
                uint32_t fb = 0; bool have_fb = false;
                for (int bb = b0; bb < b1; bb += 32) {
                    const int b = bb + lane;
                    bool in = false;
                    if (b < b1) { const int j = g.b_bcol[b] - jw; in = j >= 0 && j < wcols && (am & g.b_kmask[b]) != 0; }
                    uint32_t mask = __ballot_sync(0xffffffffu, in);
                    if (!mask) continue;
                    if (!have_fb) {
                        const uint64_t abmp = g.a_bmps[a];
                        fb = frag_shfl(abmp, load_block_vals(g.a_val, (uint32_t)g.a_off[a], __popcll(abmp), lane), p0);
                        have_fb = true;
                    }
                    // survivors two at a time; the record and the values of the next pair are loaded before this pair's MMA
                    auto take = [&](uint4& rec, uint32_t& pk) {
                        const int l = __ffs(mask) - 1;
                        mask &= mask - 1;
                        rec = __ldg(g.b_pm + 2 * (int64_t)(bb + l));
                        pk = load_block_vals(g.b_val, rec.w, __popc(rec.x) + __popc(rec.y), lane);
                    };
                    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, n0 = r0, n1 = r0; uint32_t k0 = 0, k1 = 0, m0 = 0, m1 = 0;
                    bool two = false, ntwo = false, more;
                    take(r0, k0);
                    if (mask) { take(r1, k1); two = true; }
                    for (;;) {
                        more = mask != 0;
                        if (more) { take(n0, m0); ntwo = mask != 0; if (ntwo) take(n1, m1); }
                        const uint32_t fa0 = frag_shfl(((uint64_t)r0.y << 32) | r0.x, k0, p0);
                        const uint32_t fa1 = two ? frag_shfl(((uint64_t)r1.y << 32) | r1.x, k1, p0) : 0u;
                        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                                     : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(fa0), "r"(fa1), "r"(fb));
                        float* q0 = s_dense + ((int)r0.z - jw) * 64 + lane;
                        q0[0] += d0; q0[32] += d1;
                        if (two) { float* q1 = s_dense + ((int)r1.z - jw) * 64 + lane; q1[0] += d2; q1[32] += d3; }
                        if (!more) break;
                        r0 = n0; k0 = m0; r1 = n1; k1 = m1; two = ntwo;
                    }
                }
            }
            __syncwarp();
            // the window's C blocks out: compact the dense accumulators through C's bitmaps
            while (cdone < ccount) {
                const int j = g.c_bcol[c0 + cdone] - jw;
                if (j >= wcols) break;
                const uint64_t bmp = g.c_bmps[c0 + cdone];
                float* dst = g.c_val + g.c_off[c0 + cdone];
                if ((bmp >> (63 - P0)) & 1ull) dst[rank64(bmp, P0)] = s_dense[j * 64 + lane];
                if ((bmp >> (55 - P0)) & 1ull) dst[rank64(bmp, P0 + 8)] = s_dense[j * 64 + 32 + lane];
                cdone++;
            }
            __syncwarp();
        }
    }
}

static size_t pass_smem_bytes(int pass, int T, int cap_words, int cap_c, int cap_nnz, int group) {
    size_t s = 0;
    if (pass == PASS_FILL || pass == PASS_NUMERIC) s += (size_t)cap_c * 8;
    if (pass == PASS_NUMERIC_MMA) s += (size_t)(T / 32) * QSLOTS * 8;
    s += (size_t)cap_words * ((pass == PASS_COUNT && group == 1) ? 4 : 8);
    if (pass == PASS_NUMERIC) s += (size_t)cap_c * 4 + (size_t)cap_nnz * 4;
    if (pass == PASS_NUMERIC_MMA) s += (size_t)cap_c * 64 * 4;
    s += (38 + 33 + 32 + 33 + 33 + 32 + 33) * 4 + 16;      // s_tmp, row / cursor / batch words, the per-row tables
    return s;
}

// C's derived row arrays: block-row pointers and value bases for every block row of the full matrix
// (rows outside [rb, re) are empty in a sharded product).
__global__ void finish_rows_kernel(const uint32_t* __restrict__ c_brp_local, const uint64_t* __restrict__ row_vbase, int rb, int re,
                                   int nbr, int32_t* __restrict__ brp, uint32_t* __restrict__ rvb, uint64_t* __restrict__ c_off,
                                   int64_t c_size, uint64_t c_nnz) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) c_off[c_size] = c_nnz;
    if (r > nbr) return;
    if (r < rb) { brp[r] = 0; rvb[r] = 0; }
    else if (r >= re) { brp[r] = (int32_t)c_size; rvb[r] = (uint32_t)c_nnz; }
    else { brp[r] = (int32_t)c_brp_local[r - rb]; rvb[r] = (uint32_t)row_vbase[r - rb]; }
}

}  // namespace bmsp

using namespace bmsp;

extern "C" int bmsp_debug_pair_bitmap(int64_t n, const uint64_t* a_host, const uint64_t* bt_host, uint64_t* out_host) {
    uint64_t *a = nullptr, *b = nullptr, *o = nullptr;
    cudaStream_t st = 0;
    BMSP_TRY(dev_alloc_t(&a, (size_t)n, st)); BMSP_TRY(dev_alloc_t(&b, (size_t)n, st)); BMSP_TRY(dev_alloc_t(&o, (size_t)n, st));
    BMSP_CUDA(cudaMemcpyAsync(a, a_host, 8 * n, cudaMemcpyHostToDevice, st));
    BMSP_CUDA(cudaMemcpyAsync(b, bt_host, 8 * n, cudaMemcpyHostToDevice, st));
    pair_bitmap_test_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(a, b, o, n);
    BMSP_KERNEL_CHECK();
    BMSP_CUDA(cudaMemcpyAsync(out_host, o, 8 * n, cudaMemcpyDeviceToHost, st));
    BMSP_CUDA(cudaStreamSynchronize(st));
    dev_free(a, st); dev_free(b, st); dev_free(o, st);
    return BMSP_OK;
}

template <int PASS, int MAXT>
static int launch_pass_t(const GemmArgs& g, int T, int sms, int nrows, cudaStream_t st) {
    const size_t smem = pass_smem_bytes(PASS, T, g.cap_words, g.cap_c, g.cap_nnz, g.row_list ? 1 : g.group);
    auto kern = spgemm_pass_kernel<PASS, MAXT>;
    BMSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    int occ = 0;
    BMSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    if (occ < 1) { set_error("spgemm pass %d does not fit: smem %zu", PASS, smem); return BMSP_ERR_CUDA; }
    int grid = (int)std::min<int64_t>((int64_t)sms * occ, (int64_t)nrows);
    grid = std::max(grid, 1);
    BMSP_CUDA(cudaMemsetAsync(g.work_counter, 0, sizeof(int32_t), st));
    kern<<<grid, T, smem, st>>>(g);
    BMSP_KERNEL_CHECK();
    return BMSP_OK;
}

// Heavy block rows (hub rows of power-law inputs: 10^6..10^8 candidate pairs each) are taken out of the row queue and run
// first, heaviest first, by 1024-thread CTAs on a second stream, while the 256-thread CTAs of the main launch work through the
// light rows: one CTA per row stays the unit of work, but a hub row gets four times the threads and no longer ends up as the
// tail of the launch.
struct RowSplit {
    bool active = false;
    int32_t* list = nullptr;       // device: heavy rows (descending weight) then light rows (ascending index)
    int n_heavy = 0, n_light = 0;
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    size_t heavy_scratch_off = 0;  // words: the heavy launch's slice of g_bitset / g_wrank
};

template <int PASS>
static int launch_pass(GemmArgs& g, int T, int sms, cudaStream_t st, const RowSplit& sp) {
    if (!sp.active) {
        if (T > 256) return launch_pass_t<PASS, 1024>(g, T, sms, (int)ceil_div(g.row_end - g.row_begin, g.group), st);
        return launch_pass_t<PASS, 256>(g, T, sms, (int)ceil_div(g.row_end - g.row_begin, g.group), st);
    }
    GemmArgs gh = g, gl = g;
    gh.group = gl.group = 1;
    gh.row_list = sp.list; gh.n_list = sp.n_heavy; gh.work_counter = g.work_counter + 1; gh.G = 32; gh.batch = 1;
    if (g.g_bitset) { gh.g_bitset = g.g_bitset + sp.heavy_scratch_off; gh.g_wrank = g.g_wrank + sp.heavy_scratch_off; }
    gl.row_list = sp.list + sp.n_heavy; gl.n_list = sp.n_light;
    BMSP_CUDA(cudaEventRecord(sp.fork, st));
    BMSP_CUDA(cudaStreamWaitEvent(sp.side, sp.fork, 0));
    BMSP_TRY((launch_pass_t<PASS, 1024>(gh, 1024, sms, sp.n_heavy, sp.side)));
    if (sp.n_light > 0) BMSP_TRY((launch_pass_t<PASS, 256>(gl, T, sms, sp.n_light, st)));
    BMSP_CUDA(cudaEventRecord(sp.join, sp.side));
    BMSP_CUDA(cudaStreamWaitEvent(st, sp.join, 0));
    return BMSP_OK;
}

extern "C" int bmsp_spgemm(bmsp_matrix_t A, bmsp_matrix_t Bt, const bmsp_spgemm_opts* opts, void* stream, bmsp_matrix_t* Cout,
                           bmsp_spgemm_info* info) {
    if (!A || !Bt || !Cout) { set_error("bmsp_spgemm: null argument"); return BMSP_ERR_INVALID; }
    if (A->transposed || !Bt->transposed) { set_error("bmsp_spgemm: A must be plain and B in transposed-operand form (SPGEMM.cu:1261-1262)"); return BMSP_ERR_INVALID; }
    if (A->dtype != BMSP_F16 || Bt->dtype != BMSP_F16) { set_error("bmsp_spgemm: operands must be fp16 (bmSparse_mult<half,float>)"); return BMSP_ERR_UNSUPPORTED; }
    if (A->cols != Bt->rows) { set_error("bmsp_spgemm: inner dimensions differ (%d vs %d)", A->cols, Bt->rows); return BMSP_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    touch(A, st); touch(Bt, st);
    int rb = 0, re = A->nbr;
    // brow_range_set: the range is taken literally ([r, r) multiplies nothing: a shard that owns no rows); without it the legacy
    // rule applies -- [0, 0) means all rows
    if (opts && (opts->brow_range_set || opts->brow_begin != 0 || opts->brow_end != 0)) { rb = opts->brow_begin; re = opts->brow_end; }
    if (rb < 0 || re > A->nbr || rb > re) { set_error("bmsp_spgemm: bad block-row range [%d,%d)", rb, re); return BMSP_ERR_INVALID; }
    const int nrows = re - rb;
    const bool verbose = opts && opts->verbose;

    int dev = 0, sms = 0;
    BMSP_CUDA(cudaGetDevice(&dev));
    BMSP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};       // start, after FILL, end, after COUNT
    if (verbose) for (auto& e : ev) BMSP_CUDA(cudaEventCreate(&e));
    if (verbose) BMSP_CUDA(cudaEventRecord(ev[0], st));

    // ---- scratch
    int2* rowinfo = nullptr; unsigned long long* cand = nullptr; int32_t* small = nullptr;   // small: maxes[3], counter, pad, stats[2]
    uint32_t *row_count = nullptr, *row_surv = nullptr; uint64_t* row_nnz = nullptr; uint2* surv_list = nullptr;
    bmsp_matrix_s* C = new bmsp_matrix_s();
    touch(C, st);
    C->rows = A->rows; C->cols = Bt->cols; C->dtype = BMSP_F32; C->transposed = 0;
    C->nbr = (int32_t)ceil_div(C->rows, 8);
    uint32_t *g_bitset = nullptr, *g_wrank = nullptr;
    RowSplit sp;
    int status = BMSP_OK;
    auto cleanup = [&]() {
        dev_free(sp.list, st);
        if (sp.side) { cudaStreamSynchronize(sp.side); cudaStreamDestroy(sp.side); }
        if (sp.fork) cudaEventDestroy(sp.fork);
        if (sp.join) cudaEventDestroy(sp.join);
        dev_free(rowinfo, st); dev_free(cand, st); dev_free(small, st); dev_free(row_count, st); dev_free(row_surv, st);
        dev_free(row_nnz, st); dev_free(surv_list, st); dev_free(g_bitset, st); dev_free(g_wrank, st);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
    };
    auto fail = [&](int code) { cleanup(); bmsp_destroy(C); return code; };
#define SG_TRY(x) do { status = (x); if (status != BMSP_OK) return fail(status); } while (0)
#define SG_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return fail(cuda_fail(e__, #x, __FILE__, __LINE__)); } while (0)

    SG_TRY(dev_alloc_t(&rowinfo, (size_t)nrows + 1, st));
    SG_TRY(dev_alloc_t(&cand, (size_t)nrows + 1, st));
    SG_TRY(dev_alloc_t(&small, 24, st));
    SG_TRY(dev_alloc_t(&row_count, (size_t)nrows + 2, st));
    SG_TRY(dev_alloc_t(&row_surv, (size_t)nrows + 2, st));
    SG_TRY(dev_alloc_t(&row_nnz, (size_t)nrows + 2, st));
    SG_CUDA(cudaMemsetAsync(small, 0, 24 * sizeof(int32_t), st));
    int32_t* maxes = small; int32_t* counter = small + 4; unsigned long long* stats = (unsigned long long*)(small + 8);
    unsigned long long* sum_words = (unsigned long long*)(small + 12);
    unsigned long long* max_cand = (unsigned long long*)(small + 14);
    unsigned long long* sum_cand = (unsigned long long*)(small + 16);

    int32_t h_small[24] = {0};
    if (nrows > 0) {
        rowinfo_kernel<<<(unsigned)std::min<int64_t>(ceil_div(nrows, 8), (int64_t)sms * 8), 256, 0, st>>>(A->brp, A->bcol, Bt->brp, Bt->bcol, rb, re, rowinfo, cand, maxes, sum_words, max_cand, sum_cand);
        SG_CUDA(cudaGetLastError());
    }
    SG_CUDA(cudaMemcpyAsync(h_small, small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    const int max_words = h_small[0];
    unsigned long long h_sum_words = 0;
    memcpy(&h_sum_words, h_small + 12, sizeof(h_sum_words));
    // Shared-memory capacities decide how many CTAs an SM holds, and these passes are latency-bound: size them for the rows
    // that can use them.  When every row fits, take the maxima; when the typical row would not fit anyway (R-MAT: C rows of
    // 10^4..10^5 blocks spanning every block column), keep the CTAs small -- those rows run from global scratch / global
    // atomics either way and need the occupancy (R-MAT-17 NUMERIC: 102 KB per CTA = 2 CTAs per SM, 14 % issue utilisation).
    const double avg_words = nrows ? (double)h_sum_words / nrows : 0.0;
    unsigned long long h_max_cand = 0;
    memcpy(&h_max_cand, h_small + 14, sizeof(h_max_cand));
    {
        const char* e = getenv("BMSP_SPGEMM_HEAVY");      // candidate pairs that make a row heavy (0 disables the split; tests lower it)
        const long long heavy_min = e ? atoll(e) : (1ll << 20);
        if (heavy_min > 0 && nrows > 1 && h_max_cand >= 4ull * (unsigned long long)heavy_min) {
            std::vector<unsigned long long> hc((size_t)nrows);
            SG_CUDA(cudaMemcpyAsync(hc.data(), cand, sizeof(unsigned long long) * (size_t)nrows, cudaMemcpyDeviceToHost, st));
            SG_CUDA(cudaStreamSynchronize(st));
            std::vector<int32_t> heavy, order;
            for (int i = 0; i < nrows; i++) if (hc[i] >= (unsigned long long)heavy_min) heavy.push_back(rb + i);
            std::sort(heavy.begin(), heavy.end(), [&](int32_t x, int32_t y) { return hc[x - rb] != hc[y - rb] ? hc[x - rb] > hc[y - rb] : x < y; });
            order = heavy;
            for (int i = 0; i < nrows; i++) if (hc[i] < (unsigned long long)heavy_min) order.push_back(rb + i);
            sp.n_heavy = (int)heavy.size(); sp.n_light = nrows - sp.n_heavy;
            SG_TRY(dev_alloc_t(&sp.list, (size_t)nrows + 1, st));
            SG_CUDA(cudaMemcpyAsync(sp.list, order.data(), sizeof(int32_t) * (size_t)nrows, cudaMemcpyHostToDevice, st));
            SG_CUDA(cudaStreamSynchronize(st));       // `order` dies with this scope
            SG_CUDA(cudaStreamCreateWithFlags(&sp.side, cudaStreamNonBlocking));
            SG_CUDA(cudaEventCreateWithFlags(&sp.fork, cudaEventDisableTiming));
            SG_CUDA(cudaEventCreateWithFlags(&sp.join, cudaEventDisableTiming));
            sp.active = sp.n_heavy > 0;
        }
    }

    // ---- launch shape from averages: G lanes share one A block, each lane tests 4 B blocks per step
    const double avgB = Bt->nbr ? (double)Bt->nblk / Bt->nbr : 0.0;
    const double avgA = A->nbr ? (double)A->nblk / A->nbr : 0.0;
    int G = 1;
    while (G < 32 && G * 4 < avgB) G <<= 1;
    const double avg_cand = avgA * avgB;
    // Tiny rows (stencils: 25 candidate pairs per block row) are processed 16 rows at a time by 128-thread CTAs: one warp per row
    // leaves the SM at 32 one-warp CTAs, each waiting on its own chain of dependent loads (P4096 A*A: 21.7 ms, cuSPARSE 13.0 ms)
    int group = 1;
    if (avg_cand <= 96 && !sp.active && (int64_t)16 * max_words <= 8192) group = 16;
    if (const char* e = getenv("BMSP_SPGEMM_GROUP")) { const int v = atoi(e); if (v >= 1 && v <= 16 && !sp.active && (int64_t)v * max_words <= 8192) group = v; }
    int T = group > 1 ? 128 : (avg_cand <= 96 ? 32 : (avg_cand <= 2048 ? 128 : 256));
    if (G > T) G = T;

    if (!Bt->pmeta && Bt->nblk > 0) {      // built once per B operand, reused by later products
        SG_TRY(dev_alloc((void**)&Bt->pmeta, 2 * sizeof(uint4) * (size_t)Bt->nblk, st));
        pack_meta_kernel<<<(unsigned)ceil_div(Bt->nblk, 256), 256, 0, st>>>(Bt->bmps, Bt->bcol, Bt->offsets, (const __half*)Bt->values, Bt->nnz,
                                                                          (uint4*)Bt->pmeta, Bt->nblk);
        SG_CUDA(cudaGetLastError());
    }
    // Fine index of B for sparse-block operands (about one value per block: uniform random, R-MAT): built once per B operand.
    {
        static const int fine_env = [] { const char* e = getenv("BMSP_SPGEMM_FINE"); return e ? atoi(e) : -1; }();      // 0 / 1 force it off / on
        const double dAq = A->nblk ? (double)A->nnz / A->nblk : 0.0, dBq = Bt->nblk ? (double)Bt->nnz / Bt->nblk : 0.0;
        const bool want = fine_env >= 0 ? fine_env == 1 : (dBq <= 2.0 && dAq <= 4.0);
        if (want && Bt->fine_state == 0 && Bt->nblk > 0 && (int64_t)Bt->nbr * 8 + 1 < 0x7FFFFFFFll) {
            uint32_t* fp = nullptr;
            SG_TRY(dev_alloc_t(&fp, (size_t)Bt->nbr * 8 + 2, st));
            SG_CUDA(cudaMemsetAsync(fp, 0, sizeof(uint32_t) * ((size_t)Bt->nbr * 8 + 2), st));
            fine_count_kernel<<<(unsigned)ceil_div(Bt->nblk, 256), 256, 0, st>>>(Bt->keys, Bt->kmask, Bt->nblk, fp);
            SG_CUDA(cudaGetLastError());
            // the longest bucket (read back with the total below): slot nbr * 8 + 1 of the zeroed array
            fine_max_kernel<<<(unsigned)ceil_div((int64_t)Bt->nbr * 8, 256), 256, 0, st>>>(fp, (int64_t)Bt->nbr * 8, fp + (size_t)Bt->nbr * 8 + 1);
            SG_CUDA(cudaGetLastError());
            uint32_t longest = 0;
            if (cudaMemcpyAsync(&longest, fp + (size_t)Bt->nbr * 8 + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) status = BMSP_ERR_CUDA;
            status = exclusive_scan_u32(fp, fp, (int64_t)Bt->nbr * 8, st);
            uint32_t total = 0;
            if (status == BMSP_OK && cudaMemcpyAsync(&total, fp + (size_t)Bt->nbr * 8, sizeof(uint32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) status = BMSP_ERR_CUDA;
            if (status == BMSP_OK && cudaStreamSynchronize(st) != cudaSuccess) status = BMSP_ERR_CUDA;
            if (status != BMSP_OK) { dev_free(fp, st); return fail(status); }
            // worth it while the buckets replicate little (a block with m inner indices sits in m buckets) and are of similar length:
            // a power-law B (R-MAT) has hub buckets of 10^4 entries next to a mean of 2, and the candidate scan, which spreads a B block
            // row over up to 32 lanes, handles those better (R-MAT-18 A*A: 164 ms by scan, 181 ms through the buckets)
            const double mean_bucket = (double)total / std::max<double>(1.0, (double)Bt->nbr * 8);
            const bool even = (double)longest <= 32.0 * std::max(mean_bucket, 1.0) || longest <= 512;
            if (((double)total <= 1.5 * (double)Bt->nblk && even) || fine_env == 1) {
                Bt->fine_ptr = fp; Bt->fine_n = total;
                SG_TRY(dev_alloc_t(&Bt->fine_bcol, (size_t)total + 4, st));
                SG_TRY(dev_alloc_t(&Bt->fine_kmask, (size_t)total + 16, st));
                SG_TRY(dev_alloc((void**)&Bt->fine_rec, 2 * sizeof(uint4) * ((size_t)total + 1), st));
                fine_fill_kernel<<<(unsigned)ceil_div((int64_t)Bt->nbr * 8, 128), 128, 0, st>>>(Bt->brp, Bt->kmask, (const uint4*)Bt->pmeta, Bt->nbr, fp, Bt->fine_bcol,
                                                                                             Bt->fine_kmask, (uint4*)Bt->fine_rec);
                SG_CUDA(cudaGetLastError());
                Bt->fine_state = 1;
            } else { dev_free(fp, st); Bt->fine_state = -1; }
        }
        if (!want && fine_env == 0) {}
    }
    const bool use_fine = Bt->fine_state == 1 && [] { const char* e = getenv("BMSP_SPGEMM_FINE"); return !e || atoi(e) != 0; }();
    // Walking B's fine index a row's passes wait on record loads and are held to 3-4 CTAs per SM by the row's bit set in shared
    // memory: more threads per CTA are more loads in flight for the same shared memory
    if (use_fine && T == 256 && !sp.active) {
        static const int t_env = [] { const char* e = getenv("BMSP_SPGEMM_T"); return e ? atoi(e) : 512; }();
        if (t_env == 256 || t_env == 512 || t_env == 1024) T = t_env;
    }
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    if (use_fine) { g.f_ptr = Bt->fine_ptr; g.f_bcol = Bt->fine_bcol; g.f_kmask = Bt->fine_kmask; g.f_rec = (const uint4*)Bt->fine_rec; }
    g.b_pm = (const uint4*)Bt->pmeta;
    g.a_brp = A->brp; g.a_bcol = A->bcol; g.a_bmps = A->bmps; g.a_kmask = A->kmask; g.a_off = A->offsets; g.a_val = (const __half*)A->values;
    g.b_brp = Bt->brp; g.b_bcol = Bt->bcol; g.b_bmps = Bt->bmps; g.b_kmask = Bt->kmask; g.b_off = Bt->offsets; g.b_val = (const __half*)Bt->values;
    g.rowinfo = rowinfo; g.row_begin = rb; g.row_end = re; g.G = G;
    // P4096 A*A: 2.1 M block rows of 25 candidate pairs -- one atomic on the queue head per row and pass cost more than the rows
    g.group = group;
    // (four / two lanes per pair with two / four rows each were measured on P4096: NUMERIC 8.35 / 8.97 ms against 7.45 ms with eight)
    g.split8 = A->nblk > 0 && (double)A->nnz / (double)A->nblk >= 4.0;
    {   // FILL sorts its pair-list segments by class only for the pass that reads them: the scalar NUMERIC pass with 8 lanes per pair
        const double dA0 = A->nblk ? (double)A->nnz / A->nblk : 0.0, dB0 = Bt->nblk ? (double)Bt->nnz / Bt->nblk : 0.0;
        int path0 = opts ? opts->numeric_path : -1;
        if (path0 < 0) path0 = (dA0 * dB0 / 8.0 >= 40.0) ? 1 : 0;
        static const int sort_env = [] { const char* e = getenv("BMSP_SPGEMM_SORT"); return e ? atoi(e) : 1; }();
        g.sort_max = (g.split8 && path0 == 0 && sort_env) ? SORT_K * T : 0;
    }
    g.batch = group > 1 ? 2 : (avg_cand <= 96 ? 32 : (avg_cand <= 2048 ? 4 : 1));
    g.cap_words = (int64_t)group * max_words <= 8192 ? std::max(1, group * max_words) : (avg_words > 4096.0 ? 256 : 8192);
    g.cap_c = 0; g.cap_nnz = 0;
    g.max_words = group * max_words;
    g.work_counter = counter; g.row_count = row_count; g.row_surv = row_surv; g.row_nnz = row_nnz; g.maxes = maxes; g.stats = stats;

    if (g.max_words > g.cap_words) {
        // over-cap rows use per-CTA global scratch; size it for the largest persistent grid (32 CTAs/SM)
        sp.heavy_scratch_off = (size_t)sms * 32 * g.max_words;                // + 2 heavy CTAs per SM
        const size_t n = sp.heavy_scratch_off + (size_t)sms * 2 * g.max_words;
        SG_TRY(dev_alloc_t(&g_bitset, n, st));
        SG_TRY(dev_alloc_t(&g_wrank, n, st));
        g.g_bitset = g_bitset; g.g_wrank = g_wrank;
    }

    // ---- COUNT: C blocks and surviving pairs per block row
    int64_t c_size = 0, n_surv = 0;
    if (nrows > 0) {
        SG_TRY(launch_pass<PASS_COUNT>(g, T, sms, st, sp));
        SG_TRY(exclusive_scan_u32(row_count, row_count, nrows, st));
        SG_TRY(exclusive_scan_u32(row_surv, row_surv, nrows, st));
        uint32_t tot[2] = {0, 0};
        SG_CUDA(cudaMemcpyAsync(&tot[0], row_count + nrows, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(&tot[1], row_surv + nrows, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(h_small, small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaStreamSynchronize(st));
        c_size = tot[0]; n_surv = tot[1];
    }
    if (verbose) SG_CUDA(cudaEventRecord(ev[3], st));
    unsigned long long h_stats[2];
    memcpy(h_stats, h_small + 8, sizeof(h_stats));
    if (use_fine) memcpy(&h_stats[0], h_small + 16, sizeof(unsigned long long));     // the fine walk never sees the pairs it skips: candidates from P0
    if (c_size > 0x7FFFFFFFll || h_stats[1] > 0xFFFFFFFFull) {
        set_error("product too large for one call: %lld C blocks, %llu surviving pairs (shard A's block rows with brow_begin/brow_end)",
                  (long long)c_size, h_stats[1]);
        return fail(BMSP_ERR_TOO_LARGE);
    }
    const int max_c = h_small[1];

    C->nblk = c_size; C->offsets_len = c_size + 1;
    SG_TRY(dev_alloc_t(&C->keys, (size_t)c_size + 2, st));
    SG_TRY(dev_alloc_t(&C->bmps, (size_t)c_size + 2, st));
    SG_TRY(dev_alloc_t(&C->offsets, (size_t)c_size + 2, st));
    SG_TRY(dev_alloc_t(&C->bcol, (size_t)c_size + 8, st));
    SG_TRY(dev_alloc_t(&C->kmask, (size_t)c_size + 16, st));
    SG_TRY(dev_alloc_t(&C->brp, (size_t)C->nbr + 1 + 8, st));
    SG_TRY(dev_alloc_t(&C->rvb, (size_t)C->nbr + 1 + 8, st));
    SG_TRY(dev_alloc_t(&surv_list, use_fine ? (size_t)1 : (size_t)n_surv + 1, st));
    g.c_brp = (const int32_t*)row_count; g.c_keys = C->keys; g.c_bmps = C->bmps; g.c_bcol = C->bcol; g.c_kmask = C->kmask;
    g.c_off = C->offsets; g.surv_list = surv_list;

    // ---- FILL: pair list, keys, bitmaps, row-relative value offsets
    const double avg_c = nrows ? (double)c_size / nrows : 0.0;
    auto clampi = [](double v, int lo, int hi) { return (int)std::min<double>(hi, std::max<double>(lo, v)); };
    g.cap_c = max_c <= 4096 ? std::max(1, max_c) : (avg_c <= 512.0 ? clampi(4.0 * avg_c, 256, 2048) : 64);
    if (c_size > 0) {
        if (max_c > g.cap_c) SG_CUDA(cudaMemsetAsync(C->bmps, 0, sizeof(uint64_t) * c_size, st));
        SG_TRY(launch_pass<PASS_FILL>(g, T, sms, st, sp));
    } else if (nrows > 0) {
        SG_CUDA(cudaMemsetAsync(row_nnz, 0, sizeof(uint64_t) * ((size_t)nrows + 1), st));
    }
    SG_TRY(exclusive_scan_u64(row_nnz, row_nnz, nrows, st));
    uint64_t c_nnz = 0;
    SG_CUDA(cudaMemcpyAsync(&c_nnz, row_nnz + nrows, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_small, small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    if (c_nnz > 0xFFFFFFFFull) { set_error("C has %llu values (> 2^32-1)", (unsigned long long)c_nnz); return fail(BMSP_ERR_TOO_LARGE); }
    const int max_rownnz = h_small[2];
    C->nnz = (int64_t)c_nnz;
    SG_TRY(dev_alloc(&C->values, (size_t)c_nnz * 4 + 16, st));
    if (verbose) SG_CUDA(cudaEventRecord(ev[1], st));

    // ---- NUMERIC: scalar lanes for sparse blocks, mma.sync for dense ones (about dA*dB/8 products per pair)
    g.c_val = (float*)C->values;
    const double dA = A->nblk ? (double)A->nnz / A->nblk : 0.0, dB = Bt->nblk ? (double)Bt->nnz / Bt->nblk : 0.0;
    int path = opts ? opts->numeric_path : -1;
    if (path < 0) path = (dA * dB / 8.0 >= 40.0) ? 1 : 0;
    if (c_nnz > 0) {
        static const int dense_env = [] { const char* e = getenv("BMSP_SPGEMM_DENSE"); return e ? atoi(e) : 1; }();
        const int max_c_row0 = std::max(1, h_small[3]);
        if (path == 1 && dense_env && !sp.active && (int64_t)max_words * 32 <= 1024) {
            // narrow rows of dense blocks: one warp per block row, dense accumulators for a window of C columns in shared memory
            constexpr int WARPS = 4;
            DenseArgs da;
            static const int wc_env = [] { const char* e = getenv("BMSP_SPGEMM_WCOLS"); return e ? atoi(e) : 16; }();
            da.c_brp = (const int32_t*)row_count; da.row_nnz = row_nnz; da.wcols = std::max(8, std::min(wc_env, 64));
            const size_t smem = (size_t)WARPS * (size_t)da.wcols * 256;
            auto kern = spgemm_dense_kernel<WARPS>;
            SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
            int occ = 0;
            SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
            if (occ < 1) { set_error("spgemm dense pass does not fit: smem %zu", smem); return fail(BMSP_ERR_CUDA); }
            SG_CUDA(cudaMemsetAsync(g.work_counter, 0, sizeof(int32_t), st));
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * occ, ceil_div(nrows, WARPS)));
            kern<<<grid, WARPS * 32, smem, st>>>(g, da);
            SG_CUDA(cudaGetLastError());
        } else if (path == 1) {
            const int max_c_row = std::max(1, h_small[3]);                         // recorded by COUNT
            g.cap_c = std::min(max_c_row, 192);
            if (max_c_row > g.cap_c) SG_CUDA(cudaMemsetAsync(C->values, 0, (size_t)c_nnz * 4, st));
            GemmArgs gm = g;
            gm.group = 1; gm.batch = avg_cand <= 96 ? 32 : 1;          // one warp per block row: pairs of an A block stay adjacent
            int Tm = 64;             // two warps per block row in the dense-block pass, sharing the row's accumulators (BC4M numeric, one box:
                                     // 32 threads 19.5 ms, 64: 17.7 ms, 128: 20.2 ms -- a row has ~10 A blocks x 4 lanes to hand out)
            if (const char* e = getenv("BMSP_SPGEMM_MMA_T")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128 || v == 256) Tm = v; }
            SG_TRY((launch_pass_t<PASS_NUMERIC_MMA, 256>(gm, Tm, sms, nrows, st)));
        } else {
            const double avg_nnz = nrows ? (double)c_nnz / nrows : 0.0;
            if (max_c <= 4096 && max_rownnz <= 12288) { g.cap_c = std::max(1, max_c); g.cap_nnz = std::max(1, max_rownnz); }
            else if (avg_c <= 512.0 && avg_nnz <= 1536.0) { g.cap_c = clampi(4.0 * avg_c, 256, 2048); g.cap_nnz = clampi(4.0 * avg_nnz, 1024, 8192); }
            else { g.cap_c = 64; g.cap_nnz = 256; }
            if (max_c > g.cap_c || max_rownnz > g.cap_nnz) SG_CUDA(cudaMemsetAsync(C->values, 0, (size_t)c_nnz * 4, st));
            SG_TRY(launch_pass<PASS_NUMERIC>(g, T, sms, st, sp));
        }
    }
    finish_rows_kernel<<<(unsigned)ceil_div(C->nbr + 1, 256), 256, 0, st>>>(row_count, row_nnz, rb, re, C->nbr, C->brp, C->rvb, C->offsets, c_size, c_nnz);
    SG_CUDA(cudaGetLastError());
    if (verbose) SG_CUDA(cudaEventRecord(ev[2], st));

    if (info) {
        memset(info, 0, sizeof(*info));
        info->candidate_pairs = (int64_t)h_stats[0]; info->surviving_pairs = (int64_t)h_stats[1];
        info->c_blocks = c_size; info->c_nnz = (int64_t)c_nnz; info->numeric_path = path;
        if (verbose) {
            SG_CUDA(cudaEventSynchronize(ev[2]));
            cudaEventElapsedTime(&info->symbolic_ms, ev[0], ev[1]);
            cudaEventElapsedTime(&info->numeric_ms, ev[1], ev[2]);
            cudaEventElapsedTime(&info->count_ms, ev[0], ev[3]);
            cudaEventElapsedTime(&info->fill_ms, ev[3], ev[1]);
            info->total_ms = info->symbolic_ms + info->numeric_ms;
        }
    }
    cleanup();
    *Cout = C;
    return BMSP_OK;
#undef SG_TRY
#undef SG_CUDA
}
