/*
 * bmsp_oracle.c -- CPU restatement of the bmSparse hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA product under bmsparse_spgemm_spmv_b200/.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product never links or calls anything in oracle/.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * upstream repository GonzaBerger/bmSparse-SPGEMM-SPMV).  Nothing here is copied:
 * the reference is thrust/CUDA, this is scalar C written from the format spec.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against
 *   (i)   the only fixture the reference ships (data/real/A_matrix.mtx, B_matrix.mtx;
 *         expected vectors in tests/golden/ragusa16.json, cross-derived with scipy),
 *   (ii)  the reference's own cusp host CSR kernels compiled from the reference tree
 *         into oracle/_ref/ (see oracle/Makefile), and
 *   (iii) outputs of the reference's bmSparse CUDA binaries run on the B200 box
 *         (tests/golden/ref_cuda_golden.json, written by tests/test_gpu_reference_cuda.py on the GPU box).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef _Float16 f16;

/* ---------------------------------------------------------------- fp16 helpers */
/* text -> double -> (valueType) cast, src/bmSpMatrix.cu:136-157 (round-to-nearest-even) */
uint16_t orc_f64_to_f16_bits(double v) { f16 h = (f16)v; uint16_t b; memcpy(&b, &h, 2); return b; }
/* text -> float -> __float2half, src/reader.cu:70-75 */
uint16_t orc_f32_to_f16_bits(float v) { f16 h = (f16)v; uint16_t b; memcpy(&b, &h, 2); return b; }
float orc_f16_bits_to_f32(uint16_t b) { f16 h; memcpy(&h, &b, 2); return (float)h; }

void orc_f32_to_f16_array(int64_t n, const float *in, uint16_t *out) {
    for (int64_t i = 0; i < n; i++) out[i] = orc_f32_to_f16_bits(in[i]);
}
void orc_f16_to_f32_array(int64_t n, const uint16_t *in, float *out) {
    for (int64_t i = 0; i < n; i++) out[i] = orc_f16_bits_to_f32(in[i]);
}

/* ---------------------------------------------------------------- COO -> bmSparse */
typedef struct { int32_t r, c; float v; } coo_t;

/* block_order, src/bmSpMatrix.cu:45-74: (row/8, col/8, then row-major | col-major in block) */
static int cmp_plain(const void *a, const void *b) {
    const coo_t *x = a, *y = b;
    int xb = x->r >> 3, yb = y->r >> 3;
    if (xb != yb) return xb < yb ? -1 : 1;
    xb = x->c >> 3; yb = y->c >> 3;
    if (xb != yb) return xb < yb ? -1 : 1;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    return 0;
}
static int cmp_transposed(const void *a, const void *b) {
    const coo_t *x = a, *y = b;
    int xb = x->r >> 3, yb = y->r >> 3;
    if (xb != yb) return xb < yb ? -1 : 1;
    xb = x->c >> 3; yb = y->c >> 3;
    if (xb != yb) return xb < yb ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    return 0;
}

/*
 * COO (0-based) -> keys/bmps/offsets/values.  src/bmSpMatrix.cu:163-216:
 *   sort by block_order; key = (row/8)<<32 | col/8 (:76-83); offsets = exclusive scan of the
 *   per-key counts (:180-194); bitmap bit = 63 - pos, pos = r%8*8+c%8 or c%8*8+r%8 when
 *   transposed (:85-101); bitmaps OR-reduced per key (:208-216).  Values keep the sorted order.
 * Outputs must be sized nnz (worst case one block per entry).  Returns the block count.
 * Duplicates are NOT merged (the reference does not either; it corrupts offsets -- callers of
 * the product get BMSP_ERR_DUPLICATE instead).  values_out are fp32 (already fp16-rounded by
 * the caller when the fp16 matrix type is wanted).
 */
int64_t orc_coo_to_bmsp(int64_t nnz, const int32_t *rows, const int32_t *cols, const float *vals,
                        int transposed, uint64_t *keys, uint64_t *bmps, uint64_t *offsets,
                        float *values_out) {
    coo_t *e = (coo_t *)malloc(sizeof(coo_t) * (size_t)(nnz > 0 ? nnz : 1));
    for (int64_t i = 0; i < nnz; i++) { e[i].r = rows[i]; e[i].c = cols[i]; e[i].v = vals[i]; }
    qsort(e, (size_t)nnz, sizeof(coo_t), transposed ? cmp_transposed : cmp_plain);
    int64_t nb = -1;
    uint64_t prev = 0;
    for (int64_t i = 0; i < nnz; i++) {
        uint64_t key = ((uint64_t)(uint32_t)(e[i].r >> 3) << 32) | (uint64_t)(uint32_t)(e[i].c >> 3);
        if (nb < 0 || key != prev) {
            nb++;
            keys[nb] = key; bmps[nb] = 0; offsets[nb] = (uint64_t)i;
            prev = key;
        }
        int ri = e[i].r & 7, ci = e[i].c & 7;
        int pos = transposed ? ci * 8 + ri : ri * 8 + ci;
        bmps[nb] |= (uint64_t)1 << (63 - pos);
        values_out[i] = e[i].v;
    }
    free(e);
    return nb + 1;
}

/* ---------------------------------------------------------------- bmSparse -> COO */
/*
 * generate_coo, src/bmSpMatrix.cu:320-363: walk blocks, walk bits MSB-first, cell i of a block is
 * (i/8, i%8); values consumed in that order.  `transposed` decodes the transposed-operand form
 * (cell i is (i%8, i/8)), which the reference never decodes but Appendix A defines.
 * Output order: block order, then bit order (NOT re-sorted by (row,col); callers sort).
 */
int64_t orc_bmsp_to_coo(int64_t nblk, const uint64_t *keys, const uint64_t *bmps, int transposed,
                        int32_t *rows, int32_t *cols) {
    int64_t n = 0;
    for (int64_t b = 0; b < nblk; b++) {
        int32_t br = (int32_t)(keys[b] >> 32), bc = (int32_t)(keys[b] & 0xFFFFFFFFu);
        for (int i = 0; i < 64; i++) {
            if (bmps[b] & ((uint64_t)1 << (63 - i))) {
                int hi = i >> 3, lo = i & 7;
                rows[n] = br * 8 + (transposed ? lo : hi);
                cols[n] = bc * 8 + (transposed ? hi : lo);
                n++;
            }
        }
    }
    return n;
}

/* rank of cell p inside a block: popcll(bmp >> (64 - p)), src/bmSparse_SPMV.cu:75-78.
 * p == 0 would shift by 64 (UB in C); SURVEY Appendix A: rank 0. */
static inline int cell_rank(uint64_t bmp, int p) { return p == 0 ? 0 : __builtin_popcountll(bmp >> (64 - p)); }

/* ---------------------------------------------------------------- SpMV */
/*
 * y = A x on the bmSparse form (non-transposed A).  Semantics of spmv_kernel,
 * src/bmSparse_SPMV.cu:172-187: thread (ri,ci) of block b multiplies the value at
 * offsets[b]+rank by x[bc*8+ci]; the 8 columns are summed per row; blocks of a block row are
 * summed.  Deviations: rows of empty block rows are written as 0; indexing
 * uses true block rows; products/accumulation in double here so the test tolerance is about the
 * GPU's fp32 order, not this file's.
 */
void orc_spmv(int num_rows, int64_t nblk, const uint64_t *keys, const uint64_t *bmps,
              const uint64_t *offsets, const float *values, const float *x, double *y) {
    for (int i = 0; i < num_rows; i++) y[i] = 0.0;
    for (int64_t b = 0; b < nblk; b++) {
        int64_t br = (int64_t)(keys[b] >> 32), bc = (int64_t)(keys[b] & 0xFFFFFFFFu);
        uint64_t bmp = bmps[b];
        for (int p = 0; p < 64; p++) {
            if (bmp & ((uint64_t)1 << (63 - p))) {
                int ri = p >> 3, ci = p & 7;
                int64_t row = br * 8 + ri;
                if (row < num_rows)
                    y[row] += (double)values[offsets[b] + cell_rank(bmp, p)] * (double)x[bc * 8 + ci];
            }
        }
    }
}

/* ---------------------------------------------------------------- SpGEMM */
/*
 * bmp_calculator, src/bmSparse_SPGEMM.cu:787-810: bit (i,j) of the pair bitmap is set iff
 * (row i of the A bitmap) & (byte j of the transposed-operand B bitmap = column j of B) != 0.
 */
uint64_t orc_pair_bitmap(uint64_t a, uint64_t bt) {
    uint64_t res = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t arow = (a << (8 * i)) & 0xFF00000000000000ull;
        for (int j = 0; j < 8; j++) {
            uint64_t bcol = (bt << (8 * j)) & 0xFF00000000000000ull;
            if (arow & bcol) res |= (0x8000000000000000ull >> (i * 8)) >> j;
        }
    }
    return res;
}

typedef struct { uint64_t key; int64_t a, b; } task_t;
static int cmp_task(const void *x, const void *y) {
    const task_t *p = x, *q = y;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    if (p->a != q->a) return p->a < q->a ? -1 : 1;   /* k ascending: a fixed, documented order */
    return 0;
}

static int64_t lower_bound_u64(const uint64_t *a, int64_t n, uint64_t v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
    return lo;
}

/*
 * C = A * B with A in plain form and B in transposed-operand form, bmSparse_mult
 * src/bmSparse_SPGEMM.cu:827-1107 restated without the thrust passes:
 *   candidate pair (a,b) iff bcol(a) == brow(b)                       (:859-932, T_2/T_3)
 *   drop pairs whose pair bitmap is empty                              (:742-757, :944-948, T_4)
 *   C key = (key_a & hi32) | (key_b & lo32), ascending unique           (:111-119, T_5/T_6)
 *   C bitmap = OR of pair bitmaps                                       (:1067-1083, T_9)
 *   offsets = exclusive scan of popcount, length nblk+1                 (:1086-1096)
 *   value of cell (r,c) = sum_k A(r,k) B(k,c) over the block's pairs   (:253-287)
 * Deviation (Appendix B): B block rows are looked up by true block-row index (binary search on
 * keys), so matrices with empty block rows work.  Products and sums in double; the product of two
 * fp16 values is exact in fp32/double, so the only rounding the GPU adds is fp32 accumulation.
 *
 * Two-call protocol: call with C_keys == NULL to get sizes (*c_nblk, *c_nnz); then call again
 * with buffers (keys/bmps: nblk, offsets: nblk+1, values: nnz doubles).
 */
int orc_spgemm(int64_t a_nblk, const uint64_t *a_keys, const uint64_t *a_bmps, const uint64_t *a_off,
               const float *a_val, int64_t b_nblk, const uint64_t *b_keys, const uint64_t *b_bmps,
               const uint64_t *b_off, const float *b_val, int64_t *c_nblk, int64_t *c_nnz,
               uint64_t *C_keys, uint64_t *C_bmps, uint64_t *C_off, double *C_val) {
    /* pass 1: count surviving tasks */
    int64_t ntask = 0;
    for (int64_t a = 0; a < a_nblk; a++) {
        uint64_t k = a_keys[a] & 0xFFFFFFFFu;
        int64_t lo = lower_bound_u64(b_keys, b_nblk, k << 32), hi = lower_bound_u64(b_keys, b_nblk, (k + 1) << 32);
        for (int64_t b = lo; b < hi; b++)
            if (orc_pair_bitmap(a_bmps[a], b_bmps[b])) ntask++;
    }
    task_t *t = (task_t *)malloc(sizeof(task_t) * (size_t)(ntask > 0 ? ntask : 1));
    int64_t n = 0;
    for (int64_t a = 0; a < a_nblk; a++) {
        uint64_t k = a_keys[a] & 0xFFFFFFFFu;
        int64_t lo = lower_bound_u64(b_keys, b_nblk, k << 32), hi = lower_bound_u64(b_keys, b_nblk, (k + 1) << 32);
        for (int64_t b = lo; b < hi; b++)
            if (orc_pair_bitmap(a_bmps[a], b_bmps[b])) {
                t[n].key = (a_keys[a] & 0xFFFFFFFF00000000ull) | (b_keys[b] & 0xFFFFFFFFull);
                t[n].a = a; t[n].b = b; n++;
            }
    }
    qsort(t, (size_t)ntask, sizeof(task_t), cmp_task);
    /* structure */
    int64_t nb = 0, nnz = 0;
    for (int64_t i = 0; i < ntask;) {
        int64_t j = i; uint64_t bmp = 0;
        while (j < ntask && t[j].key == t[i].key) { bmp |= orc_pair_bitmap(a_bmps[t[j].a], b_bmps[t[j].b]); j++; }
        if (C_keys) {
            C_keys[nb] = t[i].key; C_bmps[nb] = bmp; C_off[nb] = (uint64_t)nnz;
            double acc[64];
            for (int p = 0; p < 64; p++) acc[p] = 0.0;
            for (int64_t q = i; q < j; q++) {
                uint64_t ab = a_bmps[t[q].a], bb = b_bmps[t[q].b];
                const float *av = a_val + a_off[t[q].a], *bv = b_val + b_off[t[q].b];
                for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) for (int k = 0; k < 8; k++) {
                    int pa = r * 8 + k, pb = c * 8 + k;      /* B^t: pos = col*8 + row, bmSpMatrix.cu:91-95 */
                    if ((ab >> (63 - pa)) & 1 && (bb >> (63 - pb)) & 1)
                        acc[r * 8 + c] += (double)av[cell_rank(ab, pa)] * (double)bv[cell_rank(bb, pb)];
                }
            }
            int64_t w = nnz;
            for (int p = 0; p < 64; p++) if ((bmp >> (63 - p)) & 1) C_val[w++] = acc[p];
        }
        nnz += __builtin_popcountll(bmp);
        nb++;
        i = j;
    }
    if (C_keys) C_off[nb] = (uint64_t)nnz;
    *c_nblk = nb; *c_nnz = nnz;
    free(t);
    return 0;
}

/* ---------------------------------------------------------------- cusp host CSR kernels ("port") */
/*
 * CSR SpMV, cusp/cusp/system/detail/sequential/multiply/csr_spmv.h:56-73: row loop, accumulator
 * starts at initialize(y[i]) == 0 (generic/multiply.inl:93-105), reduce = plus, combine = multiplies.
 * threads > 1 follows cusp/cusp/system/omp/detail/multiply/csr_spmv.h (parallel for over rows).
 */
void orc_csr_spmv(int num_rows, const int32_t *rp, const int32_t *ci, const float *v, const float *x,
                  float *y, int threads) {
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
#endif
    for (int i = 0; i < num_rows; i++) {
        float acc = 0.0f;
        for (int32_t jj = rp[i]; jj < rp[i + 1]; jj++) acc = acc + v[jj] * x[ci[jj]];
        y[i] = acc;
    }
}

/*
 * CSR SpGEMM, two-pass Gustavson.  Sequential flavour restates
 * cusp/cusp/system/detail/sequential/multiply/csr_spgemm.h:39-72 (pass 1: per-column mask count)
 * and :79-156 (pass 2: sums[] + next[] linked list per row, explicit zeros dropped :135, columns
 * come out in reverse discovery order, i.e. unsorted :153).  drop_zeros=0 gives the OMP flavour
 * (cusp/cusp/system/omp/detail/multiply/csr_spgemm.h:92-150 keeps zeros).
 * Two-call protocol: C_ci == NULL -> returns the pass-1 upper bound in *c_nnz and fills C_rp with
 * per-row bounds (prefix summed); second call fills C_ci/C_v and rewrites C_rp/*c_nnz exactly.
 */
int orc_csr_spgemm(int a_rows, int b_cols, const int32_t *a_rp, const int32_t *a_ci, const float *a_v,
                   const int32_t *b_rp, const int32_t *b_ci, const float *b_v, int64_t *c_nnz,
                   int32_t *C_rp, int32_t *C_ci, float *C_v, int drop_zeros, int threads) {
    int nt = threads > 0 ? threads : 1;
    if (!C_ci) {
        C_rp[0] = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
        {
            int32_t *mask = (int32_t *)malloc(sizeof(int32_t) * (size_t)(b_cols > 0 ? b_cols : 1));
            for (int k = 0; k < b_cols; k++) mask[k] = -1;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
            for (int i = 0; i < a_rows; i++) {
                int32_t cnt = 0;
                for (int32_t jj = a_rp[i]; jj < a_rp[i + 1]; jj++) {
                    int32_t j = a_ci[jj];
                    for (int32_t kk = b_rp[j]; kk < b_rp[j + 1]; kk++) {
                        int32_t k = b_ci[kk];
                        if (mask[k] != i) { mask[k] = i; cnt++; }
                    }
                }
                C_rp[i + 1] = cnt;
            }
            free(mask);
        }
        int64_t tot = 0;
        for (int i = 0; i < a_rows; i++) { tot += C_rp[i + 1]; C_rp[i + 1] = (int32_t)tot; }
        *c_nnz = tot;
        return 0;
    }
    /* pass 2 writes each row at its pass-1 offset, then (drop_zeros) compacts */
    int32_t *row_len = (int32_t *)malloc(sizeof(int32_t) * (size_t)(a_rows > 0 ? a_rows : 1));
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
        int32_t *next = (int32_t *)malloc(sizeof(int32_t) * (size_t)(b_cols > 0 ? b_cols : 1));
        float *sums = (float *)malloc(sizeof(float) * (size_t)(b_cols > 0 ? b_cols : 1));
        for (int k = 0; k < b_cols; k++) { next[k] = -1; sums[k] = 0.0f; }
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int i = 0; i < a_rows; i++) {
            int32_t head = -2, length = 0;
            for (int32_t jj = a_rp[i]; jj < a_rp[i + 1]; jj++) {
                int32_t j = a_ci[jj]; float v = a_v[jj];
                for (int32_t kk = b_rp[j]; kk < b_rp[j + 1]; kk++) {
                    int32_t k = b_ci[kk];
                    sums[k] = sums[k] + v * b_v[kk];
                    if (next[k] == -1) { next[k] = head; head = k; length++; }
                }
            }
            int32_t w = C_rp[i], w0 = w;
            for (int32_t jj = 0; jj < length; jj++) {
                if (!drop_zeros || sums[head] != 0.0f) { C_ci[w] = head; C_v[w] = sums[head]; w++; }
                int32_t tmp = head; head = next[head];
                next[tmp] = -1; sums[tmp] = 0.0f;
            }
            row_len[i] = w - w0;
        }
        free(next); free(sums);
    }
    /* compact rows (sequential, cheap) */
    int64_t w = 0;
    for (int i = 0; i < a_rows; i++) {
        int32_t s = C_rp[i];
        if (w != s) { memmove(C_ci + w, C_ci + s, sizeof(int32_t) * (size_t)row_len[i]); memmove(C_v + w, C_v + s, sizeof(float) * (size_t)row_len[i]); }
        C_rp[i] = (int32_t)w;
        w += row_len[i];
    }
    C_rp[a_rows] = (int32_t)w;
    *c_nnz = w;
    free(row_len);
    return 0;
}

/* ---------------------------------------------------------------- gallery */
/*
 * poisson5pt(m,n): cusp/cusp/gallery/detail/poisson.inl:28-46 with the stencil assembly of
 * stencil.inl:33-63,114-134: grid index = x + m*y, diagonal 4, the four neighbours -1 kept iff
 * inside the grid; CSR columns ascending.  rp has m*n+1 entries; returns nnz.
 */
int64_t orc_poisson5pt(int m, int n, int32_t *rp, int32_t *ci, float *v) {
    int64_t w = 0;
    for (int y = 0; y < n; y++) for (int x = 0; x < m; x++) {
        int64_t i = (int64_t)x + (int64_t)m * y;
        rp[i] = (int32_t)w;
        if (y > 0)     { ci[w] = (int32_t)(i - m); v[w++] = -1.0f; }
        if (x > 0)     { ci[w] = (int32_t)(i - 1); v[w++] = -1.0f; }
        ci[w] = (int32_t)i; v[w++] = 4.0f;
        if (x < m - 1) { ci[w] = (int32_t)(i + 1); v[w++] = -1.0f; }
        if (y < n - 1) { ci[w] = (int32_t)(i + m); v[w++] = -1.0f; }
    }
    rp[(int64_t)m * n] = (int32_t)w;
    return w;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Launchers such as torch.distributed.run export OMP_NUM_THREADS=1 to their workers; the CPU baseline must still use the
 * box's cores.  Sets the OpenMP thread count of the calling thread's subsequent parallel regions (the reference's cusp OMP
 * kernels in oracle/_ref share this OpenMP runtime) and returns the value now in effect. */
int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ================================================================ OpenMP variants for full-size parity (test infrastructure)
 * Same definitions as orc_coo_to_bmsp / orc_spgemm above (which follow the reference line by line and are pinned to its golden
 * vectors), evaluated one block row at a time so that the BASELINE.json sizes finish in seconds on the GPU box's host cores.
 * tests/test_oracle_golden.py checks them bit for bit against the scalar functions. */

/* CSR (columns ascending in each row) -> bmSparse.  A block row is the 8-way merge of its rows: blocks in ascending block
 * column (src/bmSpMatrix.cu:45-74, block_order), cells of a block in row-major (plain) or column-major (transposed operand)
 * order (:85-101).  Two-call protocol: keys == NULL returns the block count; then fill (keys/bmps/offsets: nblk, values: nnz). */
int64_t orc_csr_to_bmsp_omp(int rows, const int32_t *rp, const int32_t *ci, const float *vals, int transposed,
                            uint64_t *keys, uint64_t *bmps, uint64_t *offsets, float *values_out) {
    const int nbr = (rows + 7) / 8;
    int64_t *cnt = (int64_t *)calloc((size_t)nbr + 1, sizeof(int64_t));
#pragma omp parallel for schedule(dynamic, 256)
    for (int br = 0; br < nbr; br++) {
        int64_t p[8], e[8];
        for (int q = 0; q < 8; q++) { int r = br * 8 + q; p[q] = r < rows ? rp[r] : 0; e[q] = r < rows ? rp[r + 1] : 0; }
        int64_t n = 0;
        for (;;) {
            int64_t bc = -1;
            for (int q = 0; q < 8; q++) if (p[q] < e[q]) { int64_t c = ci[p[q]] >> 3; if (bc < 0 || c < bc) bc = c; }
            if (bc < 0) break;
            for (int q = 0; q < 8; q++) while (p[q] < e[q] && (ci[p[q]] >> 3) == bc) p[q]++;
            n++;
        }
        cnt[br + 1] = n;
    }
    for (int br = 0; br < nbr; br++) cnt[br + 1] += cnt[br];
    const int64_t nblk = cnt[nbr];
    if (!keys) { free(cnt); return nblk; }
#pragma omp parallel for schedule(dynamic, 256)
    for (int br = 0; br < nbr; br++) {
        int64_t p[8], e[8];
        for (int q = 0; q < 8; q++) { int r = br * 8 + q; p[q] = r < rows ? rp[r] : 0; e[q] = r < rows ? rp[r + 1] : 0; }
        int64_t b = cnt[br];
        int64_t w = br * 8 < rows ? rp[br * 8] : 0;         /* values of a block row are contiguous: they start at its first row */
        for (;;) {
            int64_t bc = -1;
            for (int q = 0; q < 8; q++) if (p[q] < e[q]) { int64_t c = ci[p[q]] >> 3; if (bc < 0 || c < bc) bc = c; }
            if (bc < 0) break;
            uint64_t bmp = 0;
            int64_t s[8], t[8];
            for (int q = 0; q < 8; q++) { s[q] = p[q]; while (p[q] < e[q] && (ci[p[q]] >> 3) == bc) p[q]++; t[q] = p[q]; }
            keys[b] = ((uint64_t)(uint32_t)br << 32) | (uint64_t)(uint32_t)bc; offsets[b] = (uint64_t)w;
            if (!transposed) {
                for (int q = 0; q < 8; q++) for (int64_t i = s[q]; i < t[q]; i++) { bmp |= (uint64_t)1 << (63 - (q * 8 + (ci[i] & 7))); values_out[w++] = vals[i]; }
            } else {
                for (int c = 0; c < 8; c++) for (int q = 0; q < 8; q++) {
                    /* at most one entry of row q has column bc*8+c */
                    for (int64_t i = s[q]; i < t[q]; i++) if ((ci[i] & 7) == c) { bmp |= (uint64_t)1 << (63 - (c * 8 + q)); values_out[w++] = vals[i]; }
                }
            }
            bmps[b] = bmp;
            b++;
        }
    }
    free(cnt);
    return nblk;
}

typedef struct { uint64_t key; int32_t a, b; } rtask_t;
static int cmp_rtask(const void *x, const void *y) {
    const rtask_t *p = x, *q = y;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    if (p->a != q->a) return p->a < q->a ? -1 : 1;
    return 0;
}

/* C = A * B^t-operand, one A block row per task (the C blocks of different block rows are independent).  Same candidate rule,
 * filter, key, bitmap and per-cell sum as orc_spgemm; the pairs of a C block are summed in ascending A-block order in double and
 * stored as fp32.  a_brp: block-row pointers of A, nbr + 1 entries (derived from a_keys by the caller).
 * Two-call protocol: C_keys == NULL fills row_blk / row_nnz (nbr + 1 entries each, exclusive prefix sums, [nbr] = totals). */
int orc_spgemm_omp(int32_t nbr, const int64_t *a_brp, const uint64_t *a_keys, const uint64_t *a_bmps, const uint64_t *a_off,
                   const float *a_val, int64_t b_nblk, const uint64_t *b_keys, const uint64_t *b_bmps, const uint64_t *b_off,
                   const float *b_val, int64_t *row_blk, int64_t *row_nnz, uint64_t *C_keys, uint64_t *C_bmps, uint64_t *C_off,
                   float *C_val) {
    const int fill = C_keys != NULL;
    int failed = 0;
#pragma omp parallel
    {
        rtask_t *t = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int32_t br = 0; br < nbr; br++) {
            int64_t n = 0;
            for (int64_t a = a_brp[br]; a < a_brp[br + 1]; a++) {
                const uint64_t k = a_keys[a] & 0xFFFFFFFFu;
                const int64_t lo = lower_bound_u64(b_keys, b_nblk, k << 32), hi = lower_bound_u64(b_keys, b_nblk, (k + 1) << 32);
                if (n + (hi - lo) > cap) {
                    cap = (n + (hi - lo)) * 2 + 64;
                    t = (rtask_t *)realloc(t, sizeof(rtask_t) * (size_t)cap);
                    if (!t) { failed = 1; cap = 0; n = 0; break; }
                }
                for (int64_t b = lo; b < hi; b++)
                    if (orc_pair_bitmap(a_bmps[a], b_bmps[b])) {
                        t[n].key = (a_keys[a] & 0xFFFFFFFF00000000ull) | (b_keys[b] & 0xFFFFFFFFull);
                        t[n].a = (int32_t)a; t[n].b = (int32_t)b; n++;
                    }
            }
            qsort(t, (size_t)n, sizeof(rtask_t), cmp_rtask);
            int64_t nb = 0, nnz = 0;
            const int64_t b0 = fill ? row_blk[br] : 0, v0 = fill ? row_nnz[br] : 0;
            for (int64_t i = 0; i < n;) {
                int64_t j = i; uint64_t bmp = 0;
                while (j < n && t[j].key == t[i].key) { bmp |= orc_pair_bitmap(a_bmps[t[j].a], b_bmps[t[j].b]); j++; }
                if (fill) {
                    C_keys[b0 + nb] = t[i].key; C_bmps[b0 + nb] = bmp; C_off[b0 + nb] = (uint64_t)(v0 + nnz);
                    double acc[64];
                    for (int p = 0; p < 64; p++) acc[p] = 0.0;
                    for (int64_t q = i; q < j; q++) {
                        const uint64_t ab = a_bmps[t[q].a], bb = b_bmps[t[q].b];
                        const float *av = a_val + a_off[t[q].a], *bv = b_val + b_off[t[q].b];
                        uint64_t rem = ab;
                        int ka = 0;
                        while (rem) {                                  /* A cells in bitmap order: value index ka */
                            const int pa = __builtin_clzll(rem);
                            rem &= ~(0x8000000000000000ull >> pa);
                            const int r = pa >> 3, k = pa & 7;
                            const double aval = (double)av[ka++];
                            for (int c = 0; c < 8; c++) {
                                const int pb = c * 8 + k;               /* B^t: pos = col*8 + row, bmSpMatrix.cu:91-95 */
                                if ((bb >> (63 - pb)) & 1) acc[r * 8 + c] += aval * (double)bv[cell_rank(bb, pb)];
                            }
                        }
                    }
                    int64_t w = v0 + nnz;
                    for (int p = 0; p < 64; p++) if ((bmp >> (63 - p)) & 1) C_val[w++] = (float)acc[p];
                }
                nnz += __builtin_popcountll(bmp);
                nb++;
                i = j;
            }
            if (!fill) { row_blk[br + 1] = nb; row_nnz[br + 1] = nnz; }
        }
        free(t);
    }
    if (failed) return 1;
    if (!fill) {
        row_blk[0] = 0; row_nnz[0] = 0;
        for (int32_t br = 0; br < nbr; br++) { row_blk[br + 1] += row_blk[br]; row_nnz[br + 1] += row_nnz[br]; }
    } else {
        C_off[row_blk[nbr]] = (uint64_t)row_nnz[nbr];
    }
    return 0;
}

/* y = A x, block rows in parallel (same per-row arithmetic as orc_spmv: double products summed in block / bit order) */
void orc_spmv_omp(int num_rows, int32_t nbr, const int64_t *brp, const uint64_t *keys, const uint64_t *bmps, const uint64_t *offsets,
                  const float *values, const float *x, double *y) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int32_t br = 0; br < nbr; br++) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int64_t b = brp[br]; b < brp[br + 1]; b++) {
            const int64_t bc = (int64_t)(keys[b] & 0xFFFFFFFFu);
            uint64_t rem = bmps[b];
            int64_t k = (int64_t)offsets[b];
            while (rem) {
                const int p = __builtin_clzll(rem);
                rem &= ~(0x8000000000000000ull >> p);
                acc[p >> 3] += (double)values[k++] * (double)x[bc * 8 + (p & 7)];
            }
        }
        for (int q = 0; q < 8; q++) if ((int64_t)br * 8 + q < num_rows) y[(int64_t)br * 8 + q] = acc[q];
    }
}
