// common.cuh -- shared internals of libbmsparse_b200 (handle layout, error plumbing, device helpers).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <cstdio>
#include <cstdarg>
#include "../../include/bmsparse_b200.h"

#define BMSP_PAD_BYTES 256   // every array is over-allocated so 16-byte bulk copies may overrun the tail

// Device-resident matrix.  keys/bmps/offsets/values are the interchange surface of the reference
// (include/bmSpMatrix.h:28-31), bit-identical; brp/bcol/rvb/kmask are derived once.
struct bmsp_matrix_s {
    int32_t rows = 0, cols = 0;
    int64_t nnz = 0, nblk = 0, offsets_len = 0;
    int32_t nbr = 0;            // block rows = ceil(rows/8)
    int32_t dtype = BMSP_F16;
    int32_t transposed = 0;
    uint64_t* keys = nullptr;
    uint64_t* bmps = nullptr;
    uint64_t* offsets = nullptr;
    void* values = nullptr;
    int32_t* brp = nullptr;     // [nbr+1] block-row pointers
    int32_t* bcol = nullptr;    // [nblk]  block column of each block
    uint32_t* rvb = nullptr;    // [nbr+1] first value index of each block row
    uint8_t* kmask = nullptr;   // [nblk]  OR of the 8 bitmap bytes: inner-dimension (k) occupancy
    void* pmeta = nullptr;      // [nblk]  uint4 {bitmap lo, hi, block column, value offset}; built by the first SpGEMM that uses this matrix as B
    // SpGEMM, sparse-block operands: B's blocks bucketed by inner index (block row k, bit t of kmask) -- spgemm.cu, "fine index"
    int32_t fine_state = 0;     // 0 not examined yet, 1 built, -1 not worth it (dense blocks: the buckets would replicate every block ~8 times)
    uint32_t* fine_ptr = nullptr;   // [nbr * 8 + 1] bucket offsets
    uint32_t* fine_bcol = nullptr;  // [fine_n] block column of the bucket entry
    uint8_t* fine_kmask = nullptr;  // [fine_n] kmask of the bucket entry's block (de-duplicates blocks that sit in several buckets)
    void* fine_rec = nullptr;       // [fine_n] 32-byte records {bitmap lo, hi, block column, value offset}, {first 8 values}
    int64_t fine_n = 0;
    // SpMV plan (spmv.cu)
    int32_t spmv_path = -2;     // -2 not planned yet, 0 row-tiled (dense-ish blocks), 1 block-parallel (sparse blocks)
    int32_t cap_blk = 0, cap_val = 0;   // per-tile smem capacities of the row-tiled kernels
    int32_t cap_lines = 0;      // staged x lines per tile (tile plan)
    int32_t tile_rows = 64;     // block rows per tile (64, 32 or 16)
    int32_t spmv_kernel = 0;    // row-tiled path: 0 = streaming kernel (persistent CTAs, staged tiles in a shared-memory ring), 1 = one CTA per tile
    int32_t xl_pitch = 32;      // elements between staged x lines (32 for the streaming kernel's bulk copies, 33 for the per-tile kernel)
    int32_t* spmv_sched = nullptr;  // streaming kernel: SCHED_SLOTS x {next tile to claim, groups done}; a launch takes the next slot
    uint32_t spmv_sched_next = 0;
    void* tile_rowpair = nullptr;   // [nbr+1] int2 (block_row_ptr, first value) zipped for one bulk copy per tile
    void* tile_desc = nullptr;  // [ntiles] TileDesc (spmv.cu): block / value / x-line ranges of every tile of 64 block rows
    uint32_t* tile_lines = nullptr;   // [nblk] distinct x lines (32 columns) of every tile, stored from the tile's first block index
    uint16_t* tile_xoff = nullptr;    // [nblk] per block: element offset of its 8-column x segment inside the tile's staged lines
    int32_t* work = nullptr;    // block-parallel work items (int4 per item)
    int32_t n_work = 0, n_split = 0;
    float* split_partial = nullptr;
    int32_t* split_rows = nullptr;
    int32_t* split_list = nullptr;   // the block rows that are sliced (fix-up kernel runs over these only)
    int32_t n_split_rows = 0;
    int32_t max_row_blocks = 0;
    void* host_pipe = nullptr;  // HostPipe (spmv.cu): streams, events and staging buffers of bmsp_spmv_host
    // multi-GPU product (spmv.cu, bmsp_spmv_halo): the tiles that depend on the peers or push rows to them ("boundary"), and the rest
    int32_t* halo_tiles = nullptr;      // [ntiles] boundary tiles first (ascending), then the interior tiles (ascending)
    int32_t halo_n_boundary = -1;       // -1: not classified yet
    int32_t halo_first = 0;             // the first boundary tile
    cudaStream_t halo_side = nullptr;   // the boundary launch runs here, next to the interior launch on the caller's stream
    cudaEvent_t halo_fork = nullptr, halo_join = nullptr;
    int32_t halo_key[2 + 16] = {0};     // own column range + push ranges the lists were built for
    // Stream ordering of the handle's memory: every entry point that enqueues work on the arrays records its stream here
    // (bmsp::touch).  bmsp_destroy frees on that stream, so the frees are ordered behind the kernels that still read or write the
    // arrays; a handle that was used on more than one stream is drained with a device synchronisation first.  Streams passed to
    // calls on a matrix must outlive it.
    cudaStream_t last_stream = nullptr;
    bool multi_stream = false, touched = false;
};

namespace bmsp {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define BMSP_CUDA(call)                                                            \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) return bmsp::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define BMSP_TRY(call)                     \
    do {                                   \
        int s__ = (call);                  \
        if (s__ != BMSP_OK) return s__;    \
    } while (0)

#define BMSP_KERNEL_CHECK() BMSP_CUDA(cudaGetLastError())

// record that work on m's arrays was enqueued on st (see bmsp_matrix_s::last_stream)
inline void touch(bmsp_matrix_s* m, cudaStream_t st) {
    if (m->touched && m->last_stream != st) m->multi_stream = true;
    m->last_stream = st; m->touched = true;
}

// stream-ordered allocation (+padding).  Freed with dev_free on the same stream.
int dev_alloc(void** p, size_t bytes, cudaStream_t st);
void dev_free(void* p, cudaStream_t st);
template <typename T>
inline int dev_alloc_t(T** p, size_t n, cudaStream_t st) { return dev_alloc((void**)p, n * sizeof(T), st); }

// out[0..n] = exclusive scan of in[0..n-1] (out has n+1 entries, out[n] = total).  in may be nullptr
// when a transform kernel filled out[0..n-1] in place (in == out allowed).
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t st);
int exclusive_scan_u64(const uint64_t* in, uint64_t* out, int64_t n, cudaStream_t st);

// derive brp/bcol/rvb/kmask from keys/bmps/offsets (matrix.cu)
int derive_compact(bmsp_matrix_s* m, cudaStream_t st);
int plan_spmv(bmsp_matrix_s* m, cudaStream_t st);
void spmv_host_release(bmsp_matrix_s* m);

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t kmask_of(uint64_t bmp) {
    uint32_t w = (uint32_t)(bmp >> 32) | (uint32_t)bmp;
    w |= w >> 16;
    w |= w >> 8;
    return w & 0xFFu;
}

__device__ __forceinline__ float val_to_f32(const __half v) { return __half2float(v); }
__device__ __forceinline__ float val_to_f32(const float v) { return v; }

// streaming (read-once) loads: keep them out of L1 and mark evict-first in L2 so that x stays resident
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
    return (uint64_t)__ldcs(reinterpret_cast<const unsigned long long*>(p));   // ld.global.cs: evict-first
}
__device__ __forceinline__ int32_t ld_stream_s32(const int32_t* p) {
    return __ldcs(p);
}

// ---- mbarrier + bulk async copy (TMA 1-D; SASS: UBLKCP / SYNCS) ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace bmsp
