// cusparse_baseline.cu -- cuSPARSE 12.x comparison numbers for the same inputs (not a parity target).
// Plays the role of the reference's src/cuSparse_spmv.cu (generic-API CSR SpMV, ALG1, fp32) and of
// src/cuSparse_mult.cu (whose csrgemm2 was removed in CUDA 12: rebuilt on cusparseSpGEMM_*).  Own code.
//   cusparse_baseline spmv   <csr.bin> [reps]
//   cusparse_baseline spgemm <csr.bin> [reps] [alg]      (C = A*A; alg 0 = first of ALG1, ALG3, ALG2 that fits)
// csr.bin: int64 rows, cols, nnz; int32 row_ptr[rows+1]; int32 col_idx[nnz]; float vals[nnz]
#include <cuda_runtime.h>
#include <cusparse.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
#define CS(x) do { cusparseStatus_t s = (x); if (s != CUSPARSE_STATUS_SUCCESS) { fprintf(stderr, "cuSPARSE %s at %d\n", cusparseGetErrorString(s), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: cusparse_baseline spmv|spgemm csr.bin [reps]\n"); return 2; }
    const bool gemm = !strcmp(argv[1], "spgemm");
    const int reps = argc > 3 ? atoi(argv[3]) : 3;
    FILE* f = fopen(argv[2], "rb");
    if (!f) return 3;
    long long h[3];
    if (fread(h, 8, 3, f) != 3) return 3;
    const long long rows = h[0], cols = h[1], nnz = h[2];
    std::vector<int> rp(rows + 1), ci(nnz); std::vector<float> v(nnz);
    if (fread(rp.data(), 4, rows + 1, f) != (size_t)rows + 1 || fread(ci.data(), 4, nnz, f) != (size_t)nnz || fread(v.data(), 4, nnz, f) != (size_t)nnz) return 3;
    fclose(f);
    int *d_rp, *d_ci; float* d_v;
    CK(cudaMalloc(&d_rp, 4 * (rows + 1))); CK(cudaMalloc(&d_ci, 4 * nnz)); CK(cudaMalloc(&d_v, 4 * nnz));
    CK(cudaMemcpy(d_rp, rp.data(), 4 * (rows + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ci, ci.data(), 4 * nnz, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_v, v.data(), 4 * nnz, cudaMemcpyHostToDevice));
    cusparseHandle_t hnd; CS(cusparseCreate(&hnd));
    cusparseSpMatDescr_t A;
    CS(cusparseCreateCsr(&A, rows, cols, nnz, d_rp, d_ci, d_v, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
    const float alpha = 1.f, beta = 0.f;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (!gemm) {
        float *x, *y; CK(cudaMalloc(&x, 4 * cols)); CK(cudaMalloc(&y, 4 * rows));
        std::vector<float> ones(cols, 1.f); CK(cudaMemcpy(x, ones.data(), 4 * cols, cudaMemcpyHostToDevice));
        cusparseDnVecDescr_t X, Y; CS(cusparseCreateDnVec(&X, cols, x, CUDA_R_32F)); CS(cusparseCreateDnVec(&Y, rows, y, CUDA_R_32F));
        size_t bs = 0; void* buf = nullptr;
        CS(cusparseSpMV_bufferSize(hnd, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, X, &beta, Y, CUDA_R_32F, CUSPARSE_SPMV_CSR_ALG1, &bs));
        CK(cudaMalloc(&buf, bs + 16));
        float best = 1e30f;
        for (int r = 0; r < reps + 2; r++) {
            cudaEventRecord(e0);
            CS(cusparseSpMV(hnd, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, X, &beta, Y, CUDA_R_32F, CUSPARSE_SPMV_CSR_ALG1, buf));
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 2 && ms < best) best = ms;
        }
        std::vector<float> hy(rows); CK(cudaMemcpy(hy.data(), y, 4 * rows, cudaMemcpyDeviceToHost));
        double s = 0; for (float t : hy) s += t;
        printf("CUSPARSE_SPMV_MS %.4f rows %lld nnz %lld ysum %.6g\n", best, rows, nnz, s);
        return 0;
    }
    // SpGEMM: ALG1 (the default, fastest, memory-hungry) first; when it reports insufficient resources fall back to ALG3 / ALG2,
    // the memory-bounded variants (estimateMemory with a chunk fraction), so that every BASELINE config gets a cuSPARSE number.
    float best = 1e30f; long long c_nnz = 0;
    const int want_alg = argc > 4 ? atoi(argv[4]) : 0;
    const cusparseSpGEMMAlg_t algs[3] = {CUSPARSE_SPGEMM_ALG1, CUSPARSE_SPGEMM_ALG3, CUSPARSE_SPGEMM_ALG2};
    const int alg_id[3] = {1, 3, 2};
    int used = 0;
    const cusparseOperation_t N = CUSPARSE_OPERATION_NON_TRANSPOSE;
    auto once = [&](cusparseSpGEMMAlg_t alg, float* ms_out, long long* nnz_out, bool verbose) -> int {
        cusparseSpMatDescr_t B = A, C = nullptr;
        int* c_rp = nullptr; int* c_ci = nullptr; float* c_v = nullptr;
        void *buf1 = nullptr, *buf2 = nullptr, *buf3 = nullptr;
        cusparseSpGEMMDescr_t d = nullptr;
        int rc = 1;
        size_t b1 = 0, b2 = 0, b3 = 0;
        do {
            if (cudaMalloc(&c_rp, 4 * (rows + 1)) != cudaSuccess) break;
            if (cusparseCreateCsr(&C, rows, cols, 0, c_rp, nullptr, nullptr, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F)) break;
            if (cusparseSpGEMM_createDescr(&d)) break;
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            if (cusparseSpGEMM_workEstimation(hnd, N, N, &alpha, A, B, &beta, C, CUDA_R_32F, alg, d, &b1, nullptr)) break;
            if (cudaMalloc(&buf1, b1 + 16) != cudaSuccess) break;
            if (cusparseSpGEMM_workEstimation(hnd, N, N, &alpha, A, B, &beta, C, CUDA_R_32F, alg, d, &b1, buf1)) break;
            if (alg != CUSPARSE_SPGEMM_ALG1) {
                const float frac = 0.2f;
                if (cusparseSpGEMM_estimateMemory(hnd, N, N, &alpha, A, B, &beta, C, CUDA_R_32F, alg, d, frac, &b3, nullptr, nullptr)) break;
                if (cudaMalloc(&buf3, b3 + 16) != cudaSuccess) break;
                if (cusparseSpGEMM_estimateMemory(hnd, N, N, &alpha, A, B, &beta, C, CUDA_R_32F, alg, d, frac, &b3, buf3, &b2)) break;
                cudaFree(buf3); buf3 = nullptr;
            } else {
                if (cusparseSpGEMM_compute(hnd, N, N, &alpha, A, B, &beta, C, CUDA_R_32F, alg, d, &b2, nullptr)) break;
            }
            if (cudaMalloc(&buf2, b2 + 16) != cudaSuccess) break;
            if (cusparseSpGEMM_compute(hnd, N, N, &alpha, A, B, &beta, C, CUDA_R_32F, alg, d, &b2, buf2)) break;
            int64_t cr, cc, cn;
            if (cusparseSpMatGetSize(C, &cr, &cc, &cn)) break;
            if (cudaMalloc(&c_ci, 4 * cn + 16) != cudaSuccess || cudaMalloc(&c_v, 4 * cn + 16) != cudaSuccess) break;
            if (cusparseCsrSetPointers(C, c_rp, c_ci, c_v)) break;
            if (cusparseSpGEMM_copy(hnd, N, N, &alpha, A, B, &beta, C, CUDA_R_32F, alg, d)) break;
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) break;
            cudaEventElapsedTime(ms_out, e0, e1);
            *nnz_out = cn;
            if (verbose) printf("cusparse spgemm buffers: %.1f MB + %.1f MB\n", b1 / 1e6, b2 / 1e6);
            rc = 0;
        } while (0);
        cudaGetLastError();
        if (d) cusparseSpGEMM_destroyDescr(d);
        if (C) cusparseDestroySpMat(C);
        cudaFree(buf1); cudaFree(buf2); cudaFree(buf3); cudaFree(c_rp); cudaFree(c_ci); cudaFree(c_v);
        return rc;
    };
    for (int a = 0; a < 3 && !used; a++) {
        if (want_alg && alg_id[a] != want_alg) continue;
        float ms = 0;
        if (once(algs[a], &ms, &c_nnz, true) != 0) { fprintf(stderr, "cuSPARSE SpGEMM ALG%d failed (insufficient resources?)\n", alg_id[a]); continue; }
        used = alg_id[a];
        for (int r = 0; r < reps; r++) {
            if (once(algs[a], &ms, &c_nnz, false) != 0) { used = 0; break; }
            if (ms < best) best = ms;
        }
    }
    if (!used) { fprintf(stderr, "cuSPARSE SpGEMM: no algorithm completed\n"); return 1; }
    printf("CUSPARSE_SPGEMM_MS %.3f rows %lld nnz %lld c_nnz %lld alg %d\n", best, rows, nnz, c_nnz, used);
    return 0;
}
