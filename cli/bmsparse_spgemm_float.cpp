// bmsparse_spgemm_float -- command-line driver with the reference's calling convention and output lines, so that
// spgemm_run_batch.sh runs unchanged against libbmsparse_b200.so:
//     bmsparse_spgemm_float <MatrixFolder> <A_Matrix> <B_Matrix> [segmented tc_version verbose]
// (reference main: src/bmSparse_SPGEMM.cu:1226-1288; it reads <folder>/<name>.mtx, A plain and B in transposed-operand form,
// times bmSparse_mult with a host clock and prints "bmSparse execution", "C blocks", "C nnz").  Differences, all deliberate:
// arguments are read from the positions the batch script passes them in (the reference indexes one past);
// `segmented` / `tc_version` are accepted and ignored; verbose = 1 prints the per-phase device times under the reference's
// T_n labels (SPGEMM.cu:852-1164: T_1..T_7, T_9).  The reference's nine thrust stages are three passes here, so T_1..T_4 are
// reported as one number on the T_4 line, T_5/T_6/T_9 on the T_9 line, the merged labels as 0.  "bmSparse execution" is the
// FIRST call, as in the reference's main (SPGEMM.cu:1274-1280: it times the cold call); a second line gives the warm repeat.
// Errors are reported instead of exit() inside the library.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <cuda_runtime.h>
#include "bmSpMatrix.h"

int main(int argc, char** argv) {
    if (argc < 4) {
        std::cout << "./main MatrixFolder A_Matrix B_Matrix [segmented tc_version verbose]" << std::endl;
        return 1;
    }
    const long segmented = argc > 4 ? strtol(argv[4], NULL, 10) : 0;
    const long tc_version = argc > 5 ? strtol(argv[5], NULL, 10) : 5;
    const bool verbose = argc > 6 && *argv[6] == '1';
    const std::string A_path = std::string(argv[1]) + "/" + argv[2], B_path = std::string(argv[1]) + "/" + argv[3];
    std::cout << "A matrix: " << A_path << std::endl << "B matrix: " << B_path << std::endl;
    using clk = std::chrono::steady_clock;
    auto us = [](clk::time_point a, clk::time_point b) { return (long long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
    try {
        cudaFree(0);
        auto t0 = clk::now();
        bmSpMatrix<bmsp::half_t> A(A_path + ".mtx", false), B(B_path + ".mtx", true);
        cudaDeviceSynchronize();
        std::cout << "Parsing mtx files / Loading matrices from disk BMSP: " << us(t0, clk::now()) << " μs" << std::endl;
        bmSpMatrix<float> C;
        bmsp_spgemm_info info;
        t0 = clk::now();
        bmSparse_mult<bmsp::half_t, float>(A, B, C, segmented != 0, verbose, tc_version, &info);
        cudaDeviceSynchronize();
        const long long t_cold = us(t0, clk::now());
        bmsp_spgemm_info warm_info;
        long long t_warm;
        {   // the repeat no longer pays the one-time work (module load, memory pool growth, B's packed records)
            bmSpMatrix<float> again;
            t0 = clk::now();
            bmSparse_mult<bmsp::half_t, float>(A, B, again, segmented != 0, verbose, tc_version, &warm_info);
            cudaDeviceSynchronize();
            t_warm = us(t0, clk::now());
        }
        if (verbose) {
            auto T = [](const char* n, float ms) { std::cout << n << ": " << (long long)(ms * 1e3) << " μs " << std::endl; };
            T("T_1", 0.f); T("T_2", 0.f); T("T_3", 0.f); T("T_4", warm_info.count_ms);      // block-row spans + candidate pairs + filter: one pass
            T("T_5", 0.f); T("T_6", 0.f); T("T_9", warm_info.fill_ms);                        // ordering + unique keys + bitmaps/offsets: one pass
            T("T_7", warm_info.numeric_ms);
            std::cout << "T_1-T_6,T_9 (symbolic): " << (long long)(warm_info.symbolic_ms * 1e3) << " μs" << std::endl;
            std::cout << "T_7 (numeric): " << (long long)(warm_info.numeric_ms * 1e3) << " μs" << std::endl;
            std::cout << "Task list size: " << info.surviving_pairs << " (of " << info.candidate_pairs << " candidate pairs)" << std::endl;
        }
        std::cout << "bmSparse execution: " << t_cold << " μs" << std::endl;
        std::cout << "bmSparse execution (warm, second call): " << t_warm << " μs" << std::endl;
        std::cout << "C blocks: " << C.keys.size() << std::endl;
        std::cout << "C nnz: " << C.nnz << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
