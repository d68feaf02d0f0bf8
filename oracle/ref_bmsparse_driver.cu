// ref_bmsparse_driver.cu -- runs the REFERENCE's own bmSparse CUDA operators on the B200.  TEST/BASELINE ONLY.
//
// The reference sources are compiled from where they lie under /root/reference (paths injected by
// oracle/Makefile; nothing is copied; src/bmSpMatrix.cu is compiled as its own object) with main() renamed; this file adds a driver that feeds them matrices
// through the reference's own adopting constructor (src/bmSpMatrix.cu:30-43) from raw array files written by
// the tests, times exactly the region the reference's main() times (SPGEMM.cu:1274-1280, SPMV.cu:297-304) and
// dumps the result arrays so that tests can compare them bit for bit with the oracle and the B200 kernels.
//
//   ref_spgemm <A.bin> <Bt.bin> <out.bin|-> [tc_version=5] [mode=0] [reps=1] [sample_blocks=0]
//              out "-": nothing is dumped (timing / hashes only).  sample_blocks > 0: only the first, the middle and the last
//              `sample_blocks` C blocks (keys, bitmaps, offsets, values) are dumped, for element-wise value checks at sizes
//              whose full dump would be many GB.  The SHA-256 of the complete keys / bmps / offsets arrays is always printed:
//              bit-exact structure parity at full size is checked through it.
//   ref_spmv   <A.bin> <out.bin> [reps=1]           (fp32 matrix, x = ones, as the shipped main does)
//
// .bin layout: int64 rows, cols, nblk, nnz, value_bytes; keys[nblk] bmps[nblk] offsets[nblk] u64; values[nnz].
#define main reference_main
#ifdef REF_BUILD_SPGEMM
#include REF_SPGEMM_CU
#else
#include REF_SPMV_CU
#endif
#undef main
#include <cstdio>
#include <cstring>
#include <vector>
#include <thread>
#include <string>

// SHA-256 (FIPS 180-4), own compact implementation: the full-size structure arrays are compared through their digests.
namespace sha {
static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
    0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void block(uint32_t* h, const unsigned char* p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
        const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static std::string digest(const void* data, size_t n) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    const unsigned char* p = (const unsigned char*)data;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) block(h, p + i);
    unsigned char tail[128] = {0};
    const size_t r = n - i;
    memcpy(tail, p + i, r);
    tail[r] = 0x80;
    const size_t tl = r + 9 <= 64 ? 64 : 128;
    const unsigned long long bits = (unsigned long long)n * 8ull;
    for (int k = 0; k < 8; k++) tail[tl - 1 - k] = (unsigned char)(bits >> (8 * k));
    block(h, tail);
    if (tl == 128) block(h, tail + 64);
    char out[65];
    for (int k = 0; k < 8; k++) snprintf(out + 8 * k, 9, "%08x", h[k]);
    return std::string(out, 64);
}
}  // namespace sha

struct HostMat {
    long long rows, cols, nblk, nnz, vbytes;
    std::vector<uint64_t> keys, bmps, offsets;
    std::vector<char> values;
};
static bool read_mat(const char* path, HostMat& m) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return false; }
    long long h[5];
    if (fread(h, 8, 5, f) != 5) return false;
    m.rows = h[0]; m.cols = h[1]; m.nblk = h[2]; m.nnz = h[3]; m.vbytes = h[4];
    m.keys.resize(m.nblk); m.bmps.resize(m.nblk); m.offsets.resize(m.nblk); m.values.resize(m.nnz * m.vbytes);
    bool ok = fread(m.keys.data(), 8, m.nblk, f) == (size_t)m.nblk && fread(m.bmps.data(), 8, m.nblk, f) == (size_t)m.nblk &&
              fread(m.offsets.data(), 8, m.nblk, f) == (size_t)m.nblk && fread(m.values.data(), m.vbytes, m.nnz, f) == (size_t)m.nnz;
    fclose(f);
    return ok;
}
template <typename T>
static bmSpMatrix<T>* adopt(const HostMat& m) {
    thrust::device_vector<uint64_t> k(m.keys.begin(), m.keys.end()), b(m.bmps.begin(), m.bmps.end()), o(m.offsets.begin(), m.offsets.end());
    thrust::device_vector<T> v(m.nnz);
    cudaMemcpy(thrust::raw_pointer_cast(v.data()), m.values.data(), m.nnz * sizeof(T), cudaMemcpyHostToDevice);
    return new bmSpMatrix<T>((int)m.rows, (int)m.cols, (int)m.nblk, k, b, o, v);
}

int main(int argc, char** argv) {
    cudaFree(0);
#ifdef REF_BUILD_SPGEMM
    if (argc < 4) { fprintf(stderr, "usage: ref_spgemm A.bin Bt.bin out.bin|- [tc_version] [mode] [reps] [sample_blocks]\n"); return 2; }
    long tc = argc > 4 ? atol(argv[4]) : 5; long mode = argc > 5 ? atol(argv[5]) : 0; int reps = argc > 6 ? atoi(argv[6]) : 1;
    const long long sample = argc > 7 ? atoll(argv[7]) : 0;
    HostMat ha, hb;
    if (!read_mat(argv[1], ha) || !read_mat(argv[2], hb)) return 3;
    bmSpMatrix<half>* A = adopt<half>(ha);
    bmSpMatrix<half>* B = adopt<half>(hb);
    double best = 1e30;
    bmSpMatrix<float>* C = nullptr;
    for (int r = 0; r < reps; r++) {
        delete C; C = new bmSpMatrix<float>();
        cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        bmSparse_mult<half, float>(*A, *B, *C, (bool)mode, false, tc);
        cudaDeviceSynchronize();
        double us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (us < best) best = us;
    }
    cudaError_t e = cudaGetLastError();
    printf("REF_SPGEMM_US %.1f C_blocks %zu C_nnz %d cuda_status %d\n", best, C->keys.size(), C->nnz, (int)e);
    thrust::host_vector<uint64_t> k = C->keys, b = C->bmps, o = C->offsets;
    {
        // offsets: the reference's vector has C_size + 1 entries (SPGEMM.cu:1087); digest of exactly those
        std::string hk, hb, ho;
        std::thread t1([&] { hk = sha::digest(k.data(), k.size() * 8); }), t2([&] { hb = sha::digest(b.data(), k.size() * 8); }),
            t3([&] { ho = sha::digest(o.data(), o.size() * 8); });
        t1.join(); t2.join(); t3.join();
        printf("REF_SPGEMM_SHA256 keys %s bmps %s offsets %s offsets_len %zu\n", hk.c_str(), hb.c_str(), ho.c_str(), o.size());
    }
    if (strcmp(argv[3], "-")) {
        thrust::host_vector<float> v = C->values;
        FILE* f = fopen(argv[3], "wb");
        const long long nb = (long long)k.size();
        if (sample <= 0 || 3 * sample >= nb) {
            long long h[5] = {C->num_rows, C->num_cols, nb, (long long)v.size(), (long long)o.size()};
            fwrite(h, 8, 5, f);
            fwrite(k.data(), 8, k.size(), f); fwrite(b.data(), 8, k.size(), f); fwrite(o.data(), 8, o.size(), f); fwrite(v.data(), 4, v.size(), f);
        } else {
            // sampled dump: header {rows, cols, -nblk, nnz, windows = 3}, then per window {first block, blocks, first value, values}
            // followed by its keys, bitmaps, offsets (blocks + 1 entries) and values
            long long h[5] = {C->num_rows, C->num_cols, -nb, (long long)v.size(), 3};
            fwrite(h, 8, 5, f);
            const long long starts[3] = {0, (nb - sample) / 2, nb - sample};
            for (int w = 0; w < 3; w++) {
                const long long b0 = starts[w], b1 = b0 + sample;
                const long long v0 = (long long)o[b0], v1 = (long long)o[b1];
                long long wh[4] = {b0, sample, v0, v1 - v0};
                fwrite(wh, 8, 4, f);
                fwrite(k.data() + b0, 8, sample, f); fwrite(b.data() + b0, 8, sample, f); fwrite(o.data() + b0, 8, sample + 1, f);
                fwrite(v.data() + v0, 4, v1 - v0, f);
            }
        }
        fclose(f);
    }
#else
    if (argc < 3) { fprintf(stderr, "usage: ref_spmv A.bin out.bin [reps]\n"); return 2; }
    int reps = argc > 3 ? atoi(argv[3]) : 1;
    HostMat ha;
    if (!read_mat(argv[1], ha)) return 3;
    bmSpMatrix<float>* A = adopt<float>(ha);
    thrust::device_vector<float> x(ha.cols, 1.0f), y(ha.rows + 8, 0.0f);
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        bmSparse_SpMV<float, float>(*A, thrust::raw_pointer_cast(x.data()), thrust::raw_pointer_cast(y.data()), false);
        cudaDeviceSynchronize();
        double us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (us < best) best = us;
    }
    printf("REF_SPMV_US %.1f cuda_status %d\n", best, (int)cudaGetLastError());
    thrust::host_vector<float> hy = y;
    FILE* f = fopen(argv[2], "wb");
    long long n = ha.rows;
    fwrite(&n, 8, 1, f); fwrite(hy.data(), 4, n, f); fclose(f);
#endif
    return 0;
}
