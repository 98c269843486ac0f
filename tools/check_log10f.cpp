// check_log10f.cpp -- EXHAUSTIVE host check of csrc/phmm_log10.h against the running libm's log10f:
// every non-negative float bit pattern (0 .. 0x7f800000 inclusive: zero, subnormals, normals, +inf), bitwise.
//   g++ -O2 -std=c++17 -fopenmp -mfma -ffp-contract=off tools/check_log10f.cpp -o /tmp/check_log10f && /tmp/check_log10f
// Prints the number of mismatches (0 expected on any x86-64 host whose glibc selects the FMA variant of logf).
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include "../gatk-haplotypecaller-cpp17_b200/csrc/phmm_log10.h"

int main(int argc, char** argv)
{
    const bool quick = argc > 1 && std::string(argv[1]) == "--quick";     // every 257th float, 1/20 of the doubles
    const long long fstep = quick ? 257 : 1;
    unsigned long long mism = 0, n = 0;
    uint32_t first_bad = 0;
#pragma omp parallel for reduction(+ : mism, n) schedule(static, 1 << 20)
    for (long long u = 0; u <= 0x7f800000ll; u += fstep) {
        float x; uint32_t b = (uint32_t)u; std::memcpy(&x, &b, 4);
        const float want = log10f(x), got = phmm::glibc_log10f(x);
        uint32_t wb, gb; std::memcpy(&wb, &want, 4); std::memcpy(&gb, &got, 4);
        n++;
        if (wb != gb) { mism++; first_bad = b; }
    }
    std::printf("{\"checked\": %llu, \"mismatches\": %llu, \"example_bits\": \"0x%08x\", \"libm\": \"%s\"}\n", n, mism, first_bad,
#ifdef __GLIBC__
                "glibc " 
#endif
                "");
    // double precision: sampled (2^64 inputs cannot be enumerated).  Every binade gets random mantissas, the
    // neighbourhood of 1.0 (the separate polynomial of __log) and the normalised range [0.5, 2) get dense sweeps.
    unsigned long long mism64 = 0, n64 = 0;
    const long long per_binade = quick ? 20000 : 400000;
#pragma omp parallel for reduction(+ : mism64, n64) schedule(dynamic, 8)
    for (int e = 0; e <= 0x7fe; e++) {
        uint64_t st = 0x9e3779b97f4a7c15ull * (uint64_t)(e + 1);
        for (long long j = 0; j < per_binade; j++) {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            const uint64_t b = ((uint64_t)e << 52) | (st & 0x000fffffffffffffull);
            double x; std::memcpy(&x, &b, 8);
            const double want = log10(x), got = phmm::glibc_log10(x);
            n64++;
            if (std::memcmp(&want, &got, 8)) mism64++;
        }
    }
#pragma omp parallel for reduction(+ : mism64, n64) schedule(static, 1 << 16)
    for (long long j = 0; j < (quick ? 20000000ll : 400000000ll); j++) {                  // [0.5, 2): what log10 hands to log, every 2^24-th or so value
        const uint64_t b = 0x3fe0000000000000ull + (uint64_t)j * 22517998ull + (uint64_t)(j * 2654435761ull & 0xffffff);
        double x; std::memcpy(&x, &b, 8);
        const double want = log10(x), got = phmm::glibc_log10(x);
        n64++;
        if (std::memcmp(&want, &got, 8)) mism64++;
    }
    std::printf("{\"double_checked\": %llu, \"double_mismatches\": %llu}\n", n64, mism64);
    return (mism || mism64) ? 1 : 0;
}
