// phmm_tables.cpp -- host-built probability tables for the device kernels.
//
// Product code (not the oracle): restates native/Context.h of the reference so the kernels see the
// very same constants the AVX kernels read.
//   ph2pr[x]   = 10^(-x/10), x = 0..127, generated SEPARATELY per precision: powf for float
//                (Context.h:145-147), pow for double (Context.h:105-107).  92 of the 128 float
//                entries differ bitwise from (float)double, so they are never derived from each
//                other, and never computed on the device (device powf != glibc powf).
//   mm[i,d]    = matchToMatchProb (Context.h:50-61): 1 - (10^-i/10 + 10^-d/10) evaluated through
//                the reference's table-quantised approximateLog10SumLog10 (Context.h:67-90) IN THE
//                TABLE'S OWN PRECISION, then log1p/pow in double, then narrowed.  Triangular layout
//                ((max*(max+1))>>1)+min (Context.h:123-134).  Only qualities <= 127 are reachable
//                because every byte is masked with & 127 (avx-pairhmm-template.h:110-112), so the
//                first 128*129/2 = 8256 entries are kept.
// Must be compiled without FMA contraction and without fast-math (see Makefile).
#include "phmm_tables.h"
#include "phmm_log10.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace phmm {
namespace {

constexpr int    kMaxQual = 127;
constexpr double kJacobianTolerance = 8.0;
constexpr double kJacobianStep = 0.0001;
constexpr double kJacobianInvStep = 1.0 / kJacobianStep;
constexpr int    kJacobianSize = (int)(kJacobianTolerance / kJacobianStep) + 1;

template <class T> int fast_round(T d) { return (d > (T)0.0) ? (int)(d + (T)0.5) : (int)(d - (T)0.5); }

template <class T>
T approx_log10_sum(const std::vector<T>& jac, T small, T big)
{
    if (small > big) { T t = big; big = small; small = t; }
    // (the reference's `isinf(x) == -1` tests are always false in C++11: omitted)
    T diff = big - small;
    if (diff >= (T)kJacobianTolerance) return big;
    int ind = fast_round<T>((T)(diff * ((T)kJacobianInvStep)));
    return big + jac[ind];
}

template <class T>
void build_mm(std::vector<T>& mm)
{
    std::vector<T> jac(kJacobianSize);
    for (int k = 0; k < kJacobianSize; k++)
        jac[k] = (T)(std::log10(1.0 + std::pow(10.0, -((double)k) * kJacobianStep)));
    mm.resize(kMmEntries);
    const double LN10 = std::log(10.0);
    const double INV_LN10 = 1.0 / LN10;
    for (int i = 0, offset = 0; i <= kMaxQual; offset += ++i)
        for (int j = 0; j <= i; j++) {
            double log10_sum = approx_log10_sum<T>(jac, (T)(-0.1 * i), (T)(-0.1 * j));
            double mm_log10 = std::log1p(-std::fmin(1.0, std::pow(10.0, log10_sum))) * INV_LN10;
            mm[offset + j] = (T)(std::pow(10.0, mm_log10));
        }
}

Tables* g_tables = nullptr;
std::once_flag g_once;

}  // namespace

const Tables& host_tables()
{
    std::call_once(g_once, [] {
        Tables* t = new Tables();
        build_mm<float>(t->mm_f);
        build_mm<double>(t->mm_d);
        t->ph2pr_f.resize(128);
        t->ph2pr_d.resize(128);
        for (int x = 0; x < 128; x++) {
            t->ph2pr_d[x] = std::pow(10.0, -((double)x) / 10.0);
            t->ph2pr_f[x] = powf(10.f, -((float)x) / 10.f);
        }
        t->log10_init_f = log10f(ldexpf(1.f, 120));     // Context.h:149-150
        t->log10_init_d = std::log10(std::ldexp(1.0, 1020));   // Context.h:109-110
        g_tables = t;
    });
    return *g_tables;
}

bool log10_restatement_matches_libm()
{
    static const bool ok = [] {
        uint32_t st = 0x2545f491u;
        for (int i = 0; i < 100000; i++) {                  // floats: every exponent, random mantissas (x >= 0)
            st ^= st << 13; st ^= st >> 17; st ^= st << 5;
            const float x = l10_float(st % 0x7f800000u);
            const float a = ::log10f(x), b = glibc_log10f(x);
            if (l10_bits(a) != l10_bits(b)) {
                if (getenv("PHMM_TRACE_INIT")) fprintf(stderr, "phmm init trace: log10f(%a) libm %a restated %a\n", x, a, b);
                return false;
            }
        }
        uint64_t s64 = 0x9e3779b97f4a7c15ull;
        for (int i = 0; i < 50000; i++) {                   // doubles
            s64 ^= s64 << 13; s64 ^= s64 >> 7; s64 ^= s64 << 17;
            const double x = l10_double(s64 % 0x7ff0000000000000ull);
            const double a = ::log10(x), b = glibc_log10(x);
            if (l10_bits64(a) != l10_bits64(b)) {
                if (getenv("PHMM_TRACE_INIT")) fprintf(stderr, "phmm init trace: log10(%a) libm %a restated %a\n", x, a, b);
                return false;
            }
        }
        return true;
    }();
    return ok;
}

}  // namespace phmm
