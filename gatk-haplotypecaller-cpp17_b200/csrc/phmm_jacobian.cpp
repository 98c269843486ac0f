// phmm_jacobian.cpp -- the Jacobian-logarithm table of hc::MathUtils::approximate_log10_sum_log10
// (utils/math_utils.hpp:11-30), which the device-side genotype reduction (phmm_genotype.cu) indexes.
//
// The reference fills its table with `cache[k] = std::log10(1.0 + std::pow(10.0, -TABLE_STEP * k))` inside a
// lambda that initialises a static (math_utils.hpp:24-28).  GCC evaluates that lambda AT COMPILE TIME (math
// builtins are constant expressions for it and are folded with MPFR), so the table inside the reference's binary
// holds CORRECTLY ROUNDED values: 22 930 of its 80 001 entries differ in the last bit from what glibc 2.39's
// log10 returns at run time.  To hold the same numbers this file does the same thing on purpose -- a constexpr
// table the compiler must fold (it does not build otherwise) -- and tests/test_genotype.py pins the result
// against the reference's compiled-in table (hash in tests/golden/ref_gl.json), against the oracle's independent
// binary128 computation and against mpmath.
#include "phmm_tables.h"

#include <array>
#include <cstddef>

namespace phmm {
namespace {

constexpr double kMaxTolerance = 8.0;                      // math_utils.hpp:20
constexpr double kTableStep = 0.0001;                      // math_utils.hpp:24
constexpr std::size_t kSize = static_cast<std::size_t>(kMaxTolerance / kTableStep) + 1;   // math_utils.hpp:27

constexpr std::array<double, kSize> build_table()
{
    std::array<double, kSize> cache{};
    for (std::size_t k = 0; k < kSize; k++) cache[k] = __builtin_log10(1.0 + __builtin_pow(10.0, -kTableStep * k));
    return cache;
}
constexpr std::array<double, kSize> kTable = build_table();
static_assert(kSize == 80001 && kTable[0] > 0.3010299 && kTable[0] < 0.3010300 && kTable[kSize - 1] > 0.0,
              "the Jacobian table must be evaluated by the compiler");

}  // namespace

const double* jacobian_table(int* n)
{
    if (n) *n = (int)kSize;
    return kTable.data();
}

}  // namespace phmm
