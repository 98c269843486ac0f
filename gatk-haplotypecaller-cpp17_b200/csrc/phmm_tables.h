// phmm_tables.h -- host-built probability tables (see phmm_tables.cpp).
#pragma once
#include <vector>

namespace phmm {

constexpr int kMmEntries = (128 * 129) / 2;   // triangular over qualities 0..127

struct Tables {
    std::vector<float>  ph2pr_f, mm_f;
    std::vector<double> ph2pr_d, mm_d;
    float  log10_init_f;    // log10f(2^120)
    double log10_init_d;    // log10(2^1020)
};

const Tables& host_tables();

// hc::MathUtils' Jacobian-logarithm table (utils/math_utils.hpp:17-29), 80 001 correctly rounded doubles
// (phmm_jacobian.cpp).  inv_step = 1.0 / 0.0001, max tolerance 8.0.
const double* jacobian_table(int* n);

// True when the restatement of glibc's log10f / log10 (phmm_log10.h) agrees bit for bit with the libm this
// process runs on, over a few hundred thousand sampled inputs: only then may the device take the final log10
// (a host without FMA selects another variant of logf).  Evaluated once.
bool log10_restatement_matches_libm();

}  // namespace phmm
