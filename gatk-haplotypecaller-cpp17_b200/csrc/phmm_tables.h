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

}  // namespace phmm
