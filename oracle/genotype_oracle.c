/*
 * genotype_oracle.c -- CPU restatement of the step right after the PairHMM path:
 * per-read allele marginalisation and diploid genotype likelihoods of one variant site, as
 * hc::Genetyper computes them from the capped / filtered likelihood matrix.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's checker legs); the product never links this.
 * Pinned against the reference's own code: oracle/ref_harness.cpp exposes Genetyper's marginal_likelihoods +
 * calculate_genotype_likelihoods (compiled from /root/reference), tests/golden/make_gl_golden.py writes
 * tests/golden/ref_gl.json from it, tests/test_genotype.py compares bit for bit.
 *
 * Restated (all citations relative to /root/reference/src/haplotypecaller/):
 *   genotyper/genotyper.hpp:245-264   marginal_likelihoods: max over the haplotypes carrying an allele,
 *                                      start value numeric_limits<double>::lowest()
 *   genotyper/genotyper.hpp:276-309   per read and genotype (a1 <= a2): a1 == a2 -> L[a] + log10(2);
 *                                      else approximate_log10_sum_log10(L[a1], L[a2])
 *   genotyper/genotyper.hpp:311-320   sum over the reads IN ORDER from 0.0, minus n_reads * log10(2)
 *   utils/math_utils.hpp:11-30        approximate_log10_sum_log10 with the Jacobian table: step 1e-4 up to 8.0,
 *                                      index std::round(diff * (1.0 / 0.0001))
 * The Jacobian table is NOT what the running libm would give: the reference initialises it with a lambda GCC
 * evaluates AT COMPILE TIME (math_utils.hpp:24-28: builtins folded with MPFR), so every entry is the CORRECTLY
 * ROUNDED log10(1.0 + CR(10^x)) -- 29% of the entries differ by one ulp from glibc 2.39's runtime log10.  This
 * file gets the same values independently of GCC's constant folder, through binary128 (libquadmath: 113-bit
 * results rounded once more to 53 bits are correctly rounded for all practical purposes), and
 * tests/test_genotype.py pins them against the reference's compiled-in table and against mpmath.
 * Genotype order: (a1, a2) with a1 outer, a2 >= a1 inner (genotyper.hpp:22-33, :297-307).
 */
#include <float.h>
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>

#define JAC_MAX_TOLERANCE 8.0
#define JAC_STEP 0.0001

static double* g_jac = NULL;
static int g_jac_n = 0;

static void jac_init(void)
{
    if (g_jac) return;
    g_jac_n = (int)(size_t)(JAC_MAX_TOLERANCE / JAC_STEP) + 1;          /* math_utils.hpp:23 */
    g_jac = (double*)malloc(sizeof(double) * (size_t)g_jac_n);
    for (size_t k = 0; k < (size_t)g_jac_n; k++) {                                                  /* :25-26 */
        const double x = -JAC_STEP * k;
        const double p = (double)powq(10.0Q, (__float128)x);
        const double s = 1.0 + p;
        g_jac[k] = (double)log10q((__float128)s);
    }
}

const double* oracle_jacobian_table(int* n) { jac_init(); if (n) *n = g_jac_n; return g_jac; }

static double approx_log10_sum_log10(double a, double b)                /* math_utils.hpp:11-16 */
{
    if (a > b) { double t = a; a = b; b = t; }
    const double diff = b - a;
    const double inv_step = 1.0 / JAC_STEP;
    return b + (diff < JAC_MAX_TOLERANCE ? g_jac[(size_t)round(diff * inv_step)] : 0.0);
}

/*
 * One site.  lik: the region's matrix [n_reads][n_haps] AFTER the cap (rows of erased reads are skipped through
 * keep[]).  hap_allele[h]: allele of haplotype h.  use[r]: read r overlaps the site's interval
 * (get_read_indices_to_keep, genotyper.hpp:235-244).  out: n_alleles (n_alleles + 1) / 2 doubles.
 * Returns the number of reads that entered the sums.
 */
int oracle_genotype_likelihoods(const double* lik, int n_reads, int n_haps, const uint8_t* keep, const uint8_t* use,
                                int n_alleles, const uint8_t* hap_allele, double* out)
{
    jac_init();
    const double log10_2 = log10(2.0);
    const int n_gt = n_alleles * (n_alleles + 1) / 2;
    double* al = (double*)malloc(sizeof(double) * (size_t)(n_reads > 0 ? n_reads : 1) * (size_t)n_alleles);
    int n_used = 0;
    for (int r = 0; r < n_reads; r++) {
        if ((keep && !keep[r]) || (use && !use[r])) continue;
        double* row = al + (size_t)n_used * n_alleles;
        for (int a = 0; a < n_alleles; a++) row[a] = -DBL_MAX;             /* lowest() */
        for (int h = 0; h < n_haps; h++) {
            const double v = lik[(size_t)r * n_haps + h];
            if (v > row[hap_allele[h]]) row[hap_allele[h]] = v;
        }
        n_used++;
    }
    int g = 0;
    for (int a1 = 0; a1 < n_alleles; a1++)
        for (int a2 = a1; a2 < n_alleles; a2++, g++) {
            double sum = 0.0;
            for (int r = 0; r < n_used; r++) {
                const double* row = al + (size_t)r * n_alleles;
                sum += (a1 == a2) ? row[a1] + log10_2 : approx_log10_sum_log10(row[a1], row[a2]);
            }
            out[g] = sum - (double)(size_t)n_used * log10_2;
        }
    (void)n_gt;
    free(al);
    return n_used;
}
