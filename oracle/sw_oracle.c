/* sw_oracle.c -- plain-C restatement of the reference's Smith-Waterman haplotype -> reference aligner.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu leg).  Follows
 *   smithwaterman/intel_smithwaterman.hpp:29-59   align(): the all-match shortcut (<= 2 mismatches at equal
 *                                                 length -> offset 0, "<len>M"), else runSWOnePairBT_avx2 with
 *                                                 overhang strategy 9 == SOFTCLIP
 *   smithwaterman/native/PairWiseSW.h:123-159     MAIN_CODE: the cell update and its back-track bits
 *                 :161-365                        smithWatermanBackTrack: borders, anti-diagonal order of the
 *                                                 best-end-cell search and its tie rules
 *                 :367-520                        getCIGAR: back-track state machine, soft clips, merge, print
 * The reference sweeps anti-diagonals with AVX2; a cell depends only on (i-1,j-1), (i-1,j), (i,j-1), so the
 * row-major loop below computes the same integers.  Pinned against the compiled reference (tests/test_sw.py).
 *
 * seq1 = reference window (rows i = 1..nrow), seq2 = haplotype (columns j = 1..ncol).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SW_MATCH 0
#define SW_INSERT 1
#define SW_DELETE 2
#define SW_INSERT_EXT 4
#define SW_DELETE_EXT 8
#define SW_SOFTCLIP 9
#define SW_MIN_CUTOFF (-100000000)           /* MATRIX_MIN_CUTOFF, smithwaterman_common.h */
#define SW_LOW_INIT (INT32_MIN / 2)          /* LOW_INIT_VALUE */

/* intel_smithwaterman.hpp:47-58 */
static int all_match(const uint8_t* ref, int nref, const uint8_t* alt, int nalt)
{
    if (nref != nalt) return 0;
    int mismatch = 0;
    for (int i = 0; mismatch <= 2 && i < nref; i++) if (alt[i] != ref[i]) mismatch++;
    return mismatch <= 2;
}

/* Returns the alignment offset; writes the CIGAR string (NUL-terminated) into cigar (capacity cap). */
int sw_oracle_kernel(const uint8_t* seq1, int nrow, const uint8_t* seq2, int ncol,
                     int w_match, int w_mismatch, int w_open, int w_extend, char* cigar, int cap)
{
    const size_t W = (size_t)ncol + 1;
    int32_t* H = (int32_t*)malloc(sizeof(int32_t) * 2 * W);      /* two rows */
    int32_t* F = (int32_t*)malloc(sizeof(int32_t) * W);          /* vertical gap state per column */
    uint8_t* bt = (uint8_t*)malloc((size_t)(nrow + 1) * W);
    int32_t* lastcol = (int32_t*)malloc(sizeof(int32_t) * (nrow + 1));
    int32_t* Hp = H, *Hc = H + W;
    for (int j = 0; j <= ncol; j++) { Hp[j] = 0; F[j] = SW_LOW_INIT; }            /* :204-224, borders :318-327 */
    for (int i = 1; i <= nrow; i++) {
        int32_t E = SW_LOW_INIT;                                                  /* E(i, 0) */
        Hc[0] = 0;
        for (int j = 1; j <= ncol; j++) {
            /* MAIN_CODE (:123-159) */
            const int32_t ext_h = E + w_extend, open_h = Hc[j - 1] + w_open;
            const int32_t e11 = open_h > ext_h ? open_h : ext_h;
            int ext = (open_h > ext_h) ? 0 : SW_INSERT_EXT;
            const int32_t ext_v = F[j] + w_extend, open_v = Hp[j] + w_open;
            const int32_t f11 = ext_v > open_v ? ext_v : open_v;
            if (!(open_v > ext_v)) ext |= SW_DELETE_EXT;
            const int32_t m11 = Hp[j - 1] + (seq1[i - 1] == seq2[j - 1] ? w_match : w_mismatch);
            int32_t h11 = m11 > SW_MIN_CUTOFF ? m11 : SW_MIN_CUTOFF;
            int b = SW_MATCH;
            if (e11 > h11) { b = SW_INSERT; h11 = e11; }
            if (f11 > h11) { b = SW_DELETE; h11 = f11; }
            E = e11; F[j] = f11; Hc[j] = h11;
            bt[(size_t)i * W + j] = (uint8_t)(b | ext);
        }
        lastcol[i] = Hc[ncol];
        int32_t* t = Hp; Hp = Hc; Hc = t;
    }
    /* Hp now holds row nrow.  Best end cell, in the reference's anti-diagonal order (:329-357, SOFTCLIP):
       on every anti-diagonal first the last-ROW cell, then the last-COLUMN cell. */
    int32_t maxScore = INT32_MIN; int max_i = 0, max_j = 0;
    for (int ad = 1; ad <= nrow + ncol; ad++) {
        if (ad >= nrow + 1) {                                     /* ilo == nrow + 1: cell (nrow, ad - nrow) */
            const int j = ad - nrow; const int32_t score = Hp[j];
            if (maxScore < score || (maxScore == score && abs(nrow - j) < abs(max_i - max_j))) { maxScore = score; max_i = nrow; max_j = j; }
        }
        if (ad >= ncol + 1) {                                     /* jhi == ncol + 1: cell (ad - ncol, ncol) */
            const int i = ad - ncol; const int32_t score = lastcol[i];
            if (maxScore < score || (maxScore == score && (max_j == ncol || abs(i - ncol) <= abs(max_i - max_j)))) { maxScore = score; max_i = i; max_j = ncol; }
        }
    }
    /* getCIGAR (:367-520), overhang SOFTCLIP */
    int n = 0; int cap_el = nrow + ncol + 4;
    int* op = (int*)malloc(sizeof(int) * cap_el); int* len = (int*)malloc(sizeof(int) * cap_el);
    int i = max_i, j = max_j;
    if (j < ncol) { op[n] = SW_SOFTCLIP; len[n] = ncol - j; n++; }
    int state = 0;
    while (i > 0 && j > 0) {
        const int btr = bt[(size_t)i * W + j];
        if (state == SW_INSERT_EXT) { j--; len[n - 1]++; state = btr & SW_INSERT_EXT; }
        else if (state == SW_DELETE_EXT) { i--; len[n - 1]++; state = btr & SW_DELETE_EXT; }
        else switch (btr & 3) {
            case SW_MATCH:  i--; j--; op[n] = SW_MATCH;  len[n] = 1; state = 0; n++; break;
            case SW_INSERT: j--;      op[n] = SW_INSERT; len[n] = 1; state = btr & SW_INSERT_EXT; n++; break;
            case SW_DELETE: i--;      op[n] = SW_DELETE; len[n] = 1; state = btr & SW_DELETE_EXT; n++; break;
        }
    }
    if (j > 0) { op[n] = SW_SOFTCLIP; len[n] = j; n++; }
    const int offset = i;
    int m = 0;                                                    /* merge equal neighbours (:478-495) */
    for (int k = 1; k < n; k++) {
        if (op[k] == op[m]) len[m] += len[k];
        else { m++; op[m] = op[k]; len[m] = len[k]; }
    }
    int w = 0; cigar[0] = 0;
    for (int k = (n ? m : -1); k >= 0; k--) {                     /* printed back to front (:497-518) */
        const char c = op[k] == SW_MATCH ? 'M' : op[k] == SW_INSERT ? 'I' : op[k] == SW_DELETE ? 'D' : op[k] == SW_SOFTCLIP ? 'S' : 'R';
        w += snprintf(cigar + w, (size_t)(cap - w), "%d%c", len[k], c);
        if (w >= cap - 1) break;
    }
    free(H); free(F); free(bt); free(lastcol); free(op); free(len);
    return offset;
}

/* hc::IntelSWAligner::align (intel_smithwaterman.hpp:29-44) */
int sw_oracle_align(const uint8_t* ref, int nref, const uint8_t* alt, int nalt,
                    int w_match, int w_mismatch, int w_open, int w_extend, char* cigar, int cap)
{
    if (all_match(ref, nref, alt, nalt)) { snprintf(cigar, (size_t)cap, "%dM", nref); return 0; }
    return sw_oracle_kernel(ref, nref, alt, nalt, w_match, w_mismatch, w_open, w_extend, cigar, cap);
}
