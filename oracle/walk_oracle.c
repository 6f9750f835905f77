/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's random-walk transition rule
 * over a CSR graph, for parity checks at sizes where the pure-Python oracle is too slow.
 * Never linked into or called from the product library.
 *
 * Follows (reference file:line, same as oracle/walk_oracle.py, against which it is tested):
 *   DeepWalk.walk                    shallow_encoders/graph/random_walk_generator.py:61-72
 *   Node2Vec.walk (code rule)        shallow_encoders/graph/random_walk_generator.py:94-119
 *   edge weights                     shallow_encoders/graph/random_walk_generator.py:44-53
 *   random.choices inverse CDF       CPython 3.12.3 Lib/random.py Random.choices
 *   builtin sum() int/float mix      CPython 3.12.3 Python/bltinmodule.c builtin_sum_impl
 *
 * Parity status: pinned through tests/golden/walks_*.npz (generated from the real reference).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/libwalk_oracle.so oracle/walk_oracle.c -lm
 * (-ffp-contract=off: no FMA fusion, the arithmetic must round exactly like CPython's doubles.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define RULE_REFERENCE 0
#define RULE_PAPER 1

static int member(const int32_t *col, const int32_t *col_sorted, const int64_t *rowptr, int32_t x, int32_t t) {
    int64_t lo = rowptr[x], hi = rowptr[x + 1];
    if (col_sorted) {
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo) >> 1);
            int32_t c = col_sorted[mid];
            if (c < t) lo = mid + 1; else hi = mid;
        }
        return lo < rowptr[x + 1] && col_sorted[lo] == t;
    }
    for (int64_t i = lo; i < hi; ++i)
        if (col[i] == t) return 1;
    return 0;
}

/* is_float[i] != 0 when the python object at position i is a float (else an int). */
static double py312_sum(const double *v, const unsigned char *is_float, int64_t n, int *all_int) {
    int64_t i = 0;
    double i_result = 0.0; /* exact for |sum| < 2^53 */
    while (i < n && !is_float[i]) { i_result += v[i]; ++i; }
    if (i == n) { *all_int = 1; return i_result; }
    *all_int = 0;
    double f = i_result + v[i];
    ++i;
    double c = 0.0;
    for (; i < n; ++i) {
        double x = v[i];
        if (is_float[i]) {
            double t = f + x;
            if (fabs(f) >= fabs(x)) c += (f - t) + x; else c += (x - t) + f;
            f = t;
        } else {
            f += x;
        }
    }
    if (c != 0.0 && isfinite(c)) f += c;
    return f;
}

int oracle_walks(const int64_t *rowptr, const int32_t *col, const double *w, int w_is_int,
                 const int32_t *col_sorted, int64_t n_nodes,
                 const int32_t *starts, int64_t n_walks, int walk_len,
                 double p, double q, int node2vec, int rule,
                 const double *uniforms, int32_t *out) {
    int64_t max_deg = 0;
    for (int64_t v = 0; v < n_nodes; ++v) {
        int64_t d = rowptr[v + 1] - rowptr[v];
        if (d > max_deg) max_deg = d;
    }
    double *wt = (double *)malloc(sizeof(double) * (size_t)(max_deg > 0 ? max_deg : 1));
    unsigned char *isf = (unsigned char *)malloc((size_t)(max_deg > 0 ? max_deg : 1));
    if (!wt || !isf) { free(wt); free(isf); return -1; }
    const double inv_p = 1.0 / p, inv_q = 1.0 / q;
    int status = 0;

    for (int64_t wk = 0; wk < n_walks && status == 0; ++wk) {
        int32_t prev = -1, node = starts[wk];
        int32_t *o = out + wk * (int64_t)walk_len;
        if (walk_len > 0) o[0] = node;
        for (int s = 1; s < walk_len; ++s) {
            int64_t base = rowptr[node];
            int64_t deg = rowptr[node + 1] - base;
            if (deg <= 0) { status = -2; break; } /* reference: random.choices on empty population raises */
            for (int64_t i = 0; i < deg; ++i) {
                int32_t x = col[base + i];
                double wi = w ? w[base + i] : 1.0;
                unsigned char fl = (unsigned char)(w && !w_is_int);
                if (node2vec && prev >= 0) {
                    if (x == prev) {
                        wi *= inv_p; fl = 1;
                    } else {
                        int m = member(col, col_sorted, rowptr, x, prev);
                        if ((rule == RULE_REFERENCE && m) || (rule == RULE_PAPER && !m)) { wi *= inv_q; fl = 1; }
                    }
                }
                wt[i] = wi; isf[i] = fl;
            }
            int all_int;
            double ssum = py312_sum(wt, isf, deg, &all_int);
            /* cum = accumulate(w_i / s); total = cum[-1]; idx = bisect_right(cum, u*total, 0, deg-1) */
            double cum = 0.0;
            for (int64_t i = 0; i < deg; ++i) { wt[i] = (i == 0) ? (wt[0] / ssum) : (cum + wt[i] / ssum); cum = wt[i]; }
            double target = uniforms[wk * (int64_t)(walk_len - 1) + (s - 1)] * (cum + 0.0);
            int64_t lo = 0, hi = deg - 1;
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                if (target < wt[mid]) hi = mid; else lo = mid + 1;
            }
            int32_t child = col[base + lo];
            o[s] = child;
            prev = node; node = child;
        }
    }
    free(wt); free(isf);
    return status;
}
