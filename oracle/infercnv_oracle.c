/*
 * infercnv_oracle.c - CPU restatement (float64, long-double accumulators where R uses them) of the
 * inferCNV smoothing + HMM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may call it.  The product path
 * (infercnv_b200/) never links, imports or falls back to it.
 *
 * Every function cites the reference R code it restates (paths relative to the reference
 * checkout, commit 65e6bf5, v1.23.0).  R is not installable in the build image, so this is a
 * restatement, not the reference itself; it is pinned against the reference's own known answers
 * (tests/testthat/test_infer_cnv.R:89-172, :307-360) and the bundled golden object
 * data/infercnv_object_example.rda (count.data -> expr.data) by tests/test_oracle_*.py.
 * The Viterbi / median-filter parts have no fixture in the reference: "parity unpinned" there.
 *
 * Third-party arithmetic restated here because the reference reaches it through base R
 * (R >= 4.0, `stats`/`base`; source not in the reference checkout):
 *   - mean():      long-double sum / n plus one long-double refinement pass (summary.c real_mean)
 *   - sum():       long-double accumulation
 *   - median():    partial sort; even n -> mean() of the two middle values
 *   - sd()/var():  two-pass, long-double, refined mean (cov.c)
 *   - colSums()/rowMeans(): long-double accumulation
 *   - stats::filter(sides=2): double accumulation, taps applied right-to-left (filter.c cfilter)
 *   - pnorm(log.p=TRUE, lower.tail=FALSE): nmath pnorm_both (Cody 1969 rational approximations)
 *
 * Layout everywhere: R column-major, X[g + G*c], g = gene (row), c = cell (column).
 * All index arguments are 0-based.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

typedef long double ldouble;

/* threads used by the element-wise / reduction helpers (drivers with an nthreads argument set it) */
static int g_threads = 1;
ORC_API void orc_set_num_threads(int n) { g_threads = n > 0 ? n : 1; }

/* ------------------------------------------------------------------------------------------- */
/* base-R numerics                                                                             */
/* ------------------------------------------------------------------------------------------- */

/* R mean() for doubles: summary.c real_mean(): LD sum, /n, then one refinement pass. */
static double r_mean(const double *x, int64_t n) {
    ldouble s = 0.0L;
    for (int64_t i = 0; i < n; ++i) s += x[i];
    s /= (ldouble)n;
    ldouble t = 0.0L;
    for (int64_t i = 0; i < n; ++i) t += (x[i] - s);
    s += t / (ldouble)n;
    return (double)s;
}

static double r_mean_strided(const double *x, int64_t n, int64_t stride, const int32_t *idx) {
    /* mean(x[idx]) where element i lives at x[idx[i]*stride] */
    ldouble s = 0.0L;
    for (int64_t i = 0; i < n; ++i) s += x[(int64_t)idx[i] * stride];
    s /= (ldouble)n;
    ldouble t = 0.0L;
    for (int64_t i = 0; i < n; ++i) t += (x[(int64_t)idx[i] * stride] - s);
    s += t / (ldouble)n;
    return (double)s;
}

static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* R median.default(): n odd -> middle of sort; n even -> mean(c(a, b)). Destroys `buf`. */
static double r_median_inplace(double *buf, int64_t n) {
    qsort(buf, (size_t)n, sizeof(double), cmp_double);
    if (n & 1) return buf[n / 2];
    double two[2] = {buf[n / 2 - 1], buf[n / 2]};
    return r_mean(two, 2);
}

/* R var()/sd(): cov.c two-pass with refined LD mean, divisor n-1. */
static double r_sd(const double *x, int64_t n) {
    ldouble s = 0.0L;
    for (int64_t i = 0; i < n; ++i) s += x[i];
    ldouble m = s / (ldouble)n;
    ldouble t = 0.0L;
    for (int64_t i = 0; i < n; ++i) t += (x[i] - m);
    m += t / (ldouble)n;
    ldouble ss = 0.0L;
    for (int64_t i = 0; i < n; ++i) {
        ldouble d = x[i] - m;
        ss += d * d;
    }
    return sqrt((double)(ss / (ldouble)(n - 1)));
}

/*
 * nmath pnorm_both(), upper tail, log.p = TRUE, for x >= 0 or any finite x.
 * Published algorithm: W. J. Cody (1969) "Rational Chebyshev approximations for the error
 * function", Math. Comp. 23, as used by R's nmath/pnorm.c (three ranges: |x| <= 0.67448975,
 * <= sqrt(32), and the asymptotic tail).  Called by the reference at R/inferCNV_HMM.R:1129,1156.
 */
ORC_API double orc_pnorm_upper_log(double x) {
    static const double a[5] = {2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582,
                                18154.981253343561249, 0.065682337918207449113};
    static const double b[4] = {47.20258190468824187, 976.09855173777669322, 10260.932208618978205,
                                45507.789335026729956};
    static const double c[9] = {0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979,
                                597.27027639480026226,  2494.5375852903726711, 6848.1904505362823326,
                                11602.651437647350124,  9842.7148383839780218, 1.0765576773720192317e-8};
    static const double d[8] = {22.266688044328115691, 235.38790178262499861, 1519.377599407554805,
                                6485.558298266760755,  18615.571640885098091, 34900.952721145977266,
                                38912.003286093271411, 19685.429676859990727};
    static const double p[6] = {0.21589853405795699,      0.1274011611602473639, 0.022235277870649807,
                                0.001421619193227893466,  2.9112874951168792e-5, 0.02307344176494017303};
    static const double q[5] = {1.28426009614491121, 0.468238212480865118, 0.0659881378689285515,
                                0.00378239633202758244, 7.29751555083966205e-5};
    const double M_SQRT_32 = 5.656854249492380195206754896838;
    const double M_1_SQRT_2PI = 0.398942280401432677939946059934;
    double xden, xnum, temp, del, xsq, y, cum, ccum;
    const double eps = DBL_EPSILON * 0.5;

    if (isnan(x)) return x;
    y = fabs(x);
    if (y <= 0.67448975) {
        if (y > eps) {
            xsq = x * x;
            xnum = a[4] * xsq;
            xden = xsq;
            for (int i = 0; i < 3; ++i) {
                xnum = (xnum + a[i]) * xsq;
                xden = (xden + b[i]) * xsq;
            }
        } else {
            xnum = xden = 0.0;
        }
        temp = x * (xnum + a[3]) / (xden + b[3]);
        ccum = 0.5 - temp;
        return log(ccum);
    }
    if (y <= M_SQRT_32) {
        xnum = c[8] * y;
        xden = y;
        for (int i = 0; i < 7; ++i) {
            xnum = (xnum + c[i]) * y;
            xden = (xden + d[i]) * y;
        }
        temp = (xnum + c[7]) / (xden + d[7]);
        xsq = trunc(y * 16) / 16;
        del = (y - xsq) * (y + xsq);
        cum = (-xsq * xsq * 0.5) + (-del * 0.5) + log(temp);
        if (x > 0.) return cum; /* swap_tail: upper tail of positive x is the small side */
        ccum = log1p(-exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp);
        return ccum;
    }
    if (y < 1e170) {
        xsq = 1.0 / (x * x);
        xnum = p[5] * xsq;
        xden = xsq;
        for (int i = 0; i < 4; ++i) {
            xnum = (xnum + p[i]) * xsq;
            xden = (xden + q[i]) * xsq;
        }
        temp = xsq * (xnum + p[4]) / (xden + q[4]);
        temp = (M_1_SQRT_2PI - temp) / y;
        xsq = trunc(x * 16) / 16;
        del = (x - xsq) * (x + xsq);
        cum = (-xsq * xsq * 0.5) + (-del * 0.5) + log(temp);
        if (x > 0.) return cum;
        ccum = log1p(-exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp);
        return ccum;
    }
    return (x > 0) ? -INFINITY : 0.0;
}

/* ------------------------------------------------------------------------------------------- */
/* element-wise steps                                                                          */
/* ------------------------------------------------------------------------------------------- */

/* R/inferCNV_ops.R:3082-3111 .normalize_data_matrix_by_seq_depth: x / colSums * median(colSums)
 * (normalize_factor < 0 -> use the median of the column sums, as the reference does for NA). */
ORC_API int orc_normalize_by_seq_depth(const double *X, double *Y, int64_t G, int64_t C, double normalize_factor) {
    double *cs = (double *)malloc(sizeof(double) * (size_t)C);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)C);
    if (!cs || !tmp) return -1;
    for (int64_t c = 0; c < C; ++c) {
        ldouble s = 0.0L;
        for (int64_t g = 0; g < G; ++g) s += X[g + G * c];
        cs[c] = (double)s;
    }
    if (!(normalize_factor >= 0)) {
        memcpy(tmp, cs, sizeof(double) * (size_t)C);
        normalize_factor = r_median_inplace(tmp, C);
    }
    for (int64_t c = 0; c < C; ++c)
        for (int64_t g = 0; g < G; ++g) Y[g + G * c] = (X[g + G * c] / cs[c]) * normalize_factor;
    free(cs);
    free(tmp);
    return 0;
}

/* R/inferCNV_ops.R:2756-2769 log2xplus1 */
ORC_API void orc_log2xplus1(const double *X, double *Y, int64_t n) {
    #pragma omp parallel for schedule(static) num_threads(g_threads)
    for (int64_t i = 0; i < n; ++i) Y[i] = log2(X[i] + 1.0);
}

/* R/inferCNV_ops.R:2814-2826 invert_log2: 2^x */
ORC_API void orc_invert_log2(const double *X, double *Y, int64_t n) {
    #pragma omp parallel for schedule(static) num_threads(g_threads)
    for (int64_t i = 0; i < n; ++i) Y[i] = pow(2.0, X[i]);
}

/* R/inferCNV_ops.R:2970-2983 apply_max_threshold_bounds */
ORC_API void orc_apply_max_threshold_bounds(const double *X, double *Y, int64_t n, double threshold) {
    #pragma omp parallel for schedule(static) num_threads(g_threads)
    for (int64_t i = 0; i < n; ++i) {
        double v = X[i];
        if (v > threshold) v = threshold;
        if (v < -threshold) v = -threshold;
        Y[i] = v;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* reference subtraction                                                                       */
/* ------------------------------------------------------------------------------------------- */

/* R/inferCNV_ops.R:1708-1735 .get_normal_gene_mean_bounds: means[g + G*k] = mean(X[g, group k]);
 * inv_log: log2(mean(2^x - 1) + 1). */
ORC_API int orc_ref_means(const double *X, int64_t G, int64_t C, const int32_t *grp_off, const int32_t *grp_idx,
                          int n_grp, int inv_log, double *means) {
    (void)C;
    /* mean() per gene = LD sum in list order, / n, one LD refinement pass (summary.c).  The loops
     * run cell-outer so memory is walked contiguously; per gene the additions keep list order. */
    ldouble *s = (ldouble *)malloc(sizeof(ldouble) * (size_t)G);
    ldouble *t = (ldouble *)malloc(sizeof(ldouble) * (size_t)G);
    if (!s || !t) return -1;
    for (int k = 0; k < n_grp; ++k) {
        int64_t n = grp_off[k + 1] - grp_off[k];
        const int32_t *idx = grp_idx + grp_off[k];
#pragma omp parallel num_threads(g_threads)
        {
#pragma omp for schedule(static)
            for (int64_t g = 0; g < G; ++g) s[g] = 0.0L;
            for (int64_t i = 0; i < n; ++i) {
                const double *col = X + G * (int64_t)idx[i];
#pragma omp for schedule(static)
                for (int64_t g = 0; g < G; ++g) s[g] += inv_log ? (pow(2.0, col[g]) - 1.0) : col[g];
            }
#pragma omp for schedule(static)
            for (int64_t g = 0; g < G; ++g) {
                s[g] /= (ldouble)n;
                t[g] = 0.0L;
            }
            for (int64_t i = 0; i < n; ++i) {
                const double *col = X + G * (int64_t)idx[i];
#pragma omp for schedule(static)
                for (int64_t g = 0; g < G; ++g) t[g] += ((inv_log ? (pow(2.0, col[g]) - 1.0) : col[g]) - s[g]);
            }
#pragma omp for schedule(static)
            for (int64_t g = 0; g < G; ++g) {
                double m = (double)(s[g] + t[g] / (ldouble)n);
                means[g + G * k] = inv_log ? log2(m + 1.0) : m;
            }
        }
    }
    free(s);
    free(t);
    return 0;
}

/* R/inferCNV_ops.R:1742-1786 .subtract_expr */
ORC_API void orc_subtract_ref(const double *X, double *Y, int64_t G, int64_t C, const double *means, int n_grp,
                              int use_bounds) {
    double *lo = (double *)malloc(sizeof(double) * (size_t)G * 3);
    double *hi = lo + G, *mid = lo + 2 * G;
    double *gm = (double *)malloc(sizeof(double) * (size_t)(n_grp > 0 ? n_grp : 1));
    for (int64_t g = 0; g < G; ++g) {
        double l = means[g], h = means[g];
        for (int k = 0; k < n_grp; ++k) {
            double m = means[g + G * k];
            gm[k] = m;
            if (m < l) l = m;
            if (m > h) h = m;
        }
        lo[g] = l;
        hi[g] = h;
        mid[g] = r_mean(gm, n_grp);
    }
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (int64_t c = 0; c < C; ++c) {
        for (int64_t g = 0; g < G; ++g) {
            double x = X[g + G * c];
            double y;
            if (use_bounds) {
                if (x > hi[g]) y = x - hi[g];
                else if (x < lo[g]) y = x - lo[g];
                else y = 0.0;
            } else {
                y = x - mid[g];
            }
            Y[g + G * c] = y;
        }
    }
    free(lo);
    free(gm);
}

/* ------------------------------------------------------------------------------------------- */
/* smoothing                                                                                   */
/* ------------------------------------------------------------------------------------------- */

/* Literal transcription of .smooth_helper + .smooth_center_helper
 * (R/inferCNV_ops.R:2483-2532, :2640-2661) for one NA-free vector; window odd, >= 3. */
static void smooth_helper_literal(const double *obs, double *out, int64_t n, int window) {
    int h = (window - 1) / 2;
    memcpy(out, obs, sizeof(double) * (size_t)n);
    if (n >= window) {
        /* stats::filter(vals, custom_filter, sides=2): out[i] = sum_j f[j]*x[i+h-j], double acc */
        double denom = (double)h * (double)h + (double)window;
        double *f = (double *)malloc(sizeof(double) * (size_t)window);
        for (int j = 0; j < window; ++j) {
            int w = (j <= h) ? (j + 1) : (window - j);
            f[j] = (double)w / denom;
        }
        for (int64_t i = h; i + h < n; ++i) {
            double z = 0.0;
            for (int j = 0; j < window; ++j) z += f[j] * obs[i + h - j];
            out[i] = z;
        }
        free(f);
    }
    int64_t iter = (n > window) ? h : (n + 1) / 2;
    for (int64_t te = 1; te <= iter; ++te) { /* tail_end, 1-based as in R */
        int64_t end_tail = n - te + 1;
        int64_t d_left = te - 1;
        int64_t d_right = n - te;
        if (d_right > h) d_right = h;
        int64_t r_left = h - d_left;
        int64_t r_right = h - d_right;
        double denominator = (((double)(window - 1) / 2.0) * ((double)(window - 1) / 2.0) + (double)window) -
                             ((double)(r_left * (r_left + 1)) / 2.0) - ((double)(r_right * (r_right + 1)) / 2.0);
        /* numerator_range = numerator_counts_vector[(h+1-d_left):(h+1+d_right)] = weights h+1-|k| */
        int64_t len = te + d_right; /* left chunk obs[1:(te+d_right)] */
        ldouble sl = 0.0L, sr = 0.0L;
        for (int64_t q = 0; q < len; ++q) {
            int64_t pos = (h + 1 - d_left) + q;               /* 1-based index into the 2h+1 weights */
            double w = (double)((pos <= h + 1) ? pos : (2 * h + 2 - pos));
            sl += (ldouble)(obs[q] * w);                        /* R: sum(chunk * range): double product, LD sum */
            /* right chunk obs[(end_tail-d_right):n] times rev(range) */
            int64_t posr = (h + 1 - d_left) + (len - 1 - q);
            double wr = (double)((posr <= h + 1) ? posr : (2 * h + 2 - posr));
            sr += (ldouble)(obs[(end_tail - d_right - 1) + q] * wr);
        }
        out[te - 1] = (double)sl / denominator;
        out[end_tail - 1] = (double)sr / denominator;
    }
}

/* Unified closed form: truncated, renormalised triangle. */
static void smooth_unified(const double *x, double *out, int64_t n, int window) {
    int64_t h = (window - 1) / 2;
    for (int64_t i = 0; i < n; ++i) {
        int64_t lo = i - h < 0 ? 0 : i - h;
        int64_t hi = i + h > n - 1 ? n - 1 : i + h;
        ldouble s = 0.0L;
        for (int64_t j = lo; j <= hi; ++j) {
            int64_t dist = j > i ? j - i : i - j;
            s += (ldouble)x[j] * (ldouble)(h + 1 - dist);
        }
        int64_t rl = h - i > 0 ? h - i : 0;
        int64_t rr = h - (n - 1 - i) > 0 ? h - (n - 1 - i) : 0;
        double D = (double)((h + 1) * (h + 1)) - (double)(rl * (rl + 1)) / 2.0 - (double)(rr * (rr + 1)) / 2.0;
        out[i] = (double)(s / (ldouble)D);
    }
}

/* R/inferCNV_ops.R:2406-2466 smooth_by_chromosome + .smooth_window.
 * chr_start/chr_len: K contiguous row ranges.  literal != 0 -> transcription of the R loops,
 * else the closed form.  window < 2 -> copy (R/inferCNV_ops.R:2444-2447); chromosomes with a
 * single gene are skipped (:2417).  Even windows are rejected (-2): the reference's behaviour
 * there is accidental. */
ORC_API int orc_smooth_by_chromosome(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                                     const int32_t *chr_len, int K, int window, int literal, int nthreads) {
    if (X != Y) memcpy(Y, X, sizeof(double) * (size_t)(G * C));
    if (window < 2) return 0;
    if ((window & 1) == 0) return -2;
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t c = 0; c < C; ++c) {
        double *tmp = (double *)malloc(sizeof(double) * (size_t)G);
        for (int k = 0; k < K; ++k) {
            int64_t n = chr_len[k];
            if (n < 2) continue;
            const double *src = X + chr_start[k] + G * c;
            if (literal) smooth_helper_literal(src, tmp, n, window);
            else smooth_unified(src, tmp, n, window);
            memcpy(Y + chr_start[k] + G * c, tmp, sizeof(double) * (size_t)n);
        }
        free(tmp);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* centring                                                                                    */
/* ------------------------------------------------------------------------------------------- */

/* R/inferCNV_ops.R:2094-2109 .center_columns */
ORC_API int orc_center_columns(const double *X, double *Y, int64_t G, int64_t C, int use_median, int nthreads) {
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t c = 0; c < C; ++c) {
        double m;
        if (use_median) {
            double *tmp = (double *)malloc(sizeof(double) * (size_t)G);
            memcpy(tmp, X + G * c, sizeof(double) * (size_t)G);
            m = r_median_inplace(tmp, G);
            free(tmp);
        } else {
            m = r_mean(X + G * c, G);
        }
        for (int64_t g = 0; g < G; ++g) Y[g + G * c] = X[g + G * c] - m;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* median filter                                                                               */
/* ------------------------------------------------------------------------------------------- */

/* R/noise_reduction.R:43-113 apply_median_filtering/.median_filter.  Blocks = (chromosome) x
 * (one index list: a subcluster for observations, a whole group for references), in list order.
 * The reference's window radius is half_window+1 = (window_size+1)/2 (noise_reduction.R:102-106). */
ORC_API int orc_median_filter(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                              const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx,
                              int n_grp, int window_size, int nthreads) {
    if (window_size < 2 || (window_size & 1) == 0) return -2;
    if (X != Y) memcpy(Y, X, sizeof(double) * (size_t)(G * C));
    int64_t r = (window_size + 1) / 2;
    (void)nthreads;
    for (int b = 0; b < n_grp; ++b) {
        int64_t m = grp_off[b + 1] - grp_off[b];
        const int32_t *idx = grp_idx + grp_off[b];
        for (int k = 0; k < K; ++k) {
            int64_t n = chr_len[k], s = chr_start[k];
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
            for (int64_t j = 0; j < m; ++j) {
                double buf[1024];
                int64_t ya = j - r < 0 ? 0 : j - r;
                int64_t yb = j + r > m - 1 ? m - 1 : j + r;
                /* the ifelse() at noise_reduction.R:105: posy >= ydim-(half_window+1) -> ydim; for
                 * posy == ydim-r exactly posy+r == ydim as well, so it equals the clamp. */
                for (int64_t i = 0; i < n; ++i) {
                    int64_t xa = i - r < 0 ? 0 : i - r;
                    int64_t xb = i + r > n - 1 ? n - 1 : i + r;
                    int cnt = 0;
                    for (int64_t jj = ya; jj <= yb; ++jj)
                        for (int64_t ii = xa; ii <= xb; ++ii) buf[cnt++] = X[s + ii + G * (int64_t)idx[jj]];
                    Y[s + i + G * (int64_t)idx[j]] = r_median_inplace(buf, cnt);
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* Viterbi                                                                                     */
/* ------------------------------------------------------------------------------------------- */

/* R/inferCNV_HMM.R:1101-1176 Viterbi.dthmm.adj for one sequence.
 * Pi is m x m column-major as R stores it (Pi[j + m*k] = P(j -> k)).  states out: 1..m.
 * Returns 0, or -3 for "Problems With Underflow" (HMM.R:1165), -4 for non-finite input.
 * *min_margin (optional) receives the smallest gap between the winning and the runner-up
 * candidate over all arg-max decisions of the traceback - a diagnostic, not part of R. */
ORC_API int orc_viterbi_seq(const double *x, int64_t n, int m, const double *Pi, const double *delta,
                            const double *mean, const double *sd, int32_t *states, double *min_margin) {
    if (min_margin) *min_margin = INFINITY;
    if (n < 2) { /* HMM.R:1104-1107 */
        for (int64_t i = 0; i < n; ++i) states[i] = 3;
        return 0;
    }
    if (m > 16) return -5;
    double sdv[16], logPi[256], logdelta[16];
    memcpy(sdv, sd, sizeof(double) * (size_t)m);
    double sdm = r_median_inplace(sdv, m); /* HMM.R:1122 */
    for (int i = 0; i < m * m; ++i) logPi[i] = log(Pi[i]);
    for (int k = 0; k < m; ++k) logdelta[k] = log(delta[k]);
    double *nu = (double *)malloc(sizeof(double) * (size_t)(n * m));
    if (!nu) return -1;
    double e[16];
    for (int64_t i = 0; i < n; ++i) {
        if (!isfinite(x[i])) { free(nu); return -4; }
        ldouble s = 0.0L;
        for (int k = 0; k < m; ++k) {
            double z = fabs(x[i] - mean[k]) / sdm;
            double lq = orc_pnorm_upper_log(z);
            e[k] = 1.0 / (-1.0 * lq);
            s += e[k];
        }
        double sum = (double)s;
        for (int k = 0; k < m; ++k) e[k] = log(e[k] / sum);
        if (i == 0) {
            for (int k = 0; k < m; ++k) nu[k] = logdelta[k] + e[k];
        } else {
            const double *prev = nu + (i - 1) * m;
            for (int k = 0; k < m; ++k) {
                double best = prev[0] + logPi[0 + m * k];
                for (int j = 1; j < m; ++j) {
                    double v = prev[j] + logPi[j + m * k];
                    if (v > best) best = v;
                }
                nu[i * m + k] = best + e[k];
            }
        }
    }
    const double *last = nu + (n - 1) * m;
    for (int k = 0; k < m; ++k)
        if (last[k] == -INFINITY) { free(nu); return -3; }
    double mm = INFINITY;
    int y = 0;
    {
        double best = last[0], second = -INFINITY;
        for (int k = 1; k < m; ++k) {
            if (last[k] > best) { second = best; best = last[k]; y = k; }
            else if (last[k] > second) second = last[k];
        }
        if (best - second < mm) mm = best - second;
    }
    states[n - 1] = y + 1;
    for (int64_t i = n - 2; i >= 0; --i) {
        const double *row = nu + i * m;
        int arg = 0;
        double best = logPi[0 + m * y] + row[0], second = -INFINITY;
        for (int j = 1; j < m; ++j) {
            double v = logPi[j + m * y] + row[j];
            if (v > best) { second = best; best = v; arg = j; }
            else if (v > second) second = v;
        }
        if (best - second < mm) mm = best - second;
        y = arg;
        states[i] = y + 1;
    }
    if (min_margin) *min_margin = mm;
    free(nu);
    return 0;
}

/* Drivers: predict_CNV_via_HMM_on_indiv_cells (R/inferCNV_HMM.R:284-324, i3 twin
 * R/inferCNV_i3HMM.R:180-225) when n_grp == 0; group modes (HMM.R:345-408, 509-567; i3HMM.R:249-389)
 * when n_grp > 0: x = rowMeans(X[chr, group]) (LD sum / count), per-group sd (sd + m*b), trace
 * written to every cell of the group.  Cells in no group keep -1 (HMM.R:368 init).
 * states: int32 G x C.  margins (optional): per (sequence) min margin, size K*C or K*n_grp. */
ORC_API int orc_viterbi_matrix(const double *X, int64_t G, int64_t C, const int32_t *chr_start,
                               const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx,
                               int n_grp, int m, const double *Pi, const double *delta, const double *mean,
                               const double *sd, int32_t *states, double *margins, int nthreads) {
    int rc_all = 0;
    (void)nthreads;
    if (n_grp == 0) {
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
        for (int64_t c = 0; c < C; ++c) {
            for (int k = 0; k < K; ++k) {
                double mg;
                int rc = orc_viterbi_seq(X + chr_start[k] + G * c, chr_len[k], m, Pi, delta, mean, sd,
                                         states + chr_start[k] + G * c, &mg);
                if (margins) margins[k + (int64_t)K * c] = mg;
                if (rc) {
#pragma omp critical
                    rc_all = rc;
                }
            }
        }
        return rc_all;
    }
    for (int64_t i = 0; i < G * C; ++i) states[i] = -1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
    for (int b = 0; b < n_grp; ++b) {
        int64_t cnt = grp_off[b + 1] - grp_off[b];
        const int32_t *idx = grp_idx + grp_off[b];
        double *x = (double *)malloc(sizeof(double) * (size_t)G);
        int32_t *st = (int32_t *)malloc(sizeof(int32_t) * (size_t)G);
        for (int64_t g = 0; g < G; ++g) { /* rowMeans: LD sum / n, no refinement */
            ldouble s = 0.0L;
            for (int64_t i = 0; i < cnt; ++i) s += X[g + G * (int64_t)idx[i]];
            x[g] = (double)(s / (ldouble)cnt);
        }
        for (int k = 0; k < K; ++k) {
            double mg;
            int rc = orc_viterbi_seq(x + chr_start[k], chr_len[k], m, Pi, delta, mean, sd + (int64_t)m * b,
                                     st + chr_start[k], &mg);
            if (margins) margins[k + (int64_t)K * b] = mg;
            if (rc) {
#pragma omp critical
                rc_all = rc;
            }
        }
        for (int64_t i = 0; i < cnt; ++i)
            for (int k = 0; k < K; ++k)
                for (int64_t g = chr_start[k]; g < chr_start[k] + chr_len[k]; ++g)
                    states[g + G * (int64_t)idx[i]] = st[g];
        free(x);
        free(st);
    }
    return rc_all;
}

/* ------------------------------------------------------------------------------------------- */
/* i3 parameterisation + denoise (reductions over reference cells)                              */
/* ------------------------------------------------------------------------------------------- */

/* R/inferCNV_i3HMM.R:17-30: mu = mean(X[, cells]), sigma = sd(X[, cells]) over ALL values. */
ORC_API int orc_mean_sd_over_cells(const double *X, int64_t G, const int32_t *idx, int64_t n_idx, double *mu,
                                   double *sigma) {
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(G * n_idx));
    if (!tmp) return -1;
    for (int64_t i = 0; i < n_idx; ++i) memcpy(tmp + G * i, X + G * (int64_t)idx[i], sizeof(double) * (size_t)G);
    *mu = r_mean(tmp, G * n_idx);
    *sigma = r_sd(tmp, G * n_idx);
    free(tmp);
    return 0;
}

/* parallelDist(t(X[, cells])), method "euclidean" - the input of every hclust() in the reference
 * (R/inferCNV_tumor_subclusters.R:191,411,472,582,609; R/inferCNV_ops.R:1930,3242; R/inferCNV_heatmap.R:719,755,1062,1079).
 * parallelDist (CRAN, an Imports: dependency of the reference's DESCRIPTION, not vendored under /root/reference; no version
 * pin there) computes sqrt(accu(square(A - B))) per pair with Armadillo; stats::dist (R's src/library/stats/src/distance.c,
 * R_euclidean) computes the same sum gene by gene: dev = x[i] - x[j]; dist += dev * dev; sqrt(dist).  This restates the latter
 * (the two differ only in the order of the additions, ~1e-16 relative); the result is R's "dist" vector: the strict lower
 * triangle by columns.  cells = NULL: all C columns.  PARITY UNPINNED by the reference (it holds no distance vectors):
 * tests/test_oracle_dist.py pins this against scipy.spatial.distance.pdist instead. */
ORC_API int orc_pairwise_dist(const double *X, int64_t G, int64_t C, const int32_t *cells, int64_t n, double *out, int nthreads) {
    if (!cells) n = C;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (int64_t a = 0; a < n; ++a) {
        const double *xa = X + G * (cells ? (int64_t)cells[a] : a);
        double *o = out + n * a - a * (a + 1) / 2 - a - 1;
        for (int64_t b = a + 1; b < n; ++b) {
            const double *xb = X + G * (cells ? (int64_t)cells[b] : b);
            double d = 0.0;
            for (int64_t g = 0; g < G; ++g) {
                const double dev = xa[g] - xb[g];
                d += dev * dev;
            }
            o[b] = sqrt(d);
        }
    }
    return 0;
}

/* R/inferCNV_ops.R:2302-2346 clear_noise_via_ref_mean_sd (noise_logistic = FALSE):
 * mu = mean(X[, ref]); s = mean_c sd(X[, ref_c]) * sd_amplifier; values strictly inside
 * (mu - s, mu + s) become mu.  Used only to check the bundled golden end to end. */
ORC_API int orc_clear_noise_via_ref_mean_sd(const double *X, double *Y, int64_t G, int64_t C, const int32_t *idx,
                                            int64_t n_idx, double sd_amplifier) {
    double mu, dummy;
    if (orc_mean_sd_over_cells(X, G, idx, n_idx, &mu, &dummy)) return -1;
    double *sds = (double *)malloc(sizeof(double) * (size_t)n_idx);
    for (int64_t i = 0; i < n_idx; ++i) sds[i] = r_sd(X + G * (int64_t)idx[i], G);
    double s = r_mean(sds, n_idx) * sd_amplifier;
    free(sds);
    double hi = mu + s, lo = mu - s;
    for (int64_t i = 0; i < G * C; ++i) {
        double v = X[i];
        Y[i] = (v > lo && v < hi) ? mu : v;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* the whole smooth block, run() steps 4..14 (R/inferCNV_ops.R:614-1031)                        */
/* ------------------------------------------------------------------------------------------- */

/* X: depth-normalised expression (after step 3).  Steps: log2(x+1) [if apply_log], subtract ref
 * (bounds), clamp +-threshold, smooth(window), centre by median, subtract ref again, 2^x. */
ORC_API int orc_smooth_block(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                             const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx, int n_grp,
                             int apply_log, double threshold, int window, int use_bounds, int nthreads) {
    int64_t n = G * C;
    orc_set_num_threads(nthreads);
    double *means = (double *)malloc(sizeof(double) * (size_t)(G * n_grp));
    if (!means) return -1;
    if (apply_log) orc_log2xplus1(X, Y, n);
    else if (X != Y) memcpy(Y, X, sizeof(double) * (size_t)n);
    orc_ref_means(Y, G, C, grp_off, grp_idx, n_grp, 0, means);
    orc_subtract_ref(Y, Y, G, C, means, n_grp, use_bounds);
    orc_apply_max_threshold_bounds(Y, Y, n, threshold);
    double *T = (double *)malloc(sizeof(double) * (size_t)n);
    if (!T) { free(means); return -1; }
    int rc = orc_smooth_by_chromosome(Y, T, G, C, chr_start, chr_len, K, window, 0, nthreads);
    if (rc) { free(means); free(T); return rc; }
    orc_center_columns(T, T, G, C, 1, nthreads);
    orc_ref_means(T, G, C, grp_off, grp_idx, n_grp, 0, means);
    orc_subtract_ref(T, Y, G, C, means, n_grp, use_bounds);
    orc_invert_log2(Y, Y, n);
    free(T);
    free(means);
    return 0;
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Number of processors the machine offers, NOT what OMP_NUM_THREADS asks for (torchrun exports OMP_NUM_THREADS=1 to
 * every rank, which would silently turn the CPU arm of bench.py into a one-thread run). */
ORC_API int orc_num_procs(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* ---- synthetic workload (bench.py's CPU arm): CPU twin of the device generator -------------------------------------
 * Restates infercnv_b200/csrc/icnv_synth.cu (not a reference function: the workload model of SURVEY section 8d) so that
 * the CPU arm can draw the same cells without loading the product library.  Every value is a pure function of
 * (seed, global cell, gene).  Values are counts; libm differences between host and device can only move a count when a
 * uniform lands within an ulp of a CDF step. */
static uint64_t syn_mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static uint64_t syn_hash3(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
    return syn_mix64(syn_mix64(syn_mix64(seed ^ 0x243f6a8885a308d3ull) + a) * 0x9fb21c651e98df25ull + b) ^
           syn_mix64(c + 0x13198a2e03707344ull);
}
static double syn_u01(uint64_t h) { return ((double)(h >> 11) + 1.0) * (1.0 / 9007199254740992.0); }
static double syn_normal(uint64_t h1, uint64_t h2) {
    return sqrt(-2.0 * log(syn_u01(h1))) * cos(6.283185307179586476925286766559 * syn_u01(h2));
}

ORC_API int orc_synth(double *X, int64_t G, const int64_t *cells, int64_t n_cells, int64_t C_total, const int32_t *chr_start,
                      const int32_t *chr_len, int K, uint64_t seed, int nthreads) {
    int32_t *chr_of = (int32_t *)calloc((size_t)G, sizeof(int32_t));
    double *m_g = (double *)malloc(sizeof(double) * (size_t)G);
    if (!chr_of || !m_g) { free(chr_of); free(m_g); return -1; }
    for (int k = 0; k < K; ++k)
        for (int32_t g = chr_start[k]; g < chr_start[k] + chr_len[k] && g < G; ++g) chr_of[g] = k;
    for (int64_t g = 0; g < G; ++g)
        m_g[g] = exp(0.5 + syn_normal(syn_hash3(seed, 4, (uint64_t)g, 0), syn_hash3(seed, 4, (uint64_t)g, 1)));
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t ci = 0; ci < n_cells; ++ci) {
        const uint64_t cell = (uint64_t)cells[ci];
        const double f_c = exp(0.2 * syn_normal(syn_hash3(seed, 1, cell, 0), syn_hash3(seed, 1, cell, 1)));
        int ev_chr[3] = {-1, -1, -1};
        double ev_mul[3] = {1.0, 1.0, 1.0};
        if ((int64_t)cell >= C_total / 10 && syn_u01(syn_hash3(seed, 2, cell, 0)) < 0.3)
            for (int e = 0; e < 3; ++e) {
                const uint64_t h = syn_hash3(seed, 3, cell, (uint64_t)e);
                ev_chr[e] = (int)(h % (uint64_t)K);
                ev_mul[e] = ((h >> 40) & 1ull) ? 1.5 : 0.5;
            }
        double *col = X + G * ci;
        for (int64_t g = 0; g < G; ++g) {
            double cnv = 1.0;
            for (int e = 0; e < 3; ++e)
                if (chr_of[g] == ev_chr[e]) cnv = ev_mul[e];
            const double mean = m_g[g] * f_c * cnv;
            double prod = 1.0;
            for (int i = 0; i < 10; ++i) prod *= syn_u01(syn_hash3(seed, 5 + (uint64_t)i, cell, (uint64_t)g));
            const double lambda = -log(prod) * (mean * 0.1);
            double x;
            if (lambda < 12.0) {
                const double u = syn_u01(syn_hash3(seed, 20, cell, (uint64_t)g));
                double pmf = exp(-lambda), cdf = pmf;
                int k = 0;
                while (u > cdf && k < 64) {
                    ++k;
                    pmf *= lambda / (double)k;
                    cdf += pmf;
                }
                x = (double)k;
            } else {
                const double z = syn_normal(syn_hash3(seed, 21, cell, (uint64_t)g), syn_hash3(seed, 22, cell, (uint64_t)g));
                x = floor(lambda + sqrt(lambda) * z + 0.5);
                if (x < 0.0) x = 0.0;
            }
            col[g] = x;
        }
    }
    free(chr_of);
    free(m_g);
    return 0;
}
