// icnv_reduce.cu - small reductions and element-wise passes around the hot path:
//   K6  per-cell sum / sd (column_stats_kernel): feeds mu / sigma over the reference cells
//       (.i3HMM_get_sd_trend_by_num_cells_fit, R/inferCNV_i3HMM.R:17-30) and the denoise threshold
//       (clear_noise_via_ref_mean_sd, R/inferCNV_ops.R:2302-2346); one CTA per cell with a fixed reduction
//       tree, so per-cell statistics - and anything combined from them in list order - do not depend on
//       how the cells are spread over GPUs
//   depth normalisation, denoise and the stand-alone element-wise steps
#include <algorithm>
#include <cmath>

#include "icnv_common.cuh"

namespace icnv {

// element-wise steps of run(): log2xplus1 (ops.R:2756-2769), invert_log2 (ops.R:2814-2826),
// apply_max_threshold_bounds (ops.R:2970-2983).  Inside the fused block these ride on the loads / stores of
// cell_pipeline_kernel; as stand-alone steps they are one streaming pass (16 B per element, HBM-bound).
__global__ void __launch_bounds__(256) elementwise_kernel(const double *__restrict__ X, double *__restrict__ Y, int64_t n,
                                                          int op, double param, int *err_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) {
        double v = X[i];
        bad |= !is_finite_d(v);
        if (op == 0) v = log2(v + 1.0);
        else if (op == 1) v = exp2(v);
        else v = fmin(fmax(v, -param), param);
        Y[i] = v;
    }
    if (bad && err_flag) atomicExch(err_flag, 1);
}

// ---- SURVEY section 8(f) "next" rows adjacent to the path: depth normalisation (step 3) and denoise (step 22) ----

// one CTA per listed cell: out[i] = (sum, sd with n-1) of the cell's G values; fixed reduction tree
__global__ void __launch_bounds__(256) column_stats_kernel(const double *__restrict__ X, int64_t G,
                                                           const int32_t *__restrict__ cells, int64_t n_cells,
                                                           double *__restrict__ sums, double *__restrict__ sds) {
    __shared__ double sh[8];
    __shared__ double bc;
    for (int64_t ci = blockIdx.x; ci < n_cells; ci += gridDim.x) {
        const double *col = X + G * (cells ? (int64_t)cells[ci] : ci);
        double s = 0.0;
        for (int64_t g = threadIdx.x; g < G; g += 256) s += col[g];
        s = warp_sum_d(s);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += sh[w];
            bc = t;
        }
        __syncthreads();
        const double total = bc;
        if (sums && threadIdx.x == 0) sums[ci] = total;
        if (sds) {
            const double mean = total / (double)G;
            double q = 0.0;
            for (int64_t g = threadIdx.x; g < G; g += 256) {
                const double d = col[g] - mean;
                q = fma(d, d, q);
            }
            q = warp_sum_d(q);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = q;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0.0;
                for (int w = 0; w < 8; ++w) t += sh[w];
                sds[ci] = sqrt(t / (double)(G - 1));
            }
        }
        __syncthreads();
    }
}

// .normalize_data_matrix_by_seq_depth (ops.R:3082-3111): (x / colSum) * normalize_factor
__global__ void __launch_bounds__(256) scale_columns_kernel(const double *__restrict__ X, double *__restrict__ Y, int64_t G,
                                                            int64_t C, const double *__restrict__ sums, double factor) {
    for (int64_t c = blockIdx.x; c < C; c += gridDim.x) {
        const double cs = sums[c];
        for (int64_t g = threadIdx.x; g < G; g += 256) Y[g + G * c] = (X[g + G * c] / cs) * factor;
    }
}

// clear_noise_via_ref_mean_sd (ops.R:2302-2346): values strictly inside (lo, hi) become mu
__global__ void __launch_bounds__(256) clear_noise_kernel(const double *__restrict__ X, double *__restrict__ Y, int64_t n,
                                                          double lo, double hi, double mu) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double v = X[i];
        Y[i] = (v > lo && v < hi) ? mu : v;
    }
}

// .get_average_bounds (ops.R:2734-2742): quantile(x, na.rm=TRUE)[[1]] / [[5]] of a cell = its smallest / largest
// non-NA value.  One CTA per cell; min / max do not depend on the order.
__global__ void __launch_bounds__(256) column_minmax_kernel(const double *__restrict__ X, int64_t G, int64_t C,
                                                            double *__restrict__ mins, double *__restrict__ maxs) {
    __shared__ double sh_lo[8], sh_hi[8];
    for (int64_t c = blockIdx.x; c < C; c += gridDim.x) {
        const double *col = X + G * c;
        double lo = INFINITY, hi = -INFINITY;
        for (int64_t g = threadIdx.x; g < G; g += 256) {
            const double v = col[g];
            lo = fmin(lo, v);      // fmin / fmax return the non-NaN operand: na.rm = TRUE
            hi = fmax(hi, v);
        }
        lo = warp_min_d(lo);
        hi = warp_max_d(hi);
        if ((threadIdx.x & 31) == 0) {
            sh_lo[threadIdx.x >> 5] = lo;
            sh_hi[threadIdx.x >> 5] = hi;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) {
                lo = fmin(lo, sh_lo[w]);
                hi = fmax(hi, sh_hi[w]);
            }
            mins[c] = lo;
            maxs[c] = hi;
        }
        __syncthreads();
    }
}

// .remove_outliers_norm (ops.R:2051-2052): data[data < lower] <- lower; data[data > upper] <- upper
__global__ void __launch_bounds__(256) clamp_bounds_kernel(const double *__restrict__ X, double *__restrict__ Y, int64_t n,
                                                           double lower, double upper) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double v = X[i];
        if (v < lower) v = lower;
        if (v > upper) v = upper;
        Y[i] = v;
    }
}

// assign_HMM_states_to_proxy_expr_vals (R/inferCNV_HMM.R:1191-1206; m = 6: 1..6 -> 0, .5, 1, 1.5, 2, 3) and
// i3HMM_assign_HMM_states_to_proxy_expr_vals (R/inferCNV_i3HMM.R:405-417; m = 3: 1..3 -> .5, 1, 1.5); the sequence of
// masked assignments there never re-maps a value it has just written, so it is a lookup; other values pass through
__global__ void __launch_bounds__(256) proxy_vals_kernel(const double *__restrict__ X, double *__restrict__ Y, int64_t n,
                                                         int m) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double v = X[i];
        double out = v;
        if (m == 6) {
            if (v == 1.0) out = 0.0;
            else if (v == 2.0) out = 0.5;
            else if (v == 3.0) out = 1.0;
            else if (v == 4.0) out = 1.5;
            else if (v == 5.0) out = 2.0;
            else if (v == 6.0) out = 3.0;
        } else {
            if (v == 1.0) out = 0.5;
            else if (v == 2.0) out = 1.0;
            else if (v == 3.0) out = 1.5;
        }
        Y[i] = out;
    }
}

// .apply_logistic_val_adj (R/inferCNV_heatmap.R:2792-2810) with .logistic (R/SplatterScrape.R:210-212):
// val = |x - mean|; p = 1 / (1 + exp(-slope (val - midpt))); x -> mean +- p val
__global__ void __launch_bounds__(256) logistic_adj_kernel(const double *__restrict__ X, double *__restrict__ Y, int64_t n,
                                                           double expr_mean, double delta_midpt, double slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double x = X[i];
        const double val = fabs(x - expr_mean);
        const double p = 1.0 / (1.0 + exp(-slope * (val - delta_midpt)));
        double out = x;
        if (x > expr_mean) out = expr_mean + p * val;
        else if (x < expr_mean) out = expr_mean - p * val;
        Y[i] = out;
    }
}

}  // namespace icnv

using namespace icnv;

extern "C" ICNV_API int icnv_dev_column_stats_f64(const double *X, int64_t G, const int32_t *cells, int64_t n_cells, double *sums,
                                         double *sds, void *stream) {
    ICNV_REQUIRE_READY();
    if (!X || G <= 1 || n_cells <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_dev_column_stats_f64: bad argument");
    int64_t blocks = std::min<int64_t>(n_cells, (int64_t)ctx().sm_count * 8);
    column_stats_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(X, G, cells, n_cells, sums, sds);
    ICNV_CHECK_LAUNCH("column_stats_kernel");
    return ICNV_OK;
}

extern "C" int icnv_dev_scale_columns_f64(const double *X, double *Y, int64_t G, int64_t C, const double *sums, double factor,
                                          void *stream) {
    ICNV_REQUIRE_READY();
    int64_t blocks = std::min<int64_t>(C, (int64_t)ctx().sm_count * 8);
    scale_columns_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(X, Y, G, C, sums, factor);
    ICNV_CHECK_LAUNCH("scale_columns_kernel");
    return ICNV_OK;
}

extern "C" int icnv_dev_clear_noise_f64(const double *X, double *Y, int64_t n, double lo, double hi, double mu, void *stream) {
    ICNV_REQUIRE_READY();
    int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 16);
    clear_noise_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(X, Y, n, lo, hi, mu);
    ICNV_CHECK_LAUNCH("clear_noise_kernel");
    return ICNV_OK;
}

extern "C" int icnv_dev_elementwise_f64(const double *X, double *Y, int64_t n, int op, double param, int *err_flag,
                                        void *stream) {
    ICNV_REQUIRE_READY();
    if (!X || !Y || n <= 0 || op < 0 || op > 2) return set_error(ICNV_E_BAD_ARG, "icnv_dev_elementwise_f64: bad argument");
    int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 16);
    elementwise_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(X, Y, n, op, param, err_flag);
    ICNV_CHECK_LAUNCH("elementwise_kernel");
    return ICNV_OK;
}

extern "C" int icnv_dev_column_minmax_f64(const double *X, int64_t G, int64_t C, double *mins, double *maxs, void *stream) {
    ICNV_REQUIRE_READY();
    if (!X || !mins || !maxs || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_dev_column_minmax_f64: bad argument");
    int64_t blocks = std::min<int64_t>(C, (int64_t)ctx().sm_count * 8);
    column_minmax_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(X, G, C, mins, maxs);
    ICNV_CHECK_LAUNCH("column_minmax_kernel");
    return ICNV_OK;
}

extern "C" int icnv_dev_clamp_bounds_f64(const double *X, double *Y, int64_t n, double lower, double upper, void *stream) {
    ICNV_REQUIRE_READY();
    if (!X || !Y || n <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_dev_clamp_bounds_f64: bad argument");
    int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 16);
    clamp_bounds_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(X, Y, n, lower, upper);
    ICNV_CHECK_LAUNCH("clamp_bounds_kernel");
    return ICNV_OK;
}

extern "C" int icnv_dev_logistic_adj_f64(const double *X, double *Y, int64_t n, double expr_mean, double delta_midpt,
                                         double slope, void *stream) {
    ICNV_REQUIRE_READY();
    if (!X || !Y || n <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_dev_logistic_adj_f64: bad argument");
    int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 16);
    logistic_adj_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(X, Y, n, expr_mean, delta_midpt, slope);
    ICNV_CHECK_LAUNCH("logistic_adj_kernel");
    return ICNV_OK;
}

extern "C" int icnv_dev_proxy_vals_f64(const double *X, double *Y, int64_t n, int m, void *stream) {
    ICNV_REQUIRE_READY();
    if (!X || !Y || n <= 0 || (m != 6 && m != 3)) return set_error(ICNV_E_BAD_ARG, "icnv_dev_proxy_vals_f64: bad argument");
    int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 16);
    proxy_vals_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(X, Y, n, m);
    ICNV_CHECK_LAUNCH("proxy_vals_kernel");
    return ICNV_OK;
}
