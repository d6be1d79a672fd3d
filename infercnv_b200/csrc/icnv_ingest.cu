// icnv_ingest.cu - the steps in front of the path: gene filters on the raw counts
// (run() step 2) and counts ingest incl. sparse input, so the matrix can cross PCIe once, as counts.
//
//   .below_min_mean_expr_cutoff   R/inferCNV_ops.R:2149-2158  rowMeans(expr) < cutoff          gene_stats_*_kernel
//   require_above_min_cells_ref   R/inferCNV_ops.R:2177-2209  sum(x > 0 & !is.na(x)) >= n      gene_stats_*_kernel
//   remove_genes                  R/inferCNV.R:445-457        expr.data[-idx, ]                gather_rows_kernel
//   dgCMatrix counts (R/inferCNV.R:158-160) -> the same statistics and the depth-normalised dense matrix
//   (.normalize_data_matrix_by_seq_depth, ops.R:3082-3111) straight from the compressed columns: csc_*_kernel
//
// All HBM-bound streaming / scatter work.  Per-gene sums over a dense matrix are accumulated as fixed chunks of
// 32 consecutive cells combined in order (the same scheme as the group means of icnv_smooth.cu), so they do not
// depend on the launch geometry; the sparse variant uses atomics, which is exact for count data (integer sums
// below 2^53 do not depend on the order).
#include <algorithm>
#include <cmath>
#include <vector>

#include "icnv_common.cuh"

namespace icnv {

constexpr int IG_CHUNK = 32;      // cells per partial sum (fixed: results must not depend on the grid)

// psum[g + G*q], ppos[g + G*q]: sum / number of values > 0 of gene g over cells [q*IG_CHUNK, (q+1)*IG_CHUNK)
__global__ void __launch_bounds__(256) gene_stats_partial_kernel(const double *__restrict__ X, int64_t G, int64_t ldx,
                                                                 int64_t C, double *__restrict__ psum,
                                                                 uint32_t *__restrict__ ppos) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= G) return;
    for (int64_t q = blockIdx.y; q * IG_CHUNK < C; q += gridDim.y) {
        const int64_t c0 = q * IG_CHUNK, c1 = min(C, c0 + IG_CHUNK);
        double s = 0.0;
        uint32_t n = 0;
        // eight independent loads in flight per thread (the counted loop issued load, add, load, add: 0.48 of the HBM
        // roofline measured).  The loads are unconditional - past the chunk's end they re-read its last cell - so that they
        // can be issued back to back; only the accumulation is predicated, and the sum keeps its cell order.
        const int nvalid = (int)(c1 - c0);
        const double *__restrict__ col0 = X + g + ldx * c0;
        for (int j0 = 0; j0 < IG_CHUNK; j0 += 8) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = col0[ldx * (int64_t)min(j0 + j, nvalid - 1)];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j0 + j < nvalid) {
                    s += v[j];
                    n += (v[j] > 0.0) ? 1u : 0u;      // NaN > 0 is false: x > 0 & !is.na(x)
                }
        }
        psum[g + G * q] = s;
        ppos[g + G * q] = n;
    }
}

__global__ void __launch_bounds__(256) gene_stats_combine_kernel(const double *__restrict__ psum,
                                                                 const uint32_t *__restrict__ ppos, int64_t G,
                                                                 int64_t n_chunks, double *__restrict__ sums,
                                                                 int32_t *__restrict__ npos) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= G) return;
    double s = 0.0;
    uint32_t n = 0;
    for (int64_t q = 0; q < n_chunks; ++q) {
        s += psum[g + G * q];
        n += ppos[g + G * q];
    }
    sums[g] = s;
    npos[g] = (int32_t)n;
}

// remove_genes: Y[i + n_keep*c] = X[keep[i] + ldx*c]
__global__ void __launch_bounds__(256) gather_rows_kernel(const double *__restrict__ X, int64_t ldx,
                                                          const int32_t *__restrict__ keep, int64_t n_keep,
                                                          double *__restrict__ Y, int64_t C) {
    for (int64_t c = blockIdx.x; c < C; c += gridDim.x) {
        const double *col = X + ldx * c;
        double *out = Y + n_keep * c;
        for (int64_t i = threadIdx.x; i < n_keep; i += 256) out[i] = col[keep[i]];
    }
}

// compressed sparse columns (dgCMatrix: p = column pointers, i = row indices, x = values)
__global__ void __launch_bounds__(256) csc_gene_stats_kernel(const int32_t *__restrict__ ri, const double *__restrict__ x,
                                                             int64_t nnz, double *__restrict__ sums,
                                                             int32_t *__restrict__ npos) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) {
        const double v = x[k];
        atomicAdd(sums + ri[k], v);
        if (v > 0.0) atomicAdd(npos + ri[k], 1);
    }
}

// one warp per column: cs[c] = sum of the column's kept entries (lane-strided partial sums + a fixed shuffle tree)
__global__ void __launch_bounds__(256) csc_col_sums_kernel(const int32_t *__restrict__ p, const int32_t *__restrict__ ri,
                                                           const double *__restrict__ x,
                                                           const int32_t *__restrict__ keep_map, int64_t C,
                                                           double *__restrict__ cs) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < C; c += warps) {
        double s = 0.0;
        for (int32_t k = p[c] + lane; k < p[c + 1]; k += 32)
            if (!keep_map || keep_map[ri[k]] >= 0) s += x[k];
        s = warp_sum_d(s);
        if (lane == 0) cs[c] = s;
    }
}

// one CTA per column: Y[:, c] = (0 / cs) * factor everywhere, then (x / cs) * factor at the kept stored entries
__global__ void __launch_bounds__(256) csc_expand_kernel(const int32_t *__restrict__ p, const int32_t *__restrict__ ri,
                                                         const double *__restrict__ x,
                                                         const int32_t *__restrict__ keep_map, int64_t G_out, int64_t C,
                                                         const double *__restrict__ cs, double factor,
                                                         double *__restrict__ Y) {
    for (int64_t c = blockIdx.x; c < C; c += gridDim.x) {
        const double s = cs[c];
        const double zero_val = (0.0 / s) * factor;     // NaN for an all-zero cell, as data / colSums does in R
        double *out = Y + G_out * c;
        for (int64_t g = threadIdx.x; g < G_out; g += 256) out[g] = zero_val;
        __syncthreads();
        for (int32_t k = p[c] + threadIdx.x; k < p[c + 1]; k += 256) {
            const int32_t r = keep_map ? keep_map[ri[k]] : ri[k];
            if (r >= 0) out[r] = (x[k] / s) * factor;
        }
        __syncthreads();
    }
}

// scale_infercnv_expr (R/inferCNV_ops.R:3174-3186: t(scale(t(expr)))) - per gene, over the cells: centre at the mean,
// divide by sqrt(sum((x - mean)^2) / (C - 1)).  The mean is refined once (mean0 + mean(x - mean0), what R's long-double
// accumulation amounts to in double), so a constant gene is centred to exact zeros and comes out NaN (0 / 0) as in R.
// Partial sums in the same fixed chunks as the plain sums.  SQUARE = false: sum of (x - mean); true: sum of squares.
template <bool SQUARE>
__global__ void __launch_bounds__(256) gene_centered_partial_kernel(const double *__restrict__ X, int64_t G, int64_t ldx,
                                                                    int64_t C, const double *__restrict__ mean,
                                                                    double *__restrict__ part) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= G) return;
    const double m = mean[g];
    for (int64_t q = blockIdx.y; q * IG_CHUNK < C; q += gridDim.y) {
        const int64_t c0 = q * IG_CHUNK, c1 = min(C, c0 + IG_CHUNK);
        double s = 0.0;
        for (int64_t c = c0; c < c1; ++c) {
            const double d = X[g + ldx * c] - m;
            s = SQUARE ? fma(d, d, s) : s + d;
        }
        part[g + G * q] = s;
    }
}

// mode 0: out[g] = sum / C; mode 1: out[g] += sum / C (mean refinement); mode 2: out[g] = sum
__global__ void __launch_bounds__(256) combine_partial_kernel(const double *__restrict__ part, int64_t G, int64_t n_chunks,
                                                              int64_t C, int mode, double *__restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= G) return;
    double s = 0.0;
    for (int64_t q = 0; q < n_chunks; ++q) s += part[g + G * q];
    if (mode == 0) out[g] = s / (double)C;
    else if (mode == 1) out[g] = out[g] + s / (double)C;
    else out[g] = s;
}

__global__ void __launch_bounds__(256) scale_rows_kernel(const double *__restrict__ X, double *__restrict__ Y, int64_t G, int64_t C,
                                                         const double *__restrict__ mean, const double *__restrict__ ss) {
    for (int64_t c = blockIdx.x; c < C; c += gridDim.x)
        for (int64_t g = threadIdx.x; g < G; g += 256) {
            const double sd = sqrt(ss[g] / (double)(C > 1 ? C - 1 : 1));     // scale(): max(1, n - 1)
            Y[g + G * c] = (X[g + G * c] - mean[g]) / sd;
        }
}

}  // namespace icnv

using namespace icnv;

extern "C" {

/* per-gene sum and number of positive values over the C columns of a device matrix */
int icnv_dev_gene_stats_f64(const double *X, int64_t G, int64_t ldx, int64_t C, double *d_sums, int32_t *d_npos,
                            void *stream) {
    ICNV_REQUIRE_READY();
    cudaStream_t st = pick_stream(stream);
    if (!X || !d_sums || !d_npos || G <= 0 || C <= 0 || ldx < G)
        return set_error(ICNV_E_BAD_ARG, "icnv_dev_gene_stats_f64: bad argument");
    const int64_t n_chunks = (C + IG_CHUNK - 1) / IG_CHUNK;
    char *buf = (char *)scratch(SLOT_PARTIAL, (size_t)G * (size_t)n_chunks * 12);
    if (!buf) return ICNV_E_NOMEM;
    double *psum = (double *)buf;
    uint32_t *ppos = (uint32_t *)(buf + (size_t)G * (size_t)n_chunks * 8);
    dim3 grid((unsigned)((G + 255) / 256), (unsigned)std::min<int64_t>(n_chunks, 65535));
    gene_stats_partial_kernel<<<grid, 256, 0, st>>>(X, G, ldx, C, psum, ppos);
    ICNV_CHECK_LAUNCH("gene_stats_partial_kernel");
    gene_stats_combine_kernel<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(psum, ppos, G, n_chunks, d_sums, d_npos);
    ICNV_CHECK_LAUNCH("gene_stats_combine_kernel");
    return ICNV_OK;
}

/* Y = per-gene z-scores of X over its C columns (X, Y: G x C, ld = G); d_mean, d_ss: scratch of G doubles each */
int icnv_dev_scale_rows_f64(const double *X, double *Y, int64_t G, int64_t C, double *d_mean, double *d_ss, void *stream) {
    ICNV_REQUIRE_READY();
    cudaStream_t st = pick_stream(stream);
    if (!X || !Y || !d_mean || !d_ss || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_dev_scale_rows_f64: bad argument");
    const int64_t n_chunks = (C + IG_CHUNK - 1) / IG_CHUNK;
    char *buf = (char *)scratch(SLOT_PARTIAL, (size_t)G * (size_t)n_chunks * 12);
    if (!buf) return ICNV_E_NOMEM;
    double *part = (double *)buf;
    uint32_t *ppos = (uint32_t *)(buf + (size_t)G * (size_t)n_chunks * 8);
    const dim3 grid((unsigned)((G + 255) / 256), (unsigned)std::min<int64_t>(n_chunks, 65535));
    const unsigned gblocks = (unsigned)((G + 255) / 256);
    gene_stats_partial_kernel<<<grid, 256, 0, st>>>(X, G, G, C, part, ppos);
    ICNV_CHECK_LAUNCH("gene_stats_partial_kernel");
    combine_partial_kernel<<<gblocks, 256, 0, st>>>(part, G, n_chunks, C, 0, d_mean);
    ICNV_CHECK_LAUNCH("combine_partial_kernel");
    gene_centered_partial_kernel<false><<<grid, 256, 0, st>>>(X, G, G, C, d_mean, part);
    ICNV_CHECK_LAUNCH("gene_centered_partial_kernel");
    combine_partial_kernel<<<gblocks, 256, 0, st>>>(part, G, n_chunks, C, 1, d_mean);
    ICNV_CHECK_LAUNCH("combine_partial_kernel");
    gene_centered_partial_kernel<true><<<grid, 256, 0, st>>>(X, G, G, C, d_mean, part);
    ICNV_CHECK_LAUNCH("gene_centered_partial_kernel");
    combine_partial_kernel<<<gblocks, 256, 0, st>>>(part, G, n_chunks, C, 2, d_ss);
    ICNV_CHECK_LAUNCH("combine_partial_kernel");
    const int64_t blocks = std::min<int64_t>(C, (int64_t)ctx().sm_count * 8);
    scale_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, Y, G, C, d_mean, d_ss);
    ICNV_CHECK_LAUNCH("scale_rows_kernel");
    return ICNV_OK;
}

int icnv_dev_gather_rows_f64(const double *X, int64_t ldx, const int32_t *d_keep, int64_t n_keep, double *Y, int64_t C,
                             void *stream) {
    ICNV_REQUIRE_READY();
    if (!X || !d_keep || !Y || n_keep <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_dev_gather_rows_f64: bad argument");
    const int64_t blocks = std::min<int64_t>(C, (int64_t)ctx().sm_count * 8);
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(X, ldx, d_keep, n_keep, Y, C);
    ICNV_CHECK_LAUNCH("gather_rows_kernel");
    return ICNV_OK;
}

int icnv_dev_csc_gene_stats_f64(const int32_t *d_i, const double *d_x, int64_t nnz, int64_t G, double *d_sums,
                                int32_t *d_npos, void *stream) {
    ICNV_REQUIRE_READY();
    cudaStream_t st = pick_stream(stream);
    if (!d_sums || !d_npos || G <= 0 || nnz < 0 || (nnz > 0 && (!d_i || !d_x)))
        return set_error(ICNV_E_BAD_ARG, "icnv_dev_csc_gene_stats_f64: bad argument");
    ICNV_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * (size_t)G, st));
    ICNV_CUDA(cudaMemsetAsync(d_npos, 0, sizeof(int32_t) * (size_t)G, st));
    if (nnz == 0) return ICNV_OK;
    const int64_t blocks = std::min<int64_t>((nnz + 255) / 256, (int64_t)ctx().sm_count * 16);
    csc_gene_stats_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_i, d_x, nnz, d_sums, d_npos);
    ICNV_CHECK_LAUNCH("csc_gene_stats_kernel");
    return ICNV_OK;
}

int icnv_dev_csc_col_sums_f64(const int32_t *d_p, const int32_t *d_i, const double *d_x, const int32_t *d_keep_map,
                              int64_t C, double *d_cs, void *stream) {
    ICNV_REQUIRE_READY();
    if (!d_p || !d_cs || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_dev_csc_col_sums_f64: bad argument");
    const int64_t blocks = std::min<int64_t>((C + 7) / 8, (int64_t)ctx().sm_count * 16);
    csc_col_sums_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(d_p, d_i, d_x, d_keep_map, C, d_cs);
    ICNV_CHECK_LAUNCH("csc_col_sums_kernel");
    return ICNV_OK;
}

int icnv_dev_csc_expand_f64(const int32_t *d_p, const int32_t *d_i, const double *d_x, const int32_t *d_keep_map,
                            int64_t G_out, int64_t C, const double *d_cs, double factor, double *Y, void *stream) {
    ICNV_REQUIRE_READY();
    if (!d_p || !d_cs || !Y || G_out <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_dev_csc_expand_f64: bad argument");
    const int64_t blocks = std::min<int64_t>(C, (int64_t)ctx().sm_count * 8);
    csc_expand_kernel<<<(unsigned)blocks, 256, 0, pick_stream(stream)>>>(d_p, d_i, d_x, d_keep_map, G_out, C, d_cs, factor, Y);
    ICNV_CHECK_LAUNCH("csc_expand_kernel");
    return ICNV_OK;
}

}  // extern "C"
