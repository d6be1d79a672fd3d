// icnv_smooth.cu - the smooth block of infercnv::run() as sm_100a kernels.
//
//   K1  group_partial_sums / combine_partials / bounds_from_means
//         .get_normal_gene_mean_bounds (R/inferCNV_ops.R:1708-1735): per-gene mean over each
//         reference group, summed in a fixed chunk order so the result does not depend on how
//         the cells are spread over GPUs.
//   K2  cell_pipeline_kernel: one CTA per cell, the cell's whole gene vector resident in shared
//         memory, all of [log2(x+1)] -> .subtract_expr (ops.R:1742-1786) -> clamp (ops.R:2970-2983)
//         -> .smooth_helper pyramid (ops.R:2483-2532, 2640-2661) -> .center_columns median
//         (ops.R:2094-2109) -> .subtract_expr again -> 2^x (ops.R:2814-2826) between ONE read and
//         ONE write of the column.
//
// Data layout: X[g + ld*c], a cell's genes contiguous (R column-major).  HBM-bound integer-free
// streaming work: no tensor cores; what matters is coalesced 8/16-byte accesses, the column kept
// on chip between the stages, and a grid that fills 148 SMs x resident CTAs.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "icnv_common.cuh"

namespace icnv {

// =================================================================================================
// K1: group means
// =================================================================================================

// grid: (ceil(G/256), n_chunks); thread = gene, block row = chunk of list entries.
// Adjacent threads read adjacent genes of the same cell: fully coalesced.  The sum over the
// chunk's cells runs in list order in every configuration (determinism across GPU counts).
__global__ void __launch_bounds__(256) group_partial_sums_kernel(const double *__restrict__ X, int64_t G, int64_t ldx,
                                                                 const int32_t *__restrict__ cells, int64_t n_cells,
                                                                 int chunk, int apply_log,
                                                                 double *__restrict__ partial) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int64_t q = blockIdx.y;
    int64_t i0 = q * chunk;
    int64_t i1 = i0 + chunk < n_cells ? i0 + chunk : n_cells;
    double s = 0.0;
    int64_t i = i0;
    // 4 independent loads in flight per thread, added in list order
    for (; i + 4 <= i1; i += 4) {
        double v0 = X[g + ldx * (int64_t)cells[i]];
        double v1 = X[g + ldx * (int64_t)cells[i + 1]];
        double v2 = X[g + ldx * (int64_t)cells[i + 2]];
        double v3 = X[g + ldx * (int64_t)cells[i + 3]];
        if (apply_log) {
            v0 = log2(v0 + 1.0);
            v1 = log2(v1 + 1.0);
            v2 = log2(v2 + 1.0);
            v3 = log2(v3 + 1.0);
        }
        s += v0;
        s += v1;
        s += v2;
        s += v3;
    }
    for (; i < i1; ++i) {
        double v = X[g + ldx * (int64_t)cells[i]];
        if (apply_log) v = log2(v + 1.0);
        s += v;
    }
    partial[g + G * q] = s;
}

__global__ void __launch_bounds__(256) combine_partials_kernel(const double *__restrict__ partial, int64_t G,
                                                               int64_t n_chunks, double inv_count_num,
                                                               double *__restrict__ means) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double s = 0.0;
    for (int64_t q = 0; q < n_chunks; ++q) s += partial[g + G * q];
    means[g] = s / inv_count_num;  // a true division by the count, as mean() does
}

__global__ void __launch_bounds__(256) bounds_from_means_kernel(const double *__restrict__ means, int64_t G, int n_grp,
                                                                double *__restrict__ lo, double *__restrict__ hi,
                                                                double *__restrict__ mid) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double mn = means[g], mx = means[g], s = 0.0;
    for (int k = 0; k < n_grp; ++k) {
        double m = means[g + G * k];
        mn = fmin(mn, m);
        mx = fmax(mx, m);
        s += m;
    }
    lo[g] = mn;
    hi[g] = mx;
    if (mid) mid[g] = s / (double)n_grp;
}

// inv_log variant of the group mean: log2(mean(2^x - 1) + 1) (ops.R:1714-1717)
__global__ void __launch_bounds__(256) group_partial_sums_invlog_kernel(const double *__restrict__ X, int64_t G,
                                                                        int64_t ldx, const int32_t *__restrict__ cells,
                                                                        int64_t n_cells, int chunk,
                                                                        double *__restrict__ partial) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int64_t q = blockIdx.y;
    int64_t i0 = q * chunk;
    int64_t i1 = i0 + chunk < n_cells ? i0 + chunk : n_cells;
    double s = 0.0;
    for (int64_t i = i0; i < i1; ++i) s += exp2(X[g + ldx * (int64_t)cells[i]]) - 1.0;
    partial[g + G * q] = s;
}

__global__ void __launch_bounds__(256) invlog_finish_kernel(double *__restrict__ means, int64_t n) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) means[g] = log2(means[g] + 1.0);
}

// =================================================================================================
// K2: fused per-cell pipeline
// =================================================================================================

struct Seg {
    int start;  // first gene of the segment
    int len;    // number of genes (0 = idle thread)
    int cs;     // chromosome start
    int ce;     // chromosome end (exclusive)
};

struct CellParams {
    const double *X;
    int64_t G, ldx;
    const int32_t *cols;
    int64_t n_cols;
    double *Y;
    int64_t ldy;
    const Seg *segs;
    int apply_log;
    const double *lo1, *hi1, *mid1;
    double threshold;
    int window, h;
    int center;  // 0 none, 1 median, 2 mean
    const double *lo2, *hi2, *mid2;
    int apply_exp2;
    int *err_flag;
    int s_elems;  // doubles reserved for the column (G rounded up to even)
};

constexpr int LMAX = 24;       // genes per thread kept in registers across the median
constexpr int CAND_MAX = 64;   // candidates ranked directly at the end of the selection

// ---- block-wide reductions with one __syncthreads each (double-buffered scratch) ----------------
template <int NW>
struct Red {
    double d[2][2][NW];
    int i[2][2][NW];
};

template <int NW>
__device__ __forceinline__ void block_sum2i(Red<NW> &r, int &phase, int a, int b, int &A, int &B) {
    a = __reduce_add_sync(0xffffffffu, a);
    b = __reduce_add_sync(0xffffffffu, b);
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        r.i[phase][0][w] = a;
        r.i[phase][1][w] = b;
    }
    __syncthreads();
    int sa = 0, sb = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        sa += r.i[phase][0][k];
        sb += r.i[phase][1][k];
    }
    A = sa;
    B = sb;
    phase ^= 1;
}

// op: 0 = (min, max), 1 = (sum, sum), 2 = (max, min)
template <int NW, int OP>
__device__ __forceinline__ void block_red2d(Red<NW> &r, int &phase, double a, double b, double &A, double &B) {
    if (OP == 0) {
        a = warp_min_d(a);
        b = warp_max_d(b);
    } else if (OP == 1) {
        a = warp_sum_d(a);
        b = warp_sum_d(b);
    } else {
        a = warp_max_d(a);
        b = warp_min_d(b);
    }
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        r.d[phase][0][w] = a;
        r.d[phase][1][w] = b;
    }
    __syncthreads();
    double sa = r.d[phase][0][0], sb = r.d[phase][1][0];
#pragma unroll
    for (int k = 1; k < NW; ++k) {
        double ta = r.d[phase][0][k], tb = r.d[phase][1][k];
        if (OP == 0) {
            sa = fmin(sa, ta);
            sb = fmax(sb, tb);
        } else if (OP == 1) {
            sa += ta;
            sb += tb;
        } else {
            sa = fmax(sa, ta);
            sb = fmin(sb, tb);
        }
    }
    A = sa;
    B = sb;
    phase ^= 1;
}

// ---- exact median of the n values spread over the CTA's registers -------------------------------
// Every thread holds `len` values in y[0..len).  Returns median per R's median.default: the middle
// order statistic for odd n, the mean of the two middle ones for even n (ops.R:2098).
//
// Selection by counting: a bracket (lo, hi] known to contain both middle order statistics is
// narrowed with two pivots per round (one pass over the registers, one block reduction).  Pivots
// come from linear interpolation of the empirical CDF inside the bracket; a round that fails to
// halve the bracket is followed by a bisection round in key space, which bounds the worst case.
// Once <= CAND_MAX values remain they are gathered into shared memory and ranked directly.
template <int NT>
__device__ double block_median(const double (&y)[LMAX], int len, int n, Red<NT / 32> &red, int &phase,
                               double *cand, int *cand_n) {
    constexpr int NW = NT / 32;
    const int kA = (n - 1) >> 1, kB = n >> 1;

    // start: mean / sd bracket guess
    double s1 = 0.0, s2 = 0.0, mn = DBL_MAX, mx = -DBL_MAX;
#pragma unroll
    for (int t = 0; t < LMAX; ++t)
        if (t < len) {
            double v = y[t];
            s1 += v;
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
    double S1, dummy, MN, MX;
    block_red2d<NW, 1>(red, phase, s1, 0.0, S1, dummy);
    block_red2d<NW, 0>(red, phase, mn, mx, MN, MX);
    if (!(MN < MX)) return MN;  // all equal (or n == 1)
    const double mean = S1 / (double)n;
#pragma unroll
    for (int t = 0; t < LMAX; ++t)
        if (t < len) {
            double d = y[t] - mean;
            s2 += d * d;
        }
    double S2;
    block_red2d<NW, 1>(red, phase, s2, 0.0, S2, dummy);
    const double sd = sqrt(S2 / (double)n);

    // invariant: #(x <= lo) <= kA  and  #(x <= hi) >= kB + 1.  lo starts one ulp below the minimum.
    double lo = double_of_key(key_of(MN) - 1ull), hi = MX;
    if (!(lo < MN)) lo = double_of_key(key_of(MN) - 2ull);  // MN == +0.0: one key below is -0.0 == MN
    int Flo = 0, Fhi = n;
    double p1 = mean - 0.08 * sd, p2 = mean + 0.08 * sd;
    bool force_bisect = false;
    double a_res = 0.0, b_res = 0.0;
    bool done = false;

    for (int round = 0; round < 160 && !done; ++round) {
        int m = Fhi - Flo;
        if (m <= CAND_MAX) break;
        // ---- choose pivots strictly inside (lo, hi) ------------------------------------------
        const double lo_eff = lo;
        unsigned long long klo = key_of(lo), khi = key_of(hi);
        if (khi - klo < 2ull) {  // no double strictly between: every candidate equals hi
            a_res = b_res = hi;
            done = true;
            break;
        }
        double pmid = double_of_key(klo + ((khi - klo) >> 1));
        if (round > 0) {
            if (force_bisect) {
                p1 = p2 = pmid;
            } else {
                double f = ((double)kA + 0.5 * (double)(kB - kA) + 0.5 - (double)Flo) / (double)m;
                double wfrac = (3.0 * sqrt((double)m) + 8.0) / (double)m;
                if (wfrac > 0.5) wfrac = 0.5;
                double span = hi - lo_eff;
                double pc = lo_eff + span * f;
                p1 = pc - 0.5 * span * wfrac;
                p2 = pc + 0.5 * span * wfrac;
            }
        }
        if (!(p1 > lo && p1 < hi)) p1 = pmid;
        if (!(p2 > lo && p2 < hi)) p2 = pmid;
        if (p1 > p2) {
            double t = p1;
            p1 = p2;
            p2 = t;
        }
        // ---- count ------------------------------------------------------------------------------
        int c1 = 0, c2 = 0;
#pragma unroll
        for (int t = 0; t < LMAX; ++t)
            if (t < len) {
                c1 += (y[t] <= p1) ? 1 : 0;
                c2 += (y[t] <= p2) ? 1 : 0;
            }
        int C1, C2;
        block_sum2i<NW>(red, phase, c1, c2, C1, C2);
        // ---- narrow -----------------------------------------------------------------------------
        double split = 0.0;
        bool do_split = false;
        if (C1 >= kB + 1) {
            hi = p1;
            Fhi = C1;
        } else if (C1 > kA) {
            split = p1;
            do_split = true;
        } else if (C2 >= kB + 1) {
            lo = p1;
            Flo = C1;
            hi = p2;
            Fhi = C2;
        } else if (C2 > kA) {
            split = p2;
            do_split = true;
        } else {
            lo = p2;
            Flo = C2;
        }
        if (do_split) {  // s_kA <= split < s_kB: neighbours of the split point
            double below = -DBL_MAX, above = DBL_MAX;
#pragma unroll
            for (int t = 0; t < LMAX; ++t)
                if (t < len) {
                    double v = y[t];
                    if (v <= split) below = fmax(below, v);
                    else above = fmin(above, v);
                }
            block_red2d<NW, 2>(red, phase, below, above, a_res, b_res);
            done = true;
            break;
        }
        int m_new = Fhi - Flo;
        force_bisect = (2 * m_new > m) && !force_bisect;
    }
    if (done) return (a_res + b_res) * 0.5;

    // ---- gather the <= CAND_MAX candidates in (lo, hi] and rank them --------------------------------
    if (threadIdx.x == 0) *cand_n = 0;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < LMAX; ++t)
        if (t < len) {
            double v = y[t];
            if (v > lo && v <= hi) {
                int slot = atomicAdd(cand_n, 1);
                if (slot < CAND_MAX) cand[slot] = v;
            }
        }
    __syncthreads();
    int m = *cand_n;
    if (m > CAND_MAX) m = CAND_MAX;  // cannot happen (m == Fhi - Flo); keeps the loop bounded
    int ra = kA - Flo, rb = kB - Flo;
    if ((int)threadIdx.x < m) {
        double v = cand[threadIdx.x];
        int rank = 0;
        for (int j = 0; j < m; ++j) {
            double u = cand[j];
            rank += (u < v || (u == v && j < (int)threadIdx.x)) ? 1 : 0;
        }
        if (rank == ra) cand[CAND_MAX] = v;
        if (rank == rb) cand[CAND_MAX + 1] = v;
    }
    __syncthreads();
    double a = cand[CAND_MAX], b = cand[CAND_MAX + 1];
    __syncthreads();  // cand is reused by the next cell
    return (a + b) * 0.5;
}

// dead-band subtraction, .subtract_expr (ops.R:1764-1769): strict inequalities
__device__ __forceinline__ double sub_bounds(double x, double lo, double hi) {
    return (x > hi) ? (x - hi) : ((x < lo) ? (x - lo) : 0.0);
}

template <int NT>
__global__ void __launch_bounds__(NT, (NT == 256) ? 2 : 1) cell_pipeline_kernel(const CellParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    double *s = reinterpret_cast<double *>(smem_raw);       // the cell's gene vector
    double *invD = s + p.s_elems;                            // 1/D for one-sided truncation, h+1 entries
    double *cand = invD + (p.h + 2);                         // CAND_MAX + 2
    Red<NW> &red = *reinterpret_cast<Red<NW> *>(cand + CAND_MAX + 2);
    int *cand_n = reinterpret_cast<int *>(&red + 1);

    const int tid = threadIdx.x;
    const int G = (int)p.G;
    const int h = p.h;
    const bool do_smooth = p.window >= 2;
    int phase = 0;

    if (do_smooth) {
        double full = (double)(h + 1) * (double)(h + 1);
        for (int r = tid; r <= h; r += NT) invD[r] = 1.0 / (full - 0.5 * (double)r * (double)(r + 1));
    }
    const Seg seg = p.segs[tid];
    bool bad = false;

    for (int64_t ci = blockIdx.x; ci < p.n_cols; ci += gridDim.x) {
        const int64_t col = p.cols ? (int64_t)p.cols[ci] : ci;
        const double *__restrict__ src = p.X + p.ldx * col;
        double *__restrict__ dst = p.Y + p.ldy * ci;

        // ---- stage A: one coalesced read of the column, element-wise steps fused into it --------
        for (int g0 = tid; g0 < G; g0 += 4 * NT) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int g = g0 + u * NT;
                v[u] = (g < G) ? src[g] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int g = g0 + u * NT;
                if (g < G) {
                    double x = v[u];
                    if (!is_finite_d(x)) bad = true;
                    if (p.apply_log) x = log2(x + 1.0);
                    if (p.lo1) x = sub_bounds(x, p.lo1[g], p.hi1[g]);
                    else if (p.mid1) x = x - p.mid1[g];
                    if (p.threshold > 0.0) x = fmin(fmax(x, -p.threshold), p.threshold);
                    s[g] = x;
                }
            }
        }
        __syncthreads();

        // ---- stage B: pyramid smooth of this thread's segment, results stay in registers ---------
        double y[LMAX];
        {
            const int a = seg.start, len = seg.len, cs = seg.cs, ce = seg.ce;
            if (len > 0 && do_smooth && (ce - cs) >= 2) {
                // weighted window sum N(a) and the two half-window box sums, x = 0 outside [cs, ce)
                double N = 0.0, Ls = 0.0, Rs = 0.0;
                int jlo = max(cs, a - h), jhi = min(ce - 1, a + h);
                for (int j = jlo; j <= jhi; ++j) {
                    double v = s[j];
                    int d = j - a;
                    N = fma((double)(h + 1 - (d < 0 ? -d : d)), v, N);
                    if (d <= 0) Ls += v;
                    else Rs += v;
                }
                if (a + h + 1 < ce) Rs += s[a + h + 1];
#pragma unroll
                for (int t = 0; t < LMAX; ++t) {
                    if (t < len) {
                        int i = a + t;
                        int rl = h - (i - cs);
                        rl = rl > 0 ? rl : 0;
                        int rr = h - (ce - 1 - i);
                        rr = rr > 0 ? rr : 0;
                        double out;
                        if (rl == 0 || rr == 0) {
                            out = N * invD[rl + rr];
                        } else {  // chromosome shorter than the window: both ends truncated
                            double D = (double)(h + 1) * (double)(h + 1) - 0.5 * (double)rl * (double)(rl + 1) -
                                       0.5 * (double)rr * (double)(rr + 1);
                            out = N / D;
                        }
                        y[t] = out;
                        // slide: weights of x[i+1 .. i+h+1] grow by one, those of x[i-h .. i] shrink by one
                        N += (Rs - Ls);
                        double xin = (i + 1 < ce) ? s[i + 1] : 0.0;
                        double xoutL = (i - h >= cs) ? s[i - h] : 0.0;
                        double xinR = (i + h + 2 < ce) ? s[i + h + 2] : 0.0;
                        Ls += xin - xoutL;
                        Rs += xinR - xin;
                    } else {
                        y[t] = 0.0;
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < LMAX; ++t) y[t] = (t < len) ? s[a + t] : 0.0;
            }
        }

        // ---- stage C: per-cell centre over all genes ---------------------------------------------------
        double centre = 0.0;
        if (p.center == 1) {
            centre = block_median<NT>(y, seg.len, G, red, phase, cand, cand_n);
        } else if (p.center == 2) {
            double s1 = 0.0;
#pragma unroll
            for (int t = 0; t < LMAX; ++t)
                if (t < seg.len) s1 += y[t];
            double S1, dummy;
            block_red2d<NW, 1>(red, phase, s1, 0.0, S1, dummy);
            centre = S1 / (double)G;
        }
        __syncthreads();  // every thread is done reading s[] for the smooth
#pragma unroll
        for (int t = 0; t < LMAX; ++t)
            if (t < seg.len) s[seg.start + t] = y[t] - centre;
        __syncthreads();

        // ---- stage D: second reference subtraction + 2^x fused into the one coalesced write -------------
        for (int g = tid; g < G; g += NT) {
            double x = s[g];
            if (p.lo2) x = sub_bounds(x, p.lo2[g], p.hi2[g]);
            else if (p.mid2) x = x - p.mid2[g];
            if (p.apply_exp2) x = exp2(x);
            dst[g] = x;
        }
        __syncthreads();  // s[] is overwritten by the next cell
    }
    if (bad && p.err_flag) atomicExch(p.err_flag, 1);
}

// =================================================================================================
// host-side launchers (device-pointer ABI)
// =================================================================================================

static int build_segments(int64_t G, const int32_t *chr_start, const int32_t *chr_len, int K, int NT,
                          std::vector<Seg> &segs) {
    // cover every gene exactly once; genes outside every chromosome range are an argument error
    int64_t covered = 0;
    for (int k = 0; k < K; ++k) {
        if (chr_len[k] < 0 || chr_start[k] < 0 || (int64_t)chr_start[k] + chr_len[k] > G) return -1;
        if (k > 0 && chr_start[k] != chr_start[k - 1] + chr_len[k - 1]) return -1;
        covered += chr_len[k];
    }
    if (K > 0 && chr_start[0] != 0) return -1;
    if (covered != G) return -1;
    // L odd: adjacent threads then start an odd number of 8-byte words apart, so the strided
    // 64-bit shared-memory accesses of the smooth are bank-conflict free within a half-warp.
    for (int L = 1; L <= LMAX; L += 2) {
        int64_t n = 0;
        for (int k = 0; k < K; ++k) n += (chr_len[k] + L - 1) / L;
        if (n <= NT) {
            segs.assign(NT, Seg{0, 0, 0, 0});
            int t = 0;
            for (int k = 0; k < K; ++k) {
                int pos = chr_start[k], end = chr_start[k] + chr_len[k];
                while (pos < end) {
                    int len = end - pos < L ? end - pos : L;
                    segs[t++] = Seg{pos, len, chr_start[k], end};
                    pos += len;
                }
            }
            return L;
        }
    }
    return 0;  // does not fit
}

}  // namespace icnv

using namespace icnv;

extern "C" {

int icnv_dev_group_partial_sums_f64(const double *X, int64_t G, int64_t ldx, const int32_t *cells, int64_t n_cells,
                                    int chunk, int apply_log, double *partial, void *stream) {
    ICNV_REQUIRE_READY();
    if (!X || !cells || !partial || G <= 0 || n_cells <= 0 || chunk <= 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_dev_group_partial_sums_f64: bad argument");
    int64_t n_chunks = (n_cells + chunk - 1) / chunk;
    if (n_chunks > 65535) return set_error(ICNV_E_BAD_ARG, "too many chunks (%lld)", (long long)n_chunks);
    dim3 grid((unsigned)((G + 255) / 256), (unsigned)n_chunks);
    if (apply_log == 2)
        group_partial_sums_invlog_kernel<<<grid, 256, 0, pick_stream(stream)>>>(X, G, ldx, cells, n_cells, chunk,
                                                                              partial);
    else
        group_partial_sums_kernel<<<grid, 256, 0, pick_stream(stream)>>>(X, G, ldx, cells, n_cells, chunk, apply_log,
                                                                       partial);
    ICNV_CHECK_LAUNCH("group_partial_sums_kernel");
    return ICNV_OK;
}

int icnv_dev_combine_partials_f64(const double *partial, int64_t G, int64_t n_chunks, int64_t count, double *means,
                                  void *stream) {
    ICNV_REQUIRE_READY();
    if (!partial || !means || G <= 0 || n_chunks <= 0 || count <= 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_dev_combine_partials_f64: bad argument");
    combine_partials_kernel<<<(unsigned)((G + 255) / 256), 256, 0, pick_stream(stream)>>>(partial, G, n_chunks,
                                                                                          (double)count, means);
    ICNV_CHECK_LAUNCH("combine_partials_kernel");
    return ICNV_OK;
}

int icnv_dev_bounds_from_means_f64(const double *means, int64_t G, int n_grp, double *lo, double *hi, double *mid,
                                   void *stream) {
    ICNV_REQUIRE_READY();
    if (!means || !lo || !hi || G <= 0 || n_grp <= 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_dev_bounds_from_means_f64: bad argument");
    bounds_from_means_kernel<<<(unsigned)((G + 255) / 256), 256, 0, pick_stream(stream)>>>(means, G, n_grp, lo, hi, mid);
    ICNV_CHECK_LAUNCH("bounds_from_means_kernel");
    return ICNV_OK;
}

int icnv_dev_invlog_finish_f64(double *means, int64_t n, void *stream) {
    ICNV_REQUIRE_READY();
    invlog_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pick_stream(stream)>>>(means, n);
    ICNV_CHECK_LAUNCH("invlog_finish_kernel");
    return ICNV_OK;
}

int icnv_dev_cell_pipeline_f64(const double *X, int64_t G, int64_t ldx, const int32_t *cols, int64_t n_cols, double *Y,
                               int64_t ldy, const int32_t *chr_start, const int32_t *chr_len, int K, int apply_log,
                               const double *lo1, const double *hi1, const double *mid1, double threshold, int window,
                               int center, const double *lo2, const double *hi2, const double *mid2, int apply_exp2,
                               int *err_flag, void *stream) {
    ICNV_REQUIRE_READY();
    Ctx &c = ctx();
    if (!X || !Y || G <= 0 || n_cols < 0 || ldx < G || ldy < G || !chr_start || !chr_len || K <= 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_dev_cell_pipeline_f64: bad argument");
    if ((lo1 == nullptr) != (hi1 == nullptr) || (lo2 == nullptr) != (hi2 == nullptr))
        return set_error(ICNV_E_BAD_ARG, "lo/hi bounds must be given in pairs");
    if (window >= 2 && (window & 1) == 0)
        return set_error(ICNV_E_BAD_ARG, "window_length %d is even: the reference's behaviour is accidental there", window);
    if (center < 0 || center > 2) return set_error(ICNV_E_BAD_ARG, "center must be 0, 1 or 2");
    if (n_cols == 0) return ICNV_OK;
    if (G > (int64_t)512 * LMAX) return set_error(ICNV_E_UNSUPPORTED, "G = %lld exceeds %d genes", (long long)G, 512 * LMAX);

    int h = window >= 2 ? (window - 1) / 2 : 0;
    int s_elems = (int)((G + 1) & ~(int64_t)1);
    std::vector<Seg> segs;
    int NT = 256;
    int L = build_segments(G, chr_start, chr_len, K, NT, segs);
    if (L < 0) return set_error(ICNV_E_BAD_ARG, "chromosome ranges must tile [0, G) contiguously");
    if (L == 0) {
        NT = 512;
        L = build_segments(G, chr_start, chr_len, K, NT, segs);
        if (L <= 0) return set_error(ICNV_E_UNSUPPORTED, "G = %lld with K = %d does not fit 512 x %d", (long long)G, K, LMAX);
    }
    size_t red_bytes = (NT == 256) ? sizeof(Red<8>) : sizeof(Red<16>);
    size_t smem = sizeof(double) * ((size_t)s_elems + (size_t)(h + 2) + CAND_MAX + 2) + red_bytes + 16;
    if (smem > (size_t)c.smem_optin)
        return set_error(ICNV_E_UNSUPPORTED, "needs %zu B shared memory per CTA, device allows %d", smem, c.smem_optin);

    cudaStream_t st = pick_stream(stream);
    Seg *d_segs = (Seg *)scratch(SLOT_SEGS, sizeof(Seg) * 512);
    if (!d_segs) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemcpyAsync(d_segs, segs.data(), sizeof(Seg) * NT, cudaMemcpyHostToDevice, st));

    CellParams p;
    p.X = X; p.G = G; p.ldx = ldx; p.cols = cols; p.n_cols = n_cols; p.Y = Y; p.ldy = ldy; p.segs = d_segs;
    p.apply_log = apply_log; p.lo1 = lo1; p.hi1 = hi1; p.mid1 = mid1; p.threshold = threshold;
    p.window = window; p.h = h; p.center = center; p.lo2 = lo2; p.hi2 = hi2; p.mid2 = mid2;
    p.apply_exp2 = apply_exp2; p.err_flag = err_flag; p.s_elems = s_elems;

    int per_sm = 1;
    if (NT == 256) {
        ICNV_CUDA(cudaFuncSetAttribute(cell_pipeline_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ICNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cell_pipeline_kernel<256>, 256, smem));
    } else {
        ICNV_CUDA(cudaFuncSetAttribute(cell_pipeline_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ICNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cell_pipeline_kernel<512>, 512, smem));
    }
    if (per_sm < 1) return set_error(ICNV_E_UNSUPPORTED, "cell_pipeline_kernel does not fit on an SM");
    int64_t grid = (int64_t)c.sm_count * per_sm;  // persistent CTAs: a whole number of waves
    if (grid > n_cols) grid = n_cols;
    if (NT == 256) cell_pipeline_kernel<256><<<(unsigned)grid, 256, smem, st>>>(p);
    else cell_pipeline_kernel<512><<<(unsigned)grid, 512, smem, st>>>(p);
    ICNV_CHECK_LAUNCH("cell_pipeline_kernel");
    return ICNV_OK;
}

}  // extern "C"
