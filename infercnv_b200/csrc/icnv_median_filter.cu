// icnv_median_filter.cu - K4: apply_median_filtering / .median_filter (R/noise_reduction.R:43-113).
//
// out[i, j] = median(B[max(0,i-r) .. min(n-1,i+r), max(0,j-r) .. min(m-1,j+r)]) on every block
// B = (genes of one chromosome) x (cells of one index list, in list order), r = (window_size+1)/2
// as in the reference (noise_reduction.R:102-106: half_window + 1).  Even-count windows average
// the two middle values (median.default).  Reads the un-filtered input throughout.
//
// One thread per output element.  A CTA covers a tile of TI consecutive genes x TJ consecutive
// list positions; lanes run along genes, so each (di, dj) tap of the window is one coalesced
// 128-byte read per 16 genes and neighbouring taps hit L1/L2.  The up-to (2r+1)^2 window values of
// each thread are staged in a thread-private shared-memory column ([tap][thread], conflict free)
// and the middle order statistics come from an in-place k-th smallest selection (Wirth/Hoare).
#include <cfloat>
#include <cmath>
#include <vector>

#include "icnv_common.cuh"

namespace icnv {

struct MfTile {
    int start;  // first gene (or first list position) of the tile
    int len;    // valid entries in the tile
    int lo;     // block start (chromosome start / group offset in the concatenated list)
    int hi;     // block end, exclusive
};

struct MfParams {
    const double *X;
    double *Y;
    int64_t G;
    const int32_t *cells;     // concatenated index lists (device)
    const MfTile *gene_tiles; // gridDim.y entries
    const MfTile *cell_tiles; // gridDim.x entries
    int r;
    int *err_flag;
};

template <int NT, int TI>
__global__ void __launch_bounds__(NT) median_filter_kernel(const MfParams p) {
    extern __shared__ __align__(16) double win[];  // [tap][NT]
    const MfTile gt = p.gene_tiles[blockIdx.y];
    const MfTile ct = p.cell_tiles[blockIdx.x];
    const int ti = threadIdx.x % TI, tj = threadIdx.x / TI;
    if (ti >= gt.len || tj >= ct.len) return;
    const int i = gt.start + ti, j = ct.start + tj;  // gene row, position in the concatenated list
    const int r = p.r;
    const int xa = max(gt.lo, i - r), xb = min(gt.hi - 1, i + r);
    const int ya = max(ct.lo, j - r), yb = min(ct.hi - 1, j + r);
    double *a = win + threadIdx.x;
    int n = 0;
    bool bad = false;
    for (int jj = ya; jj <= yb; ++jj) {
        const double *col = p.X + p.G * (int64_t)p.cells[jj];
        for (int ii = xa; ii <= xb; ++ii) {
            double v = col[ii];
            bad |= !is_finite_d(v);
            a[(int64_t)n * NT] = v;
            ++n;
        }
    }
    // k-th smallest, k = (n-1)/2 (Wirth); afterwards a[0..k-1] <= a[k] <= a[k+1..n-1]
    const int k = (n - 1) >> 1;
    int l = 0, rr = n - 1;
    while (l < rr) {
        const double x = a[(int64_t)k * NT];
        int u = l, w = rr;
        do {
            while (a[(int64_t)u * NT] < x) ++u;
            while (x < a[(int64_t)w * NT]) --w;
            if (u <= w) {
                double t = a[(int64_t)u * NT];
                a[(int64_t)u * NT] = a[(int64_t)w * NT];
                a[(int64_t)w * NT] = t;
                ++u;
                --w;
            }
        } while (u <= w);
        if (w < k) l = u;
        if (k < u) rr = w;
    }
    double med = a[(int64_t)k * NT];
    if ((n & 1) == 0) {  // mean of the two middle values
        double nxt = DBL_MAX;
        for (int q = k + 1; q < n; ++q) nxt = fmin(nxt, a[(int64_t)q * NT]);
        med = (med + nxt) * 0.5;
    }
    p.Y[i + p.G * (int64_t)p.cells[j]] = med;
    if (bad && p.err_flag) atomicExch(p.err_flag, 1);
}

static void make_tiles(const int32_t *start, const int32_t *len, int nblk, int T, std::vector<MfTile> &out) {
    for (int b = 0; b < nblk; ++b) {
        int lo = start[b], hi = start[b] + len[b];
        for (int s = lo; s < hi; s += T) out.push_back(MfTile{s, (hi - s < T ? hi - s : T), lo, hi});
    }
}

}  // namespace icnv

using namespace icnv;

extern "C" int icnv_dev_median_filter_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                                          const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx,
                                          int n_grp, int window_size, void *stream) {
    ICNV_REQUIRE_READY();
    Ctx &c = ctx();
    if (!X || !Y || X == Y || G <= 0 || C <= 0 || !chr_start || !chr_len || K <= 0 || n_grp < 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_dev_median_filter_f64: bad argument (X and Y must not alias)");
    if (window_size < 3 || (window_size & 1) == 0)
        return set_error(ICNV_E_BAD_ARG, "window_size must be an odd number >= 3 (noise_reduction.R:48-50)");
    const int r = (window_size + 1) / 2;
    const int W = (2 * r + 1) * (2 * r + 1);
    int NT = 128;
    if ((size_t)W * NT * sizeof(double) > (size_t)c.smem_optin) NT = 64;
    if ((size_t)W * NT * sizeof(double) > (size_t)c.smem_optin)
        return set_error(ICNV_E_UNSUPPORTED, "window_size %d needs %zu B of shared memory per CTA", window_size,
                         (size_t)W * NT * sizeof(double));
    cudaStream_t st = pick_stream(stream);
    // cells in no list are copied through
    ICNV_CUDA(cudaMemcpyAsync(Y, X, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToDevice, st));
    if (n_grp == 0) return ICNV_OK;
    const int TI = 16, TJ = NT / TI;
    std::vector<MfTile> gt, ct;
    make_tiles(chr_start, chr_len, K, TI, gt);
    std::vector<int32_t> g_start(n_grp), g_len(n_grp);
    for (int b = 0; b < n_grp; ++b) {
        g_start[b] = grp_off[b];
        g_len[b] = grp_off[b + 1] - grp_off[b];
    }
    make_tiles(g_start.data(), g_len.data(), n_grp, TJ, ct);
    if (gt.empty() || ct.empty()) return ICNV_OK;
    if (gt.size() > 65535) return set_error(ICNV_E_UNSUPPORTED, "too many gene tiles (%zu)", gt.size());
    const int64_t n_idx = grp_off[n_grp];
    size_t bytes = sizeof(MfTile) * (gt.size() + ct.size()) + sizeof(int32_t) * (size_t)n_idx + 64;
    char *d = (char *)scratch(SLOT_MF, bytes);
    int *d_flag = (int *)scratch(SLOT_MISC, 64);
    if (!d || !d_flag) return ICNV_E_NOMEM;
    MfTile *d_gt = (MfTile *)d;
    MfTile *d_ct = d_gt + gt.size();
    int32_t *d_cells = (int32_t *)(d_ct + ct.size());
    ICNV_CUDA(cudaMemcpyAsync(d_gt, gt.data(), sizeof(MfTile) * gt.size(), cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaMemcpyAsync(d_ct, ct.data(), sizeof(MfTile) * ct.size(), cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaMemcpyAsync(d_cells, grp_idx, sizeof(int32_t) * (size_t)n_idx, cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    ICNV_CUDA(cudaStreamSynchronize(st));  // tile tables are stack-lifetime host buffers
    MfParams p{X, Y, G, d_cells, d_gt, d_ct, r, d_flag};
    size_t smem = (size_t)W * NT * sizeof(double);
    dim3 grid((unsigned)ct.size(), (unsigned)gt.size());
    if (NT == 128) {
        ICNV_CUDA(cudaFuncSetAttribute(median_filter_kernel<128, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        median_filter_kernel<128, 16><<<grid, 128, smem, st>>>(p);
    } else {
        ICNV_CUDA(cudaFuncSetAttribute(median_filter_kernel<64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        median_filter_kernel<64, 16><<<grid, 64, smem, st>>>(p);
    }
    ICNV_CHECK_LAUNCH("median_filter_kernel");
    return ICNV_OK;
}
