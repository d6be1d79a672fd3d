// icnv_median_filter.cu - K4: apply_median_filtering / .median_filter (R/noise_reduction.R:43-113).
//
// out[i, j] = median(B[max(0,i-r) .. min(n-1,i+r), max(0,j-r) .. min(m-1,j+r)]) on every block
// B = (genes of one chromosome) x (cells of one index list, in list order), r = (window_size+1)/2
// as in the reference (noise_reduction.R:102-106: half_window + 1).  Even-count windows average
// the two middle values (median.default).  Reads the un-filtered input throughout.
//
// One thread per output element.  Window sizes 3..9 (radius 2..5) run median_filter_select_kernel: the tile's halo is
// staged in shared memory and every thread selects its window's median by counting (icnv_median_select.cuh - no
// data-dependent partitioning, so the lanes of a warp stay together).  Other window sizes use the generic kernel
// below (in-place Wirth selection over a private index array).
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "icnv_common.cuh"
#include "icnv_median_select.cuh"

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
    int use_net;              // radius 4 only: full windows by the key network (window_median_net81), 0 = counting selection
};

// One CTA = a tile of TI genes x TJ list positions (TI*TJ threads, one output each).  The tile's halo
// ((TI+2r) x (TJ+2r) values, clamped to the chromosome / index-list block) is loaded ONCE into shared
// memory, coalesced along genes; every thread then selects its median over a private array of 16-bit
// indices into that shared halo - nothing but indices ever moves.  12 KB of shared memory per 64-thread
// CTA for the default window (against 83 KB of private value copies before) keeps ~36 warps per SM busy
// on what is a chain of dependent shared-memory reads.
constexpr int MF_TI = 8, MF_TJ = 8, MF_NT = MF_TI * MF_TJ;

__global__ void __launch_bounds__(MF_NT) median_filter_kernel(const MfParams p) {
    extern __shared__ __align__(16) unsigned char mf_smem[];
    const int r = p.r;
    const int HR = MF_TI + 2 * r, HC = MF_TJ + 2 * r;       // halo rows (genes, fast) x cols (cells)
    const int W = (2 * r + 1) * (2 * r + 1);
    double *halo = reinterpret_cast<double *>(mf_smem);
    unsigned short *idx_all = reinterpret_cast<unsigned short *>(halo + HR * HC);
    const MfTile gt = p.gene_tiles[blockIdx.y];
    const MfTile ct = p.cell_tiles[blockIdx.x];
    const int hi0 = gt.start - r, hj0 = ct.start - r;
    bool bad = false;
    // ---- halo: rows inside [gt.lo, gt.hi), list positions inside [ct.lo, ct.hi) -------------------------
    for (int e = threadIdx.x; e < HR * HC; e += MF_NT) {
        const int hr = e % HR, hc = e / HR;
        const int ii = hi0 + hr, jj = hj0 + hc;
        double v = 0.0;
        if (ii >= gt.lo && ii < gt.hi && jj >= ct.lo && jj < ct.hi) {
            v = p.X[ii + p.G * (int64_t)p.cells[jj]];
            bad |= !is_finite_d(v);
        }
        halo[e] = v;
    }
    __syncthreads();
    const int ti = threadIdx.x % MF_TI, tj = threadIdx.x / MF_TI;
    if (ti < gt.len && tj < ct.len) {
        const int i = gt.start + ti, j = ct.start + tj;
        const int xa = max(gt.lo, i - r), xb = min(gt.hi - 1, i + r);
        const int ya = max(ct.lo, j - r), yb = min(ct.hi - 1, j + r);
        unsigned short *a = idx_all + threadIdx.x;   // [tap][thread]
        int n = 0;
        for (int jj = ya; jj <= yb; ++jj)
            for (int ii = xa; ii <= xb; ++ii) {
                a[n * MF_NT] = (unsigned short)((jj - hj0) * HR + (ii - hi0));
                ++n;
            }
        (void)W;
#define MF_VAL(q) halo[a[(q) * MF_NT]]
        // k-th smallest, k = (n-1)/2 (Wirth); afterwards val[0..k-1] <= val[k] <= val[k+1..n-1]
        const int k = (n - 1) >> 1;
        int l = 0, rr = n - 1;
        while (l < rr) {
            const double x = MF_VAL(k);
            int u = l, w = rr;
            do {
                while (MF_VAL(u) < x) ++u;
                while (x < MF_VAL(w)) --w;
                if (u <= w) {
                    const unsigned short t = a[u * MF_NT];
                    a[u * MF_NT] = a[w * MF_NT];
                    a[w * MF_NT] = t;
                    ++u;
                    --w;
                }
            } while (u <= w);
            if (w < k) l = u;
            if (k < u) rr = w;
        }
        double med = MF_VAL(k);
        if ((n & 1) == 0) {  // mean of the two middle values
            double nxt = DBL_MAX;
            for (int q = k + 1; q < n; ++q) nxt = fmin(nxt, MF_VAL(q));
            med = (med + nxt) * 0.5;
        }
#undef MF_VAL
        p.Y[i + p.G * (int64_t)p.cells[j]] = med;
    }
    if (bad && p.err_flag) atomicExch(p.err_flag, 1);
}

// Tile of 32 genes x 8 list positions, 256 threads; a warp covers 32 consecutive genes of one cell, so its halo reads
// are consecutive doubles (no bank conflicts) and the per-thread tap lists interleave at 2-byte granularity.
constexpr int MS_TI = 32, MS_TJ = 8, MS_NT = MS_TI * MS_TJ;

// NET (radius 4 only): full windows go through the key network (window_median_net81, ~160 registers per thread); a separate
// instance, so that the counting-selection kernel keeps its 64 registers and four CTAs per SM.
template <int R, typename ListT = unsigned short, bool NET = false>
__global__ void __launch_bounds__(MS_NT, NET ? 2 : 0) median_filter_select_kernel(const MfParams p) {
    extern __shared__ __align__(16) unsigned char mf_smem[];
    constexpr int D = 2 * R + 1, HR = MS_TI + 2 * R, HC = MS_TJ + 2 * R;
    double *halo = reinterpret_cast<double *>(mf_smem);          // +inf outside the block: never counted
    double *halo0 = halo + HR * HC;                              // 0 outside the block: for the window moments
    ListT *list = reinterpret_cast<ListT *>(halo0 + HR * HC);   // [D*D][MS_NT] tap offsets
    float *kf = reinterpret_cast<float *>(list + (size_t)D * D * MS_NT);   // use_net: single-precision keys of the halo
    const MfTile gt = p.gene_tiles[blockIdx.y];
    const MfTile ct = p.cell_tiles[blockIdx.x];
    const int hi0 = gt.start - R, hj0 = ct.start - R;
    bool bad = false;
    for (int e = threadIdx.x; e < HR * HC; e += MS_NT) {
        const int hr = e % HR, hc = e / HR;
        const int ii = hi0 + hr, jj = hj0 + hc;
        double v = INFINITY, v0 = 0.0;
        if (ii >= gt.lo && ii < gt.hi && jj >= ct.lo && jj < ct.hi) {
            v = p.X[ii + p.G * (int64_t)p.cells[jj]];
            bad |= !is_finite_d(v);
            v0 = v;
        }
        halo[e] = v;
        halo0[e] = v0;
        if (NET) kf[e] = (float)v;
    }
    __syncthreads();
    const int ti = threadIdx.x % MS_TI, tj = threadIdx.x / MS_TI;
    // a tile whose every window is a full (2R+1)^2 one (no tap outside the chromosome / index-list block): CTA-uniform
    const bool full_tile = gt.len == MS_TI && ct.len == MS_TJ && gt.start - R >= gt.lo && gt.start + MS_TI + R <= gt.hi &&
                           ct.start - R >= ct.lo && ct.start + MS_TJ + R <= ct.hi;
    if (ti < gt.len && tj < ct.len) {
        const int i = gt.start + ti, j = ct.start + tj;
        double med;
        if (NET && full_tile) {
            med = window_median_net81<ListT>(halo, halo0, kf, HR, tj * HR + ti, list + threadIdx.x, MS_NT);
        } else {
            const int xa = max(gt.lo, i - R), xb = min(gt.hi - 1, i + R);
            const int ya = max(ct.lo, j - R), yb = min(ct.hi - 1, j + R);
            const int n = (xb - xa + 1) * (yb - ya + 1);
            med = window_median<R, ListT>(halo, halo0, HR, tj * HR + ti, list + threadIdx.x, MS_NT, n);
        }
        p.Y[i + p.G * (int64_t)p.cells[j]] = med;
    }
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
    bool select_kernel = (r >= 2 && r <= 5);
    int use_net = 0;
    if (c.opt_mf_kernel >= 0) {   // ICNV_MF_KERNEL, read once in icnv_init: 0 generic kernel; 2 key network for full 9 x 9 windows (A/B runs)
        select_kernel = select_kernel && (c.opt_mf_kernel != 0);
        use_net = (c.opt_mf_kernel == 2 && r == 4) ? 1 : 0;
    }
    const int TI = select_kernel ? MS_TI : MF_TI, TJ = select_kernel ? MS_TJ : MF_TJ, NT = TI * TJ;
    bool list32 = false;
    if (c.opt_mf_list32) list32 = select_kernel;   // diagnostic (ICNV_MF_LIST32 at icnv_init)
    const size_t smem = sizeof(double) * (size_t)(TI + 2 * r) * (size_t)(TJ + 2 * r) * (select_kernel ? 2 : 1) +
                        (list32 ? sizeof(unsigned) : sizeof(unsigned short)) * (size_t)W * (size_t)NT +
                        (use_net ? sizeof(float) * (size_t)(TI + 2 * r) * (size_t)(TJ + 2 * r) : 0);
    if (smem > (size_t)c.smem_optin || (TI + 2 * r) * (TJ + 2 * r) > 65535)
        return set_error(ICNV_E_UNSUPPORTED, "window_size %d needs %zu B of shared memory per CTA", window_size, smem);
    cudaStream_t st = pick_stream(stream);
    // cells in no list are copied through
    ICNV_CUDA(cudaMemcpyAsync(Y, X, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToDevice, st));
    if (n_grp == 0) return ICNV_OK;
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
    MfParams p{X, Y, G, d_cells, d_gt, d_ct, r, d_flag, use_net};
    dim3 grid((unsigned)ct.size(), (unsigned)gt.size());
    auto launch = [&](auto kern) -> int {
        ICNV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NT, smem, st>>>(p);
        return ICNV_OK;
    };
    int lrc;
    if (list32) use_net = 0;
    if (!select_kernel) lrc = launch(median_filter_kernel);
    else if (list32 && r == 5) lrc = launch(median_filter_select_kernel<5, unsigned>);
    else if (list32 && r == 4) lrc = launch(median_filter_select_kernel<4, unsigned>);
    else if (r == 2) lrc = launch(median_filter_select_kernel<2>);
    else if (r == 3) lrc = launch(median_filter_select_kernel<3>);
    else if (r == 4 && use_net) lrc = launch(median_filter_select_kernel<4, unsigned short, true>);
    else if (r == 4) lrc = launch(median_filter_select_kernel<4>);
    else lrc = launch(median_filter_select_kernel<5>);
    if (lrc) return lrc;
    ICNV_CHECK_LAUNCH("median_filter_kernel");
    return ICNV_OK;
}
