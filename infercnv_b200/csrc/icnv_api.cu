// icnv_api.cu - C ABI of libinfercnv_b200.so: lifecycle, the host-pointer entry points the R shim
// binds, and the device-side composition of the smooth block.  See include/infercnv_b200.h.
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "icnv_common.cuh"

// device-pointer helpers defined in the other translation units
extern "C" {
int icnv_dev_invlog_finish_f64(double *means, int64_t n, void *stream);
int icnv_dev_widen_states(const uint8_t *s, int32_t *out, int64_t n, void *stream);
int icnv_dev_narrow_states(const int32_t *s, uint8_t *out, int64_t n, void *stream);
int icnv_dev_scatter_group_states(const uint8_t *gs, int64_t G, int64_t C, const int32_t *grp_of, int32_t *out,
                                  void *stream);
int icnv_dev_elementwise_f64(const double *X, double *Y, int64_t n, int op, double param, int *err_flag, void *stream);
int icnv_dev_column_stats_f64(const double *X, int64_t G, const int32_t *cells, int64_t n_cells, double *sums, double *sds,
                              void *stream);
int icnv_dev_scale_columns_f64(const double *X, double *Y, int64_t G, int64_t C, const double *sums, double factor,
                               void *stream);
int icnv_dev_clear_noise_f64(const double *X, double *Y, int64_t n, double lo, double hi, double mu, void *stream);
int icnv_dev_proxy_vals_f64(const double *X, double *Y, int64_t n, int m, void *stream);
int icnv_dev_column_minmax_f64(const double *X, int64_t G, int64_t C, double *mins, double *maxs, void *stream);
int icnv_dev_clamp_bounds_f64(const double *X, double *Y, int64_t n, double lower, double upper, void *stream);
int icnv_dev_logistic_adj_f64(const double *X, double *Y, int64_t n, double expr_mean, double delta_midpt, double slope,
                              void *stream);
int icnv_dev_gene_stats_f64(const double *X, int64_t G, int64_t ldx, int64_t C, double *d_sums, int32_t *d_npos, void *stream);
int icnv_dev_gather_rows_f64(const double *X, int64_t ldx, const int32_t *d_keep, int64_t n_keep, double *Y, int64_t C,
                             void *stream);
int icnv_dev_scale_rows_f64(const double *X, double *Y, int64_t G, int64_t C, double *d_sums, double *d_ss, void *stream);
int icnv_dev_csc_gene_stats_f64(const int32_t *d_i, const double *d_x, int64_t nnz, int64_t G, double *d_sums,
                                int32_t *d_npos, void *stream);
int icnv_dev_csc_col_sums_f64(const int32_t *d_p, const int32_t *d_i, const double *d_x, const int32_t *d_keep_map,
                              int64_t C, double *d_cs, void *stream);
int icnv_dev_csc_expand_f64(const int32_t *d_p, const int32_t *d_i, const double *d_x, const int32_t *d_keep_map,
                            int64_t G_out, int64_t C, const double *d_cs, double factor, double *Y, void *stream);
int icnv_dev_scatter_states_per_chr_u8(const uint8_t *gs, int64_t G, int64_t C, const int32_t *d_chr_id, const int32_t *d_grp_of,
                                       uint8_t *out, void *stream);
int icnv_dev_apply_consensus_u8(const uint8_t *S, const uint8_t *cons, int64_t G, int64_t C, const int32_t *d_chr_of,
                                const int32_t *d_grp_of, uint8_t *out, void *stream);
int icnv_dev_state_consensus_u8(const uint8_t *S, int64_t G, int64_t lds, const int32_t *d_cells, const int32_t *h_grp_off,
                                int n_grp, uint8_t *d_cons, int *d_flag, void *stream);
int icnv_dev_cnv_regions_u8(const uint8_t *d_seqs, int64_t G, int64_t lds, int64_t n_seq, const int32_t *d_cols,
                            const int32_t *chr_start, const int32_t *chr_len, int K, const double *gene_start,
                            const double *gene_stop, int64_t *n_regions, void *stream);
int icnv_dev_cnv_regions_records(int64_t *n, const int32_t **seq, const int32_t **chr, const int32_t **first_gene,
                                 const int32_t **last_gene, const int32_t **state, const double **start,
                                 const double **end);
}
extern "C" ICNV_API void icnv_combine_cell_stats(const double *sums, const double *sds, int64_t n, int64_t G, double *mu,
                                                double *sigma);

namespace icnv {

static thread_local char g_err[512] = "";

// One context per device the library was initialised on.  A host thread works on one of them at a time (slot 0 unless
// the multi-device host pipeline put the thread on another one): every helper below - scratch(), pick_stream(), the
// launch counter - goes through ctx().
static Ctx g_ctx[ICNV_MAX_DEVICES];
static thread_local int tl_slot = 0;
static int g_slots = 0;

Ctx &ctx() { return g_ctx[tl_slot]; }
Ctx &ctx_of(int slot) { return g_ctx[slot]; }
int current_slot() { return tl_slot; }
void set_current_slot(int s) { tl_slot = s; }
int device_slots() { return g_slots; }

// ---- host copy pool: pageable caller memory <-> the pinned staging ring --------------------------------------------
// cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread at ~10 GB/s and blocks it, which
// serialises the three-stream slab pipeline.  The library therefore owns pinned slabs and a few copy threads that move
// the caller's slab into / out of them at memory bandwidth, so PCIe sees pinned transfers only.
// One thread's share of a staging copy, with NON-TEMPORAL stores: the destination (a pinned ring slot the DMA engine reads
// next, or the caller's result matrix) is not read again by this thread, so writing around the cache saves the write-allocate
// read of every line - glibc's memcpy only does that above ~3/4 of the last-level cache per call, which is why the
// end-to-end time dropped by 18 % between 2048- and 4096-cell slabs before this (profiles/r02_e2e_probe.txt).
static void stream_copy(char *d, const char *s, size_t n) {
#if defined(__x86_64__) && defined(__SSE2__)
    if (n < (size_t)1 << 16) {
        memcpy(d, s, n);
        return;
    }
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    memcpy(d, s, head);
    d += head;
    s += head;
    n -= head;
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(d), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 48), e);
        s += 64;
        d += 64;
    }
    _mm_sfence();
    memcpy(d, s, n - blocks * 64);
#else
    memcpy(d, s, n);
#endif
}

class CopyPool {
public:
    explicit CopyPool(int n) : n_(n) {
        for (int t = 0; t < n_; ++t) th_.emplace_back([this, t] { loop(t); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return n_; }
    // copies up to two (dst, src, bytes) jobs, cut into equal byte ranges over the threads; returns when both are done
    void copy2(void *d0, const void *s0, size_t n0, void *d1, const void *s1, size_t n1) {
        if (n0 + n1 == 0) return;
        std::unique_lock<std::mutex> lk(mu_);
        job_[0] = {static_cast<char *>(d0), static_cast<const char *>(s0), n0};
        job_[1] = {static_cast<char *>(d1), static_cast<const char *>(s1), n1};
        pending_ = n_;
        ++gen_;
        cv_.notify_all();
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    struct Job {
        char *d;
        const char *s;
        size_t n;
    };
    void loop(int t) {
        unsigned long long seen = 0;
        for (;;) {
            Job j[2];
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                j[0] = job_[0];
                j[1] = job_[1];
            }
            const size_t total = j[0].n + j[1].n;
            const size_t per = ((total + (size_t)n_ - 1) / (size_t)n_ + 4095) & ~(size_t)4095;
            size_t lo = per * (size_t)t, hi = std::min(total, lo + per);
            for (int k = 0; k < 2 && lo < hi; ++k) {   // the byte range [lo, hi) of job 0 followed by job 1
                const size_t base = k == 0 ? 0 : j[0].n;
                const size_t a = std::max(lo, base), b = std::min(hi, base + j[k].n);
                if (a < b) stream_copy(j[k].d + (a - base), j[k].s + (a - base), b - a);
            }
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    Job job_[2] = {};
    int pending_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

static int g_host_threads = 0;   // 0 = default (icnv_set_host_threads)
static int pool_threads_per_device() {
    int n = g_host_threads;
    if (n <= 0) {
        const unsigned hw = std::thread::hardware_concurrency();
        n = (int)std::min<unsigned>(16u, std::max<unsigned>(2u, hw / 2));
    }
    const int slots = std::max(1, g_slots);
    return std::max(2, n / slots);
}

// pinned staging buffer `i` of the current device context (grow-only, freed in icnv_shutdown)
static void *pinned(int i, size_t bytes) {
    Ctx &c = ctx();
    if (c.pin_bytes[i] >= bytes) return c.pin_ptr[i];
    if (c.pin_ptr[i]) {
        cudaFreeHost(c.pin_ptr[i]);
        c.pin_ptr[i] = nullptr;
        c.pin_bytes[i] = 0;
    }
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        set_error(ICNV_E_NOMEM, "cudaHostAlloc(%zu) for the staging ring failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    c.pin_ptr[i] = p;
    c.pin_bytes[i] = bytes;
    return p;
}

// true when the driver can DMA straight from / to p (cudaHostAlloc / cudaHostRegister memory)
static bool is_pinned(const void *p) {
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void *scratch(int slot, size_t bytes) {
    Ctx &c = ctx();
    if (bytes == 0) bytes = 16;
    if (c.slot_bytes[slot] >= bytes) return c.slot_ptr[slot];
    if (slot == SLOT_SEGS) c.up_segs.clear();          // a new allocation does not hold what was uploaded before
    if (slot == SLOT_VIT_ITEMS) c.up_items.clear();
    if (c.slot_ptr[slot]) {
        cudaStreamSynchronize(c.stream);
        cudaFree(c.slot_ptr[slot]);
        c.slot_ptr[slot] = nullptr;
        c.slot_bytes[slot] = 0;
    }
    size_t want = bytes + 256;  // slack so 16-byte over-reads at a column tail stay inside
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        set_error(ICNV_E_NOMEM, "cudaMalloc(%zu) for scratch slot %d failed: %s", want, slot, cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    c.slot_ptr[slot] = p;
    c.slot_bytes[slot] = bytes;
    return p;
}

static int validate_chr(int64_t G, const int32_t *chr_start, const int32_t *chr_len, int K) {
    if (!chr_start || !chr_len || K <= 0) return set_error(ICNV_E_BAD_ARG, "chromosome ranges missing");
    int64_t pos = 0;
    for (int k = 0; k < K; ++k) {
        if (chr_start[k] != pos || chr_len[k] < 0)
            return set_error(ICNV_E_BAD_ARG, "chromosome ranges must tile [0, G) contiguously (chr %d)", k);
        pos += chr_len[k];
    }
    if (pos != G) return set_error(ICNV_E_BAD_ARG, "chromosome ranges cover %lld genes, G = %lld", (long long)pos, (long long)G);
    return ICNV_OK;
}

static int validate_groups(int64_t C, const int32_t *grp_off, const int32_t *grp_idx, int n_grp, bool allow_empty) {
    if (n_grp < 0) return set_error(ICNV_E_BAD_ARG, "n_grp < 0");
    if (n_grp == 0) return allow_empty ? ICNV_OK : set_error(ICNV_E_BAD_ARG, "at least one cell group is required");
    if (!grp_off || !grp_idx) return set_error(ICNV_E_BAD_ARG, "group index lists missing");
    if (grp_off[0] != 0) return set_error(ICNV_E_BAD_ARG, "grp_off[0] must be 0");
    for (int k = 0; k < n_grp; ++k) {
        if (grp_off[k + 1] <= grp_off[k]) return set_error(ICNV_E_BAD_ARG, "group %d is empty", k);
    }
    for (int32_t i = 0; i < grp_off[n_grp]; ++i)
        if (grp_idx[i] < 0 || grp_idx[i] >= C) return set_error(ICNV_E_BAD_ARG, "cell index %d out of range", grp_idx[i]);
    return ICNV_OK;
}

static int check_flag(int *d_flag, cudaStream_t st) {
    int h = 0;
    ICNV_CUDA(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    if (h & 1) return set_error(ICNV_E_NONFINITE, "non-finite value (NA/NaN/Inf) in the expression matrix");
    if (h & 2) return set_error(ICNV_E_UNDERFLOW, "Problems With Underflow");
    return ICNV_OK;
}

// group means of the columns listed in d_cells[off[k]..off[k+1]) -> d_means (G x n_grp)
static int dev_group_means(const double *dX, int64_t G, int64_t ldx, const int32_t *d_cells, const int32_t *h_off,
                           int n_grp, int apply_log, double *d_means, cudaStream_t st) {
    const int chunk = 32;  // fixed: the summation tree must not depend on the device count
    int64_t max_chunks = 0;
    for (int k = 0; k < n_grp; ++k) max_chunks = std::max<int64_t>(max_chunks, (h_off[k + 1] - h_off[k] + chunk - 1) / chunk);
    double *d_part = (double *)scratch(SLOT_PARTIAL, sizeof(double) * (size_t)G * (size_t)max_chunks);
    if (!d_part) return ICNV_E_NOMEM;
    for (int k = 0; k < n_grp; ++k) {
        int64_t n = h_off[k + 1] - h_off[k];
        int64_t n_chunks = (n + chunk - 1) / chunk;
        int rc = icnv_dev_group_partial_sums_f64(dX, G, ldx, d_cells + h_off[k], n, chunk, apply_log, d_part, st);
        if (rc) return rc;
        rc = icnv_dev_combine_partials_f64(d_part, G, n_chunks, n, d_means + G * k, st);
        if (rc) return rc;
    }
    return ICNV_OK;
}

static void destroy_streams(Ctx &c) {
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.s_h2d) cudaStreamDestroy(c.s_h2d);
    if (c.s_d2h) cudaStreamDestroy(c.s_d2h);
    c.stream = c.s_h2d = c.s_d2h = nullptr;
    for (int i = 0; i < 2; ++i) {
        if (c.ev_h2d[i]) cudaEventDestroy(c.ev_h2d[i]);
        if (c.ev_comp[i]) cudaEventDestroy(c.ev_comp[i]);
        if (c.ev_d2h[i]) cudaEventDestroy(c.ev_d2h[i]);
        c.ev_h2d[i] = c.ev_comp[i] = c.ev_d2h[i] = nullptr;
    }
}

}  // namespace icnv

using namespace icnv;

extern "C" {

// ---- lifecycle ------------------------------------------------------------------------------------

int icnv_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(ICNV_E_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return n;
}

static void free_slot(Ctx &c) {
    cudaSetDevice(c.device);
    if (c.stream) cudaStreamSynchronize(c.stream);
    for (int s = 0; s < SLOT_COUNT; ++s) {
        if (c.slot_ptr[s]) cudaFree(c.slot_ptr[s]);
        c.slot_ptr[s] = nullptr;
        c.slot_bytes[s] = 0;
    }
    for (int i = 0; i < 6; ++i) {
        if (c.pin_ptr[i]) cudaFreeHost(c.pin_ptr[i]);
        c.pin_ptr[i] = nullptr;
        c.pin_bytes[i] = 0;
    }
    destroy_streams(c);
    c.ready = false;
}

static int env_int(const char *name, int dflt, bool *set = nullptr) {
    const char *e = getenv(name);
    if (set) *set = e != nullptr;
    return e ? atoi(e) : dflt;
}

// Tuning switches come from the environment ONCE, here - never at launch time - and every one that departs from the
// default is reported, so a stray variable in a user's shell cannot silently change which kernel runs.
static void read_options(Ctx &c, bool verbose) {
    bool set;
    if (const char *e = getenv("ICNV_HMM_MODE")) c.hmm_mode = (e[0] == '0' || e[0] == 'e') ? 0 : (e[0] == '2' ? 2 : 1);
    c.opt_cell_kernel = env_int("ICNV_CELL_KERNEL", 0);
    c.opt_cell_nt = env_int("ICNV_CELL_NT", 0);
    c.opt_cell_variant = env_int("ICNV_CELL_VARIANT", -1);
    c.opt_cell_padq = env_int("ICNV_CELL_PADQ", 1);
    c.opt_cell_lfix = env_int("ICNV_CELL_LFIX", 1);
    c.opt_vfast_warps = env_int("ICNV_VFAST_WARPS", 0);
    c.opt_mf_kernel = env_int("ICNV_MF_KERNEL", -1);
    c.opt_mf_list32 = env_int("ICNV_MF_LIST32", 0);
    c.opt_vit_evict = env_int("ICNV_VIT_EVICT", 2);
    c.opt_slab_cells = env_int("ICNV_SLAB_CELLS", 0, &set);
    if (c.opt_slab_cells < 32 || c.opt_slab_cells > 65536) c.opt_slab_cells = 0;
    if (!verbose) return;
    static const char *names[] = {"ICNV_HMM_MODE", "ICNV_CELL_KERNEL", "ICNV_CELL_NT", "ICNV_CELL_VARIANT", "ICNV_CELL_PADQ",
                                  "ICNV_CELL_LFIX", "ICNV_VFAST_WARPS", "ICNV_MF_KERNEL", "ICNV_MF_LIST32", "ICNV_SLAB_CELLS", "ICNV_VIT_EVICT"};
    for (const char *n : names)
        if (const char *e = getenv(n)) fprintf(stderr, "[infercnv_b200] non-default tuning switch %s=%s (read once at icnv_init)\n", n, e);
}

static int init_slot(int slot, int device) {
    Ctx &c = g_ctx[slot];
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.ready && c.device == device) return ICNV_OK;
    if (c.ready) free_slot(c);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(ICNV_E_NO_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return set_error(ICNV_E_NO_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return set_error(ICNV_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                         prop.major, prop.minor);
    c.sm_count = prop.multiProcessorCount;
    c.smem_optin = (int)prop.sharedMemPerBlockOptin;
    e = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c.s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c.s_d2h, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c.ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c.ev_comp[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c.ev_d2h[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return set_error(ICNV_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    c.device = device;
    c.launches = 0;
    c.table_uploaded = false;
    c.table32_uploaded = false;
    c.math_tables_uploaded = false;
    c.up_segs.clear();
    c.up_items.clear();
    c.hmm_list_count = nullptr;
    c.hmm_seq_counts = nullptr;
    c.hmm_seq_k = 0;
    c.rg_n = 0;
    read_options(c, slot == 0);
    c.ready = true;
    return ICNV_OK;
}

static std::mutex g_init_mu;

int icnv_init(int device) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    Ctx &c0 = g_ctx[0];
    if (device < 0) {
        if (c0.ready) return ICNV_OK;   // keep whatever icnv_init / icnv_init_devices set up
        device = 0;
    }
    if (c0.ready && c0.device == device && g_slots == 1) return ICNV_OK;
    int n = icnv_device_count();
    if (n <= 0) return set_error(ICNV_E_NO_DEVICE, "no CUDA device available (the library has no CPU fallback)");
    if (device >= n) return set_error(ICNV_E_BAD_ARG, "device %d out of range (%d devices)", device, n);
    for (int s = 1; s < g_slots; ++s)
        if (g_ctx[s].ready) free_slot(g_ctx[s]);
    int rc = init_slot(0, device);
    if (rc) return rc;
    g_slots = 1;
    cudaSetDevice(device);
    return ICNV_OK;
}

int icnv_init_devices(int n_devices, const int *device_ids) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    int n = icnv_device_count();
    if (n <= 0) return set_error(ICNV_E_NO_DEVICE, "no CUDA device available (the library has no CPU fallback)");
    if (n_devices <= 0) n_devices = n;   // "all of them"
    if (n_devices > ICNV_MAX_DEVICES || n_devices > n)
        return set_error(ICNV_E_BAD_ARG, "icnv_init_devices: %d devices requested, %d present (at most %d)", n_devices, n,
                         ICNV_MAX_DEVICES);
    for (int i = 0; i < n_devices; ++i) {
        const int d = device_ids ? device_ids[i] : i;
        if (d < 0 || d >= n) return set_error(ICNV_E_BAD_ARG, "icnv_init_devices: device %d out of range", d);
        for (int j = 0; j < i; ++j)
            if ((device_ids ? device_ids[j] : j) == d) return set_error(ICNV_E_BAD_ARG, "icnv_init_devices: device %d listed twice", d);
    }
    for (int s = n_devices; s < g_slots; ++s)
        if (g_ctx[s].ready) free_slot(g_ctx[s]);
    for (int i = 0; i < n_devices; ++i) {
        int rc = init_slot(i, device_ids ? device_ids[i] : i);
        if (rc) return rc;
    }
    g_slots = n_devices;
    cudaSetDevice(g_ctx[0].device);
    return ICNV_OK;
}

int icnv_devices_in_use(void) { return g_slots; }

int icnv_set_host_threads(int n) {
    if (n < 0 || n > 256) return set_error(ICNV_E_BAD_ARG, "icnv_set_host_threads: 0 (default) .. 256");
    g_host_threads = n;
    return ICNV_OK;
}

void icnv_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    for (int s = 0; s < ICNV_MAX_DEVICES; ++s) {
        Ctx &c = g_ctx[s];
        std::lock_guard<std::mutex> lk2(c.mu);
        if (c.ready) free_slot(c);
    }
    g_slots = 0;
}

const char *icnv_last_error(void) { return g_err; }
const char *icnv_version(void) { return "infercnv_b200 0.1.0 (sm_100a)"; }
int64_t icnv_launch_count(void) {
    int64_t n = 0;
    for (int s = 0; s < std::max(1, g_slots); ++s) n += g_ctx[s].launches.load();
    return n;
}

// ---- device-side composition of the smooth block ---------------------------------------------------------

int icnv_dev_smooth_block_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                              const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx, int n_grp,
                              int apply_log, double threshold, int window, int use_bounds, void *stream) {
    ICNV_REQUIRE_READY();
    if (!X || !Y || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_dev_smooth_block_f64: bad argument");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    rc = validate_groups(C, grp_off, grp_idx, n_grp, false);
    if (rc) return rc;
    cudaStream_t st = pick_stream(stream);
    const int64_t n_ref = grp_off[n_grp];

    // index lists: [ref cells as given | 0..n_ref-1 (columns of the compact pass-1 buffer)]
    int32_t *d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * 2 * (size_t)n_ref);
    if (!d_idx) return ICNV_E_NOMEM;
    std::vector<int32_t> iota((size_t)n_ref);
    std::iota(iota.begin(), iota.end(), 0);
    ICNV_CUDA(cudaMemcpyAsync(d_idx, grp_idx, sizeof(int32_t) * (size_t)n_ref, cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaMemcpyAsync(d_idx + n_ref, iota.data(), sizeof(int32_t) * (size_t)n_ref, cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaStreamSynchronize(st));  // iota is a stack-lifetime host buffer

    double *d_means = (double *)scratch(SLOT_MEANS, sizeof(double) * (size_t)G * (size_t)n_grp);
    double *d_b = (double *)scratch(SLOT_BOUNDS, sizeof(double) * (size_t)G * 6);
    double *d_T = (double *)scratch(SLOT_TMP, sizeof(double) * (size_t)G * (size_t)n_ref);
    int *d_flag = (int *)scratch(SLOT_MISC, 64);
    if (!d_means || !d_b || !d_T || !d_flag) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    double *lo1 = d_b, *hi1 = d_b + G, *mid1 = d_b + 2 * G, *lo2 = d_b + 3 * G, *hi2 = d_b + 4 * G, *mid2 = d_b + 5 * G;

    // pass 0: step-8 reference means on (log-transformed) input   (ops.R:771 -> :1708)
    rc = dev_group_means(X, G, G, d_idx, grp_off, n_grp, apply_log ? 1 : 0, d_means, st);
    if (rc) return rc;
    rc = icnv_dev_bounds_from_means_f64(d_means, G, n_grp, lo1, hi1, mid1, st);
    if (rc) return rc;
    // pass 1: steps 4..11 on the reference cells only -> compact buffer T (one column per list entry)
    rc = icnv_dev_cell_pipeline_f64(X, G, G, d_idx, n_ref, d_T, G, chr_start, chr_len, K, apply_log,
                                    use_bounds ? lo1 : nullptr, use_bounds ? hi1 : nullptr, use_bounds ? nullptr : mid1,
                                    threshold, window, 1, nullptr, nullptr, nullptr, 0, d_flag, st);
    if (rc) return rc;
    // step-12 reference means on the centred, smoothed reference cells   (ops.R:952)
    rc = dev_group_means(d_T, G, G, d_idx + n_ref, grp_off, n_grp, 0, d_means, st);
    if (rc) return rc;
    rc = icnv_dev_bounds_from_means_f64(d_means, G, n_grp, lo2, hi2, mid2, st);
    if (rc) return rc;
    // pass 2: every cell, read once, written once
    rc = icnv_dev_cell_pipeline_f64(X, G, G, nullptr, C, Y, G, chr_start, chr_len, K, apply_log,
                                    use_bounds ? lo1 : nullptr, use_bounds ? hi1 : nullptr, use_bounds ? nullptr : mid1,
                                    threshold, window, 1, use_bounds ? lo2 : nullptr, use_bounds ? hi2 : nullptr,
                                    use_bounds ? nullptr : mid2, 1, d_flag, st);
    if (rc) return rc;
    return ICNV_OK;
}

// ---- host-pointer entry points --------------------------------------------------------------------------------

#define ICNV_HOST_PROLOGUE()                       \
    ICNV_REQUIRE_READY();                          \
    Ctx &c = ctx();                                \
    std::lock_guard<std::mutex> lk(c.mu);          \
    ICNV_CUDA(cudaSetDevice(c.device));            \
    cudaStream_t st = c.stream;                    \
    (void)st

static int upload_matrix(const double *X, int64_t n, double **dX, int slot, cudaStream_t st) {
    *dX = (double *)scratch(slot, sizeof(double) * (size_t)n);
    if (!*dX) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemcpyAsync(*dX, X, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    return ICNV_OK;
}

int icnv_ref_means_f64(const double *X, int64_t G, int64_t C, const int32_t *grp_off, const int32_t *grp_idx, int n_grp,
                       int inv_log, double *means) {
    ICNV_HOST_PROLOGUE();
    if (!X || !means || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_ref_means_f64: bad argument");
    int rc = validate_groups(C, grp_off, grp_idx, n_grp, false);
    if (rc) return rc;
    double *dX;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    int64_t n_ref = grp_off[n_grp];
    int32_t *d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)n_ref);
    double *d_means = (double *)scratch(SLOT_MEANS, sizeof(double) * (size_t)G * (size_t)n_grp);
    if (!d_idx || !d_means) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemcpyAsync(d_idx, grp_idx, sizeof(int32_t) * (size_t)n_ref, cudaMemcpyHostToDevice, st));
    rc = dev_group_means(dX, G, G, d_idx, grp_off, n_grp, inv_log ? 2 : 0, d_means, st);
    if (rc) return rc;
    if (inv_log && (rc = icnv_dev_invlog_finish_f64(d_means, G * n_grp, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(means, d_means, sizeof(double) * (size_t)G * (size_t)n_grp, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

// one pass of the per-cell kernel over a host matrix: upload, run, download
static int host_cell_pipeline(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                              const int32_t *chr_len, int K, const double *h_means, int n_grp, int use_bounds,
                              int window, int center, cudaStream_t st) {
    double *dX, *dY;
    int rc;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(G * C));
    int *d_flag = (int *)scratch(SLOT_MISC, 64);
    if (!dY || !d_flag) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    double *lo = nullptr, *hi = nullptr, *mid = nullptr;
    if (h_means) {
        double *d_means = (double *)scratch(SLOT_MEANS, sizeof(double) * (size_t)G * (size_t)n_grp);
        double *d_b = (double *)scratch(SLOT_BOUNDS, sizeof(double) * (size_t)G * 6);
        if (!d_means || !d_b) return ICNV_E_NOMEM;
        ICNV_CUDA(cudaMemcpyAsync(d_means, h_means, sizeof(double) * (size_t)G * (size_t)n_grp, cudaMemcpyHostToDevice, st));
        if ((rc = icnv_dev_bounds_from_means_f64(d_means, G, n_grp, d_b, d_b + G, d_b + 2 * G, st))) return rc;
        if (use_bounds) {
            lo = d_b;
            hi = d_b + G;
        } else {
            mid = d_b + 2 * G;
        }
    }
    std::vector<int32_t> one_start{0}, one_len{(int32_t)G};
    const int32_t *cs = chr_start ? chr_start : one_start.data();
    const int32_t *cl = chr_len ? chr_len : one_len.data();
    int kk = chr_start ? K : 1;
    rc = icnv_dev_cell_pipeline_f64(dX, G, G, nullptr, C, dY, G, cs, cl, kk, 0, lo, hi, mid, 0.0, window, center, nullptr,
                                    nullptr, nullptr, 0, d_flag, st);
    if (rc) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    return check_flag(d_flag, st);
}

int icnv_subtract_ref_f64(const double *X, double *Y, int64_t G, int64_t C, const double *means, int n_grp,
                          int use_bounds) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || !means || G <= 0 || C <= 0 || n_grp <= 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_subtract_ref_f64: bad argument");
    return host_cell_pipeline(X, Y, G, C, nullptr, nullptr, 0, means, n_grp, use_bounds, 0, 0, st);
}

int icnv_smooth_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start, const int32_t *chr_len,
                    int K, int window) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_smooth_f64: bad argument");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    if (window >= 2 && (window & 1) == 0)
        return set_error(ICNV_E_BAD_ARG, "window_length %d is even: refusing (the reference's result is accidental)", window);
    return host_cell_pipeline(X, Y, G, C, chr_start, chr_len, K, nullptr, 0, 0, window, 0, st);
}

int icnv_center_f64(const double *X, double *Y, int64_t G, int64_t C, int use_median) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_center_f64: bad argument");
    return host_cell_pipeline(X, Y, G, C, nullptr, nullptr, 0, nullptr, 0, 0, 0, use_median ? 1 : 2, st);
}

// ---- slab pipeline: cells are independent once the reference means are known, so the matrix moves
// through the GPU in slabs of SLAB_CELLS columns - H2D of slab i+1, kernels on slab i and D2H of slab
// i-1 run concurrently on three streams (double-buffered device slabs).  PCIe, not the kernels, bounds
// the host-pointer entry points; this hides everything but the slower copy direction. ---------------
static const int64_t SLAB_CELLS = 1024;
static const int64_t SLAB_CELLS_SMOOTH = 256;

struct HmmModel {
    int m;
    const double *Pi, *delta, *mean, *sd;
};

// Y (optional) = smooth block of X; states (optional) = per-cell Viterbi of the block's output (or of X
// itself when do_smooth == 0).
// states (int32, -1 = unassigned) or states8 (uint8, 255 = unassigned): at most one of them is non-NULL.
// Works on the cells [c_lo, c_hi) of the G x C host matrices on the calling thread's device context; the reference
// groups index the whole matrix (every device reduces all reference cells itself, in list order).
static int host_pipeline_range(Ctx &c, const double *X, double *Y, int32_t *states, uint8_t *states8, int64_t G, int64_t c_lo,
                               int64_t c_hi, const int32_t *chr_start, const int32_t *chr_len, int K, int do_smooth,
                               const int32_t *grp_off, const int32_t *grp_idx, int n_grp, int apply_log, double threshold,
                               int window, int use_bounds, const HmmModel *hmm, CopyPool *pool) {
    cudaStream_t sc = c.stream, sh = c.s_h2d, sd = c.s_d2h;
    int rc;
    const int64_t C = c_hi - c_lo;
    int *d_flag = (int *)scratch(SLOT_MISC, 64);
    if (!d_flag) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), sc));
    double *lo1 = nullptr, *hi1 = nullptr, *mid1 = nullptr, *lo2 = nullptr, *hi2 = nullptr, *mid2 = nullptr;
    // page-locked caller memory goes over PCIe directly; anything else through the pinned ring (pool != NULL)
    const bool stage_in = pool && !is_pinned(X);
    const bool stage_out = pool && ((Y && !is_pinned(Y)) || (states && !is_pinned(states)) || (states8 && !is_pinned(states8)));

    // ---- slabs ------------------------------------------------------------------------------------------
    // Fill and drain of the copy pipeline cost one slab each way.  Without the HMM a slab's kernel time (~0.05 ms per 256
    // cells) is far below its PCIe time (0.37 ms), so small slabs only shorten fill / drain; the per-cell Viterbi is bounded
    // below by its longest chromosome's serial recursion whatever the slab size, so HMM calls keep the measured 1024.
    int64_t slab_cells = hmm ? SLAB_CELLS : SLAB_CELLS_SMOOTH;
    if (stage_in || stage_out) {
        // pageable caller memory: fewer, larger hand-overs to the copy threads - about 320 MB of input per slab (4000 cells at
        // 10 000 genes: 362 ms against 386 ms with 1024-cell slabs for c3, profiles/r02_e2e_probe.txt), but at least ten slabs
        // so that filling and draining the pipeline stays a small share
        int64_t by_bytes = ((int64_t)(320e6 / (8.0 * (double)G)) + 31) / 32 * 32;
        by_bytes = std::max<int64_t>(256, std::min<int64_t>(8192, by_bytes));
        const int64_t by_count = std::max<int64_t>(SLAB_CELLS, ((C + 9) / 10 + 31) / 32 * 32);
        slab_cells = std::min(by_bytes, by_count);
    }
    if (c.opt_slab_cells) slab_cells = c.opt_slab_cells;
    const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(slab_cells, C));
    const size_t slab_elems = (size_t)G * (size_t)slab;
    double *dIn[2], *dOut[2] = {nullptr, nullptr};
    uint8_t *dSt[2] = {nullptr, nullptr};
    int32_t *dW[2] = {nullptr, nullptr};
    double *pIn[2] = {nullptr, nullptr}, *pOut[2] = {nullptr, nullptr};
    void *pSt[2] = {nullptr, nullptr};
    const size_t st_elem = states ? sizeof(int32_t) : 1;
    for (int b = 0; b < 2; ++b) {
        dIn[b] = (double *)scratch(SLOT_SLAB_IN0 + b, sizeof(double) * slab_elems);
        if (!dIn[b]) return ICNV_E_NOMEM;
        if (do_smooth) {
            dOut[b] = (double *)scratch(SLOT_SLAB_OUT0 + b, sizeof(double) * slab_elems);
            if (!dOut[b]) return ICNV_E_NOMEM;
        }
        if (hmm) {
            dSt[b] = (uint8_t *)scratch(SLOT_SLAB_ST0 + b, slab_elems);
            if (!dSt[b]) return ICNV_E_NOMEM;
            if (states) {
                dW[b] = (int32_t *)scratch(SLOT_SLAB_W0 + b, sizeof(int32_t) * slab_elems);
                if (!dW[b]) return ICNV_E_NOMEM;
            }
        }
        if (stage_in && !(pIn[b] = (double *)pinned(b, sizeof(double) * slab_elems))) return ICNV_E_NOMEM;
        if (stage_out && do_smooth && Y && !(pOut[b] = (double *)pinned(2 + b, sizeof(double) * slab_elems))) return ICNV_E_NOMEM;
        if (stage_out && hmm && !(pSt[b] = pinned(4 + b, st_elem * slab_elems))) return ICNV_E_NOMEM;
    }
    const int64_t n_slabs = (C + slab - 1) / slab;
    auto slab_cells_of = [&](int64_t i) { return std::min<int64_t>(slab, C - i * slab); };
    // results of slab j leave the pinned ring for the caller's buffers (after its D2H has completed); together with the
    // input of slab `in` entering the ring when in >= 0 - one hand-over to the copy threads for both directions
    auto host_copies = [&](int64_t in, int64_t out) -> int {
        void *d0 = nullptr, *d1 = nullptr;
        const void *s0 = nullptr, *s1 = nullptr;
        size_t n0 = 0, n1 = 0;
        if (in >= 0 && in < n_slabs && stage_in) {
            const int b = (int)(in & 1);
            if (in >= 2) ICNV_CUDA(cudaEventSynchronize(c.ev_h2d[b]));   // the ring slot has been read by slab in-2's H2D
            d0 = pIn[b];
            s0 = X + G * (c_lo + in * slab);
            n0 = sizeof(double) * (size_t)G * (size_t)slab_cells_of(in);
        }
        if (out >= 0 && out < n_slabs && stage_out) {
            const int b = (int)(out & 1);
            ICNV_CUDA(cudaEventSynchronize(c.ev_d2h[b]));
            const int64_t c0 = c_lo + out * slab, nc = slab_cells_of(out);
            if (do_smooth && Y) {
                d1 = Y + G * c0;
                s1 = pOut[b];
                n1 = sizeof(double) * (size_t)G * (size_t)nc;
            }
            if (hmm) {   // the states ride along: a third job is not worth a second hand-over
                char *dst = states ? (char *)(states + G * c0) : (char *)(states8 + G * c0);
                if (n1 == 0) {
                    d1 = dst;
                    s1 = pSt[b];
                    n1 = st_elem * (size_t)G * (size_t)nc;
                } else {
                    pool->copy2(d0, s0, n0, d1, s1, n1);
                    d0 = dst;
                    s0 = pSt[b];
                    n0 = st_elem * (size_t)G * (size_t)nc;
                    d1 = nullptr;
                    n1 = 0;
                }
            }
        }
        pool->copy2(d0, s0, n0, d1, s1, n1);
        return ICNV_OK;
    };
    // H2D of slab i (buffer i & 1) may start once the kernels of slab i-2 have consumed the buffer
    auto issue_h2d = [&](int64_t i) -> int {
        const int b = (int)(i & 1);
        const int64_t c0 = c_lo + i * slab, nc = slab_cells_of(i);
        if (i >= 2) ICNV_CUDA(cudaStreamWaitEvent(sh, c.ev_comp[b], 0));
        ICNV_CUDA(cudaMemcpyAsync(dIn[b], stage_in ? pIn[b] : X + G * c0, sizeof(double) * (size_t)G * (size_t)nc,
                                  cudaMemcpyHostToDevice, sh));
        ICNV_CUDA(cudaEventRecord(c.ev_h2d[b], sh));
        return ICNV_OK;
    };
    int64_t h2d_issued = 0;

    if (do_smooth) {
        // ---- reference pre-passes on a compact copy of the reference columns (<= ~10 % of the cells) ----
        const int64_t n_ref = grp_off[n_grp];
        double *d_ref = (double *)scratch(SLOT_REFX, sizeof(double) * (size_t)G * (size_t)n_ref);
        double *d_T = (double *)scratch(SLOT_TMP, sizeof(double) * (size_t)G * (size_t)n_ref);
        double *d_means = (double *)scratch(SLOT_MEANS, sizeof(double) * (size_t)G * (size_t)n_grp);
        double *d_b = (double *)scratch(SLOT_BOUNDS, sizeof(double) * (size_t)G * 6);
        int32_t *d_iota = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)n_ref);
        if (!d_ref || !d_T || !d_means || !d_b || !d_iota) return ICNV_E_NOMEM;
        for (int64_t i = 0; i < n_ref;) {  // runs of consecutive cells go up in one copy (staged: a ring slot at a time)
            int64_t j = i + 1;
            while (j < n_ref && grp_idx[j] == grp_idx[j - 1] + 1) ++j;
            if (!stage_in) {
                ICNV_CUDA(cudaMemcpyAsync(d_ref + G * i, X + G * (int64_t)grp_idx[i], sizeof(double) * (size_t)(G * (j - i)),
                                          cudaMemcpyHostToDevice, sc));
            } else {
                for (int64_t q = i; q < j; q += slab) {
                    const int64_t nq = std::min<int64_t>(slab, j - q);
                    const int b = (int)((q / slab) & 1);
                    ICNV_CUDA(cudaEventSynchronize(c.ev_h2d[b]));   // (a recorded-never event is complete)
                    pool->copy2(pIn[b], X + G * ((int64_t)grp_idx[i] + (q - i)), sizeof(double) * (size_t)(G * nq), nullptr, nullptr, 0);
                    ICNV_CUDA(cudaMemcpyAsync(d_ref + G * q, pIn[b], sizeof(double) * (size_t)(G * nq), cudaMemcpyHostToDevice, sc));
                    ICNV_CUDA(cudaEventRecord(c.ev_h2d[b], sc));
                }
            }
            i = j;
        }
        std::vector<int32_t> iota((size_t)n_ref);
        std::iota(iota.begin(), iota.end(), 0);
        ICNV_CUDA(cudaMemcpyAsync(d_iota, iota.data(), sizeof(int32_t) * (size_t)n_ref, cudaMemcpyHostToDevice, sc));
        if (!stage_in) {
            // the first two slabs queue up behind the reference columns and cross PCIe while the pre-passes run
            for (; h2d_issued < std::min<int64_t>(2, n_slabs); ++h2d_issued)
                if ((rc = issue_h2d(h2d_issued))) return rc;
        }
        ICNV_CUDA(cudaStreamSynchronize(sc));
        lo1 = d_b; hi1 = d_b + G; mid1 = d_b + 2 * G; lo2 = d_b + 3 * G; hi2 = d_b + 4 * G; mid2 = d_b + 5 * G;
        if ((rc = dev_group_means(d_ref, G, G, d_iota, grp_off, n_grp, apply_log ? 1 : 0, d_means, sc))) return rc;
        if ((rc = icnv_dev_bounds_from_means_f64(d_means, G, n_grp, lo1, hi1, mid1, sc))) return rc;
        rc = icnv_dev_cell_pipeline_f64(d_ref, G, G, nullptr, n_ref, d_T, G, chr_start, chr_len, K, apply_log,
                                        use_bounds ? lo1 : nullptr, use_bounds ? hi1 : nullptr, use_bounds ? nullptr : mid1,
                                        threshold, window, 1, nullptr, nullptr, nullptr, 0, d_flag, sc);
        if (rc) return rc;
        if ((rc = dev_group_means(d_T, G, G, d_iota, grp_off, n_grp, 0, d_means, sc))) return rc;
        if ((rc = icnv_dev_bounds_from_means_f64(d_means, G, n_grp, lo2, hi2, mid2, sc))) return rc;
        if (!use_bounds) lo1 = hi1 = lo2 = hi2 = nullptr;
        else mid1 = mid2 = nullptr;
    }

    if (stage_in) {   // the first slab enters the ring while the reference pre-passes run on the device
        if ((rc = host_copies(0, -1))) return rc;
    }
    for (int64_t i = 0; i < n_slabs; ++i) {
        const int b = (int)(i & 1);
        const int64_t c0 = c_lo + i * slab, nc = slab_cells_of(i);
        const size_t bytes = sizeof(double) * (size_t)G * (size_t)nc;
        if (stage_in) {
            if (h2d_issued <= i && (rc = issue_h2d(h2d_issued++))) return rc;
        } else {
            for (; h2d_issued <= std::min<int64_t>(i + 1, n_slabs - 1); ++h2d_issued)
                if ((rc = issue_h2d(h2d_issued))) return rc;
        }
        // kernels: need the slab on the device and the output buffers of slab i-2 drained
        ICNV_CUDA(cudaStreamWaitEvent(sc, c.ev_h2d[b], 0));
        if (i >= 2) ICNV_CUDA(cudaStreamWaitEvent(sc, c.ev_d2h[b], 0));
        const double *hmm_in = dIn[b];
        if (do_smooth) {
            rc = icnv_dev_cell_pipeline_f64(dIn[b], G, G, nullptr, nc, dOut[b], G, chr_start, chr_len, K, apply_log, lo1, hi1,
                                            mid1, threshold, window, 1, lo2, hi2, mid2, 1, d_flag, sc);
            if (rc) return rc;
            hmm_in = dOut[b];
        }
        if (hmm) {
            rc = icnv_dev_viterbi_f64(hmm_in, G, nc, chr_start, chr_len, K, hmm->m, hmm->Pi, hmm->delta, hmm->mean, hmm->sd,
                                      0, dSt[b], nullptr, d_flag, sc);
            if (rc) return rc;
            if (states && (rc = icnv_dev_widen_states(dSt[b], dW[b], (int64_t)G * nc, sc))) return rc;
        }
        ICNV_CUDA(cudaEventRecord(c.ev_comp[b], sc));
        // D2H (staged: the ring slot of slab i-2 was emptied by host_copies(.., i-2) below, two iterations ago)
        ICNV_CUDA(cudaStreamWaitEvent(sd, c.ev_comp[b], 0));
        if (do_smooth && Y) ICNV_CUDA(cudaMemcpyAsync(stage_out ? pOut[b] : Y + G * c0, dOut[b], bytes, cudaMemcpyDeviceToHost, sd));
        if (hmm && states)
            ICNV_CUDA(cudaMemcpyAsync(stage_out ? (int32_t *)pSt[b] : states + G * c0, dW[b],
                                      sizeof(int32_t) * (size_t)G * (size_t)nc, cudaMemcpyDeviceToHost, sd));
        if (hmm && states8)
            ICNV_CUDA(cudaMemcpyAsync(stage_out ? (uint8_t *)pSt[b] : states8 + G * c0, dSt[b], (size_t)G * (size_t)nc,
                                      cudaMemcpyDeviceToHost, sd));
        ICNV_CUDA(cudaEventRecord(c.ev_d2h[b], sd));
        // host side of the ring while the device works on slab i: slab i+1 in, slab i-1 out
        if (stage_in || stage_out)
            if ((rc = host_copies(stage_in ? i + 1 : -1, stage_out ? i - 1 : -1))) return rc;
    }
    if (stage_out && (rc = host_copies(-1, n_slabs - 1))) return rc;
    ICNV_CUDA(cudaStreamSynchronize(sh));
    ICNV_CUDA(cudaStreamSynchronize(sd));
    return check_flag(d_flag, sc);
}

// The cells are cut into one contiguous range per device context; each range is driven by its own host thread (a CUDA
// context per thread, its own streams, scratch and ring).  One context: the calling thread does the work itself.
static int host_pipeline(Ctx &c0, const double *X, double *Y, int32_t *states, uint8_t *states8, int64_t G, int64_t C,
                         const int32_t *chr_start, const int32_t *chr_len, int K, int do_smooth, const int32_t *grp_off,
                         const int32_t *grp_idx, int n_grp, int apply_log, double threshold, int window, int use_bounds,
                         const HmmModel *hmm) {
    const int n_dev = (int)std::min<int64_t>(std::max(1, device_slots()), std::max<int64_t>(1, C / 64));
    const bool pageable = !is_pinned(X) || (Y && !is_pinned(Y)) || (states && !is_pinned(states)) || (states8 && !is_pinned(states8));
    const int nt = pool_threads_per_device();
    if (n_dev <= 1) {
        std::unique_ptr<CopyPool> pool(pageable ? new CopyPool(nt) : nullptr);
        return host_pipeline_range(c0, X, Y, states, states8, G, 0, C, chr_start, chr_len, K, do_smooth, grp_off, grp_idx, n_grp,
                                   apply_log, threshold, window, use_bounds, hmm, pool.get());
    }
    std::vector<int> rcs((size_t)n_dev, ICNV_OK);
    std::vector<std::string> msgs((size_t)n_dev);
    std::vector<std::thread> th;
    const int caller_slot = current_slot();
    for (int d = 0; d < n_dev; ++d) {
        const int64_t lo = C * d / n_dev, hi = C * (d + 1) / n_dev;
        th.emplace_back([&, d, lo, hi] {
            set_current_slot(d);
            Ctx &c = ctx();
            std::unique_lock<std::mutex> lk(c.mu, std::defer_lock);
            if (d != caller_slot) lk.lock();   // the caller already holds its own context's lock
            int rc = ICNV_OK;
            if (cudaSetDevice(c.device) != cudaSuccess) rc = set_error(ICNV_E_CUDA, "cudaSetDevice(%d) failed", c.device);
            if (!rc) {
                std::unique_ptr<CopyPool> pool(pageable ? new CopyPool(nt) : nullptr);
                rc = host_pipeline_range(c, X, Y, states, states8, G, lo, hi, chr_start, chr_len, K, do_smooth, grp_off, grp_idx,
                                         n_grp, apply_log, threshold, window, use_bounds, hmm, pool.get());
            }
            rcs[(size_t)d] = rc;
            if (rc) msgs[(size_t)d] = icnv_last_error();   // the message is per thread: carry it to the caller
        });
    }
    for (auto &t : th) t.join();
    cudaSetDevice(c0.device);
    for (int d = 0; d < n_dev; ++d)
        if (rcs[(size_t)d]) return set_error(rcs[(size_t)d], "device %d: %s", ctx_of(d).device, msgs[(size_t)d].c_str());
    return ICNV_OK;
}

int icnv_smooth_block_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                          const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx, int n_grp,
                          int apply_log, double threshold, int window, int use_bounds) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_smooth_block_f64: bad argument");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    if ((rc = validate_groups(C, grp_off, grp_idx, n_grp, false))) return rc;
    return host_pipeline(c, X, Y, nullptr, nullptr, G, C, chr_start, chr_len, K, 1, grp_off, grp_idx, n_grp, apply_log,
                         threshold, window, use_bounds, nullptr);
}

/* Fused smooth block + per-cell HMM in one pass over the matrix: run() steps 4..14 and step 17 for
 * analysis_mode = "cells" with prune_outliers = FALSE (the defaults between them do not touch expr.data).
 * Saves the second upload of the matrix that two separate calls need. */
static int smooth_hmm_impl(const double *X, double *Y, int32_t *states, uint8_t *states8, int64_t G, int64_t C,
                           const int32_t *chr_start, const int32_t *chr_len, int K, const int32_t *grp_off,
                           const int32_t *grp_idx, int n_grp, int apply_log, double threshold, int window, int use_bounds,
                           int m, const double *Pi, const double *delta, const double *mean, const double *sd) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || (!states && !states8) || G <= 0 || C <= 0 || !Pi || !delta || !mean || !sd)
        return set_error(ICNV_E_BAD_ARG, "icnv_smooth_hmm_f64: bad argument");
    if (m != 6 && m != 3) return set_error(ICNV_E_BAD_ARG, "m must be 6 or 3");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    if ((rc = validate_groups(C, grp_off, grp_idx, n_grp, false))) return rc;
    HmmModel hm{m, Pi, delta, mean, sd};
    return host_pipeline(c, X, Y, states, states8, G, C, chr_start, chr_len, K, 1, grp_off, grp_idx, n_grp, apply_log,
                         threshold, window, use_bounds, &hm);
}

int icnv_smooth_hmm_f64(const double *X, double *Y, int32_t *states, int64_t G, int64_t C, const int32_t *chr_start,
                        const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx, int n_grp,
                        int apply_log, double threshold, int window, int use_bounds, int m, const double *Pi,
                        const double *delta, const double *mean, const double *sd) {
    return smooth_hmm_impl(X, Y, states, nullptr, G, C, chr_start, chr_len, K, grp_off, grp_idx, n_grp, apply_log, threshold,
                           window, use_bounds, m, Pi, delta, mean, sd);
}

int icnv_smooth_hmm_u8_f64(const double *X, double *Y, uint8_t *states, int64_t G, int64_t C, const int32_t *chr_start,
                           const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx, int n_grp,
                           int apply_log, double threshold, int window, int use_bounds, int m, const double *Pi,
                           const double *delta, const double *mean, const double *sd) {
    return smooth_hmm_impl(X, Y, nullptr, states, G, C, chr_start, chr_len, K, grp_off, grp_idx, n_grp, apply_log, threshold,
                           window, use_bounds, m, Pi, delta, mean, sd);
}

static int viterbi_impl(const double *X, int64_t G, int64_t C, const int32_t *chr_start, const int32_t *chr_len, int K,
                        const int32_t *grp_off, const int32_t *grp_idx, int n_grp, int m, const double *Pi,
                        const double *delta, const double *mean, const double *sd, int32_t *states, uint8_t *states8,
                        double *margins) {
    ICNV_HOST_PROLOGUE();
    if (!X || (!states && !states8) || G <= 0 || C <= 0 || !Pi || !delta || !mean || !sd)
        return set_error(ICNV_E_BAD_ARG, "icnv_viterbi_f64: bad argument");
    if (m != 6 && m != 3) return set_error(ICNV_E_BAD_ARG, "m must be 6 or 3");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    rc = validate_groups(C, grp_off, grp_idx, n_grp, true);
    if (rc) return rc;
    if (n_grp == 0 && !margins) {  // per-cell mode: slab pipeline (copies overlap the kernels)
        HmmModel hm{m, Pi, delta, mean, sd};
        return host_pipeline(c, X, nullptr, states, states8, G, C, chr_start, chr_len, K, 0, nullptr, nullptr, 0, 0, 0.0, 0, 0,
                             &hm);
    }
    double *dX;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    int *d_flag = (int *)scratch(SLOT_MISC, 64);
    if (!d_flag) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    int32_t *d_out = (int32_t *)scratch(SLOT_OUT, sizeof(int32_t) * (size_t)(G * C));
    if (!d_out) return ICNV_E_NOMEM;

    if (n_grp == 0) {
        uint8_t *d_st = (uint8_t *)scratch(SLOT_STATES, (size_t)(G * C));
        double *d_mg = margins ? (double *)scratch(SLOT_TMP, sizeof(double) * (size_t)K * (size_t)C) : nullptr;
        if (!d_st || (margins && !d_mg)) return ICNV_E_NOMEM;
        rc = icnv_dev_viterbi_f64(dX, G, C, chr_start, chr_len, K, m, Pi, delta, mean, sd, 0, d_st, d_mg, d_flag, st);
        if (rc) return rc;
        if ((rc = icnv_dev_widen_states(d_st, d_out, G * C, st))) return rc;
        if (margins)
            ICNV_CUDA(cudaMemcpyAsync(margins, d_mg, sizeof(double) * (size_t)K * (size_t)C, cudaMemcpyDeviceToHost, st));
    } else {
        // group modes: x = rowMeans(X[chr, group]) (HMM.R:383), per-group median sd, trace broadcast
        const int64_t n_idx = grp_off[n_grp];
        int32_t *d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)(n_idx + C));
        double *d_xm = (double *)scratch(SLOT_MEANS, sizeof(double) * (size_t)G * (size_t)n_grp);
        double *d_sd = (double *)scratch(SLOT_BOUNDS, sizeof(double) * (size_t)std::max<int64_t>(n_grp, 6 * G));
        uint8_t *d_st = (uint8_t *)scratch(SLOT_STATES, (size_t)G * (size_t)n_grp);
        double *d_mg = margins ? (double *)scratch(SLOT_TMP, sizeof(double) * (size_t)K * (size_t)n_grp) : nullptr;
        if (!d_idx || !d_xm || !d_sd || !d_st || (margins && !d_mg)) return ICNV_E_NOMEM;
        std::vector<int32_t> grp_of((size_t)C, -1);
        std::vector<double> sdm((size_t)n_grp);
        for (int b = 0; b < n_grp; ++b) {
            for (int32_t i = grp_off[b]; i < grp_off[b + 1]; ++i) grp_of[grp_idx[i]] = b;
            std::vector<double> v(sd + (size_t)m * b, sd + (size_t)m * (b + 1));
            std::sort(v.begin(), v.end());
            sdm[b] = (m & 1) ? v[m / 2] : 0.5 * (v[m / 2 - 1] + v[m / 2]);
            if (!(sdm[b] > 0.0)) return set_error(ICNV_E_BAD_ARG, "state sd must be positive (group %d)", b);
        }
        ICNV_CUDA(cudaMemcpyAsync(d_idx, grp_idx, sizeof(int32_t) * (size_t)n_idx, cudaMemcpyHostToDevice, st));
        ICNV_CUDA(cudaMemcpyAsync(d_idx + n_idx, grp_of.data(), sizeof(int32_t) * (size_t)C, cudaMemcpyHostToDevice, st));
        ICNV_CUDA(cudaMemcpyAsync(d_sd, sdm.data(), sizeof(double) * (size_t)n_grp, cudaMemcpyHostToDevice, st));
        if ((rc = dev_group_means(dX, G, G, d_idx, grp_off, n_grp, 0, d_xm, st))) return rc;
        rc = icnv_dev_viterbi_f64(d_xm, G, n_grp, chr_start, chr_len, K, m, Pi, delta, mean, d_sd, 1, d_st, d_mg, d_flag, st);
        if (rc) return rc;
        if ((rc = icnv_dev_scatter_group_states(d_st, G, C, d_idx + n_idx, d_out, st))) return rc;
        if (margins)
            ICNV_CUDA(cudaMemcpyAsync(margins, d_mg, sizeof(double) * (size_t)K * (size_t)n_grp, cudaMemcpyDeviceToHost, st));
        ICNV_CUDA(cudaStreamSynchronize(st));  // grp_of / sdm are stack-lifetime host buffers
    }
    if (states) {
        ICNV_CUDA(cudaMemcpyAsync(states, d_out, sizeof(int32_t) * (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    } else {
        uint8_t *d_n = (uint8_t *)scratch(SLOT_SLAB_ST0, (size_t)(G * C));
        if (!d_n) return ICNV_E_NOMEM;
        if ((rc = icnv_dev_narrow_states(d_out, d_n, G * C, st))) return rc;
        ICNV_CUDA(cudaMemcpyAsync(states8, d_n, (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    }
    return check_flag(d_flag, st);
}

int icnv_viterbi_f64(const double *X, int64_t G, int64_t C, const int32_t *chr_start, const int32_t *chr_len, int K,
                     const int32_t *grp_off, const int32_t *grp_idx, int n_grp, int m, const double *Pi,
                     const double *delta, const double *mean, const double *sd, int32_t *states, double *margins) {
    return viterbi_impl(X, G, C, chr_start, chr_len, K, grp_off, grp_idx, n_grp, m, Pi, delta, mean, sd, states, nullptr,
                        margins);
}

int icnv_viterbi_u8_f64(const double *X, int64_t G, int64_t C, const int32_t *chr_start, const int32_t *chr_len, int K,
                        const int32_t *grp_off, const int32_t *grp_idx, int n_grp, int m, const double *Pi,
                        const double *delta, const double *mean, const double *sd, uint8_t *states, double *margins) {
    return viterbi_impl(X, G, C, chr_start, chr_len, K, grp_off, grp_idx, n_grp, m, Pi, delta, mean, sd, nullptr, states,
                        margins);
}

int icnv_median_filter_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                           const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx, int n_grp,
                           int window_size) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_median_filter_f64: bad argument");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    if ((rc = validate_groups(C, grp_off, grp_idx, n_grp, true))) return rc;
    double *dX, *dY;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(G * C));
    int *d_flag = (int *)scratch(SLOT_MISC, 64);
    if (!dY || !d_flag) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    rc = icnv_dev_median_filter_f64(dX, dY, G, C, chr_start, chr_len, K, grp_off, grp_idx, n_grp, window_size, st);
    if (rc) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    return check_flag(d_flag, st);
}

// ---- gene filters and counts ingest (run() steps 2-3; R/inferCNV_ops.R:2124-2209, 3064-3111) ------------------------

// rowMeans as R computes it: a long-double quotient rounded to double (the sums of count data are exact integers)
static void means_from_sums(const double *sums, int64_t G, int64_t C, double *means) {
    for (int64_t g = 0; g < G; ++g) means[g] = (double)((long double)sums[g] / (long double)C);
}

static int download_gene_stats(const double *d_sums, const int32_t *d_npos, int64_t G, int64_t C, double *sums,
                               int32_t *n_pos, double *means, cudaStream_t st) {
    std::vector<double> tmp;
    double *hs = sums;
    if (!hs) {
        tmp.resize((size_t)G);
        hs = tmp.data();
    }
    ICNV_CUDA(cudaMemcpyAsync(hs, d_sums, sizeof(double) * (size_t)G, cudaMemcpyDeviceToHost, st));
    if (n_pos) ICNV_CUDA(cudaMemcpyAsync(n_pos, d_npos, sizeof(int32_t) * (size_t)G, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    if (means) means_from_sums(hs, G, C, means);
    return ICNV_OK;
}

int icnv_gene_stats_f64(const double *X, int64_t G, int64_t C, double *sums, int32_t *n_pos, double *means) {
    ICNV_HOST_PROLOGUE();
    if (!X || G <= 0 || C <= 0 || (!sums && !n_pos && !means)) return set_error(ICNV_E_BAD_ARG, "icnv_gene_stats_f64: bad argument");
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    double *d_sums = (double *)scratch(SLOT_MEANS, sizeof(double) * (size_t)G);
    int32_t *d_npos = (int32_t *)scratch(SLOT_IDX2, sizeof(int32_t) * (size_t)G);
    if (!d_sums || !d_npos) return ICNV_E_NOMEM;
    if ((rc = icnv_dev_gene_stats_f64(dX, G, G, C, d_sums, d_npos, st))) return rc;
    return download_gene_stats(d_sums, d_npos, G, C, sums, n_pos, means, st);
}

/* scale_infercnv_expr, R/inferCNV_ops.R:3174-3186 (run() step 5 when scale_data = TRUE): every gene centred and scaled
 * to unit sd (n - 1) across the cells. */
int icnv_scale_infercnv_expr_f64(const double *X, double *Y, int64_t G, int64_t C) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_scale_infercnv_expr_f64: bad argument");
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(G * C));
    double *d_stat = (double *)scratch(SLOT_MEANS, sizeof(double) * 2 * (size_t)G);
    if (!dY || !d_stat) return ICNV_E_NOMEM;
    if ((rc = icnv_dev_scale_rows_f64(dX, dY, G, C, d_stat, d_stat + G, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

static int validate_keep(const int32_t *keep, int64_t n_keep, int64_t G) {
    if (!keep || n_keep <= 0 || n_keep > G) return set_error(ICNV_E_BAD_ARG, "the list of genes to keep is empty or too long");
    for (int64_t i = 0; i < n_keep; ++i)
        if (keep[i] < 0 || keep[i] >= G || (i > 0 && keep[i] <= keep[i - 1]))
            return set_error(ICNV_E_BAD_ARG, "genes to keep must be increasing row indices in [0, G)");
    return ICNV_OK;
}

int icnv_remove_genes_f64(const double *X, int64_t G, int64_t C, const int32_t *keep, int64_t n_keep, double *Y) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_remove_genes_f64: bad argument");
    int rc = validate_keep(keep, n_keep, G);
    if (rc) return rc;
    double *dX;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(n_keep * C));
    int32_t *d_keep = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)n_keep);
    if (!dY || !d_keep) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemcpyAsync(d_keep, keep, sizeof(int32_t) * (size_t)n_keep, cudaMemcpyHostToDevice, st));
    if ((rc = icnv_dev_gather_rows_f64(dX, G, d_keep, n_keep, dY, C, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(n_keep * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

/* Row selection in any order (rows may repeat): Y (n x C) = X[idx, ].  The ingest step .order_reduce
 * (R/inferCNV.R:352-428) reorders the expression matrix to the genomic position table with it. */
int icnv_gather_genes_f64(const double *X, int64_t G, int64_t C, const int32_t *idx, int64_t n, double *Y) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || !idx || G <= 0 || C <= 0 || n <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_gather_genes_f64: bad argument");
    for (int64_t i = 0; i < n; ++i)
        if (idx[i] < 0 || idx[i] >= G) return set_error(ICNV_E_BAD_ARG, "row index %d out of range", idx[i]);
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(n * C));
    int32_t *d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)n);
    if (!dY || !d_idx) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemcpyAsync(d_idx, idx, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    if ((rc = icnv_dev_gather_rows_f64(dX, G, d_idx, n, dY, C, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(n * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

static int validate_csc(const int32_t *p, const int32_t *ri, const double *x, int64_t G, int64_t C) {
    if (!p || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "compressed-column matrix: bad argument");
    if (p[0] != 0) return set_error(ICNV_E_BAD_ARG, "compressed-column matrix: p[0] must be 0");
    for (int64_t c = 0; c < C; ++c)
        if (p[c + 1] < p[c]) return set_error(ICNV_E_BAD_ARG, "compressed-column matrix: p must be non-decreasing");
    const int64_t nnz = p[C];
    if (nnz > 0 && (!ri || !x)) return set_error(ICNV_E_BAD_ARG, "compressed-column matrix: i / x missing");
    for (int64_t k = 0; k < nnz; ++k)
        if (ri[k] < 0 || ri[k] >= G) return set_error(ICNV_E_BAD_ARG, "compressed-column matrix: row index out of range");
    return ICNV_OK;
}

// upload p / i / x into SLOT_IN (x first: 8-byte alignment)
static int upload_csc(const int32_t *p, const int32_t *ri, const double *x, int64_t C, const double **d_x,
                      const int32_t **d_i, const int32_t **d_p, cudaStream_t st) {
    const int64_t nnz = p[C];
    const size_t bx = sizeof(double) * (size_t)std::max<int64_t>(nnz, 1);
    const size_t bi = (sizeof(int32_t) * (size_t)std::max<int64_t>(nnz, 1) + 7) & ~(size_t)7;
    char *base = (char *)scratch(SLOT_IN, bx + bi + sizeof(int32_t) * (size_t)(C + 1));
    if (!base) return ICNV_E_NOMEM;
    if (nnz > 0) {
        ICNV_CUDA(cudaMemcpyAsync(base, x, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
        ICNV_CUDA(cudaMemcpyAsync(base + bx, ri, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    }
    ICNV_CUDA(cudaMemcpyAsync(base + bx + bi, p, sizeof(int32_t) * (size_t)(C + 1), cudaMemcpyHostToDevice, st));
    *d_x = (const double *)base;
    *d_i = (const int32_t *)(base + bx);
    *d_p = (const int32_t *)(base + bx + bi);
    return ICNV_OK;
}

int icnv_csc_gene_stats_f64(const int32_t *p, const int32_t *ri, const double *x, int64_t G, int64_t C, double *sums,
                            int32_t *n_pos, double *means) {
    ICNV_HOST_PROLOGUE();
    int rc = validate_csc(p, ri, x, G, C);
    if (rc) return rc;
    if (!sums && !n_pos && !means) return set_error(ICNV_E_BAD_ARG, "icnv_csc_gene_stats_f64: no output requested");
    const double *d_x;
    const int32_t *d_i, *d_p;
    if ((rc = upload_csc(p, ri, x, C, &d_x, &d_i, &d_p, st))) return rc;
    double *d_sums = (double *)scratch(SLOT_MEANS, sizeof(double) * (size_t)G);
    int32_t *d_npos = (int32_t *)scratch(SLOT_IDX2, sizeof(int32_t) * (size_t)G);
    if (!d_sums || !d_npos) return ICNV_E_NOMEM;
    if ((rc = icnv_dev_csc_gene_stats_f64(d_i, d_x, p[C], G, d_sums, d_npos, st))) return rc;
    return download_gene_stats(d_sums, d_npos, G, C, sums, n_pos, means, st);
}

int icnv_csc_normalize_f64(const int32_t *p, const int32_t *ri, const double *x, int64_t G, int64_t C, const int32_t *keep,
                           int64_t n_keep, double normalize_factor, double *Y, double *col_sums) {
    ICNV_HOST_PROLOGUE();
    int rc = validate_csc(p, ri, x, G, C);
    if (rc) return rc;
    if (!Y) return set_error(ICNV_E_BAD_ARG, "icnv_csc_normalize_f64: bad argument");
    const int64_t G_out = keep ? n_keep : G;
    if (keep && (rc = validate_keep(keep, n_keep, G))) return rc;
    const double *d_x;
    const int32_t *d_i, *d_p;
    if ((rc = upload_csc(p, ri, x, C, &d_x, &d_i, &d_p, st))) return rc;
    int32_t *d_map = nullptr;
    std::vector<int32_t> map;
    if (keep) {
        map.assign((size_t)G, -1);
        for (int64_t i = 0; i < n_keep; ++i) map[(size_t)keep[i]] = (int32_t)i;
        d_map = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)G);
        if (!d_map) return ICNV_E_NOMEM;
        ICNV_CUDA(cudaMemcpyAsync(d_map, map.data(), sizeof(int32_t) * (size_t)G, cudaMemcpyHostToDevice, st));
    }
    double *d_cs = (double *)scratch(SLOT_PARTIAL, sizeof(double) * (size_t)C);
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(G_out * C));
    if (!d_cs || !dY) return ICNV_E_NOMEM;
    if ((rc = icnv_dev_csc_col_sums_f64(d_p, d_i, d_x, d_map, C, d_cs, st))) return rc;
    std::vector<double> cs((size_t)C);
    ICNV_CUDA(cudaMemcpyAsync(cs.data(), d_cs, sizeof(double) * (size_t)C, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    if (col_sums) std::copy(cs.begin(), cs.end(), col_sums);
    if (!(normalize_factor >= 0.0)) {   // NA: median of the library sizes (ops.R:3095-3097)
        std::sort(cs.begin(), cs.end());
        normalize_factor = (C & 1) ? cs[(size_t)(C / 2)] : 0.5 * (cs[(size_t)(C / 2 - 1)] + cs[(size_t)(C / 2)]);
    }
    if (!std::isfinite(normalize_factor)) return set_error(ICNV_E_NONFINITE, "Error, normalize factor not estimated");
    if ((rc = icnv_dev_csc_expand_f64(d_p, d_i, d_x, d_map, G_out, C, d_cs, normalize_factor, dY, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(G_out * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

// ---- per-chromosome subcluster HMM (predict_CNV_via_HMM_on_tumor_subclusters_per_chr, HMM.R:412-487) --------------

int icnv_viterbi_per_chr_u8_f64(const double *X, int64_t G, int64_t C, const int32_t *chr_start, const int32_t *chr_len, int K,
                                const int32_t *chr_grp_off, const int32_t *grp_off, const int32_t *grp_idx, int m,
                                const double *Pi, const double *delta, const double *mean, const double *sd,
                                uint8_t *states) {
    ICNV_HOST_PROLOGUE();
    if (!X || !states || G <= 0 || C <= 0 || !Pi || !delta || !mean || !sd || !chr_grp_off)
        return set_error(ICNV_E_BAD_ARG, "icnv_viterbi_per_chr_u8_f64: bad argument");
    if (m != 6 && m != 3) return set_error(ICNV_E_BAD_ARG, "m must be 6 or 3");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    if (chr_grp_off[0] != 0) return set_error(ICNV_E_BAD_ARG, "chr_grp_off[0] must be 0");
    for (int k = 0; k < K; ++k)
        if (chr_grp_off[k + 1] < chr_grp_off[k]) return set_error(ICNV_E_BAD_ARG, "chr_grp_off must be non-decreasing");
    const int n_tot = chr_grp_off[K];
    if (n_tot <= 0) return set_error(ICNV_E_BAD_ARG, "no subcluster given for any chromosome");
    if ((rc = validate_groups(C, grp_off, grp_idx, n_tot, false))) return rc;
    // group of every cell on every chromosome, per-group median sd (HMM.R:1122), chromosome of every gene
    std::vector<int32_t> grp_of((size_t)K * (size_t)C, -1), chr_id((size_t)G, 0);
    std::vector<double> sdm((size_t)n_tot);
    for (int k = 0; k < K; ++k) {
        for (int32_t g = chr_start[k]; g < chr_start[k] + chr_len[k]; ++g) chr_id[(size_t)g] = k;
        for (int b = chr_grp_off[k]; b < chr_grp_off[k + 1]; ++b)
            for (int32_t i = grp_off[b]; i < grp_off[b + 1]; ++i) {
                int32_t &slot = grp_of[(size_t)k * (size_t)C + (size_t)grp_idx[i]];
                if (slot >= 0) return set_error(ICNV_E_BAD_ARG, "cell %d is in two subclusters of chromosome %d", grp_idx[i], k);
                slot = b;
            }
    }
    for (int b = 0; b < n_tot; ++b) {
        std::vector<double> v(sd + (size_t)m * b, sd + (size_t)m * (b + 1));
        std::sort(v.begin(), v.end());
        sdm[(size_t)b] = (m & 1) ? v[m / 2] : 0.5 * (v[m / 2 - 1] + v[m / 2]);
        if (!(sdm[(size_t)b] > 0.0)) return set_error(ICNV_E_BAD_ARG, "state sd must be positive (group %d)", b);
    }
    double *dX;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    const int64_t n_idx = grp_off[n_tot];
    int32_t *d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)(n_idx + K * C + G));
    double *d_xm = (double *)scratch(SLOT_MEANS, sizeof(double) * (size_t)G * (size_t)n_tot);
    double *d_sd = (double *)scratch(SLOT_BOUNDS, sizeof(double) * (size_t)n_tot);
    uint8_t *d_st = (uint8_t *)scratch(SLOT_STATES, (size_t)G * (size_t)n_tot);
    uint8_t *d_out = (uint8_t *)scratch(SLOT_SLAB_ST0, (size_t)(G * C));
    int *d_flag = (int *)scratch(SLOT_MISC, 64);
    if (!d_idx || !d_xm || !d_sd || !d_st || !d_out || !d_flag) return ICNV_E_NOMEM;
    int32_t *d_grp_of = d_idx + n_idx, *d_chr_id = d_grp_of + (int64_t)K * C;
    ICNV_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    ICNV_CUDA(cudaMemcpyAsync(d_idx, grp_idx, sizeof(int32_t) * (size_t)n_idx, cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaMemcpyAsync(d_grp_of, grp_of.data(), sizeof(int32_t) * grp_of.size(), cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaMemcpyAsync(d_chr_id, chr_id.data(), sizeof(int32_t) * (size_t)G, cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaMemcpyAsync(d_sd, sdm.data(), sizeof(double) * (size_t)n_tot, cudaMemcpyHostToDevice, st));
    // rowMeans of every (chromosome, subcluster) column over ALL genes - only its own chromosome's rows are used
    // below; the Viterbi of the other chromosomes' rows is redundant work on a G x n_tot matrix, small next to the
    // G x C upload
    if ((rc = dev_group_means(dX, G, G, d_idx, grp_off, n_tot, 0, d_xm, st))) return rc;
    rc = icnv_dev_viterbi_f64(d_xm, G, n_tot, chr_start, chr_len, K, m, Pi, delta, mean, d_sd, 1, d_st, nullptr, d_flag, st);
    if (rc) return rc;
    if ((rc = icnv_dev_scatter_states_per_chr_u8(d_st, G, C, d_chr_id, d_grp_of, d_out, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(states, d_out, (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    return check_flag(d_flag, st);   // synchronises: the host tables above outlive the copies
}

// ---- CNV region calling on the state matrix (R/inferCNV_HMM.R:706-1087) ----------------------------------------

static int check_state_flag(int *d_flag, cudaStream_t st) {
    int h = 0;
    ICNV_CUDA(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    if (h & 4) return set_error(ICNV_E_BAD_ARG, "state outside 0..6 / 255 in the state matrix");
    return ICNV_OK;
}

// upload the states and the cell lists, run the consensus into SLOT_RG_CONS
static int host_consensus(const uint8_t *states, int64_t G, int64_t C, const int32_t *grp_off, const int32_t *grp_idx,
                          int n_grp, uint8_t **d_cons_out, cudaStream_t st) {
    uint8_t *dS = (uint8_t *)scratch(SLOT_STATES, (size_t)(G * C));
    int32_t *d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)grp_off[n_grp]);
    uint8_t *d_cons = (uint8_t *)scratch(SLOT_RG_CONS, (size_t)G * (size_t)n_grp);
    int *d_flag = (int *)scratch(SLOT_MISC, 64);
    if (!dS || !d_idx || !d_cons || !d_flag) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemcpyAsync(dS, states, (size_t)(G * C), cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaMemcpyAsync(d_idx, grp_idx, sizeof(int32_t) * (size_t)grp_off[n_grp], cudaMemcpyHostToDevice, st));
    ICNV_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    int rc = icnv_dev_state_consensus_u8(dS, G, G, d_idx, grp_off, n_grp, d_cons, d_flag, st);
    if (rc) return rc;
    if ((rc = check_state_flag(d_flag, st))) return rc;
    *d_cons_out = d_cons;
    return ICNV_OK;
}

int icnv_state_consensus_u8(const uint8_t *states, int64_t G, int64_t C, const int32_t *grp_off, const int32_t *grp_idx,
                            int n_grp, uint8_t *consensus) {
    ICNV_HOST_PROLOGUE();
    if (!states || !consensus || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_state_consensus_u8: bad argument");
    int rc = validate_groups(C, grp_off, grp_idx, n_grp, false);
    if (rc) return rc;
    uint8_t *d_cons;
    if ((rc = host_consensus(states, G, C, grp_off, grp_idx, n_grp, &d_cons, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(consensus, d_cons, (size_t)G * (size_t)n_grp, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

/* get_predicted_CNV_regions(by = "subcluster") + the overwrite loop of HMM.R:472-483 as one pass */
int icnv_apply_state_consensus_u8(const uint8_t *states, int64_t G, int64_t C, const int32_t *chr_start,
                                  const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx, int n_grp,
                                  uint8_t *out) {
    ICNV_HOST_PROLOGUE();
    if (!states || !out || G <= 0 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_apply_state_consensus_u8: bad argument");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    if ((rc = validate_groups(C, grp_off, grp_idx, n_grp, false))) return rc;
    std::vector<int32_t> tab((size_t)(G + C), -1);     // chr_of[G] (-1: chromosome with < 2 genes) | grp_of[C]
    for (int k = 0; k < K; ++k)
        for (int32_t g = chr_start[k]; g < chr_start[k] + chr_len[k]; ++g) tab[(size_t)g] = chr_len[k] >= 2 ? k : -1;
    for (int b = 0; b < n_grp; ++b)
        for (int32_t i = grp_off[b]; i < grp_off[b + 1]; ++i) {
            if (tab[(size_t)G + (size_t)grp_idx[i]] >= 0) return set_error(ICNV_E_BAD_ARG, "cell %d is in two groups", grp_idx[i]);
            tab[(size_t)G + (size_t)grp_idx[i]] = b;
        }
    uint8_t *d_cons;
    if ((rc = host_consensus(states, G, C, grp_off, grp_idx, n_grp, &d_cons, st))) return rc;   // states now in SLOT_STATES
    const uint8_t *dS = (const uint8_t *)ctx().slot_ptr[SLOT_STATES];
    int32_t *d_tab = (int32_t *)scratch(SLOT_IDX2, sizeof(int32_t) * tab.size());
    uint8_t *d_out = (uint8_t *)scratch(SLOT_SLAB_ST0, (size_t)(G * C));
    if (!d_tab || !d_out) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemcpyAsync(d_tab, tab.data(), sizeof(int32_t) * tab.size(), cudaMemcpyHostToDevice, st));
    if ((rc = icnv_dev_apply_consensus_u8(dS, d_cons, G, C, d_tab, d_tab + G, d_out, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(out, d_out, (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

static int validate_gene_positions(const double *gene_start, const double *gene_stop, int64_t G) {
    if (!gene_start || !gene_stop) return set_error(ICNV_E_BAD_ARG, "gene start / stop positions missing");
    for (int64_t g = 0; g < G; ++g)
        if (!std::isfinite(gene_start[g]) || !std::isfinite(gene_stop[g]))
            return set_error(ICNV_E_NONFINITE, "non-finite gene position at gene %lld", (long long)g);
    return ICNV_OK;
}

int icnv_cnv_regions_u8(const uint8_t *seqs, int64_t G, int64_t n_seq, const int32_t *chr_start, const int32_t *chr_len,
                        int K, const double *gene_start, const double *gene_stop, int64_t *n_regions) {
    ICNV_HOST_PROLOGUE();
    if (!seqs || !n_regions || G <= 0 || n_seq <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_cnv_regions_u8: bad argument");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    if ((rc = validate_gene_positions(gene_start, gene_stop, G))) return rc;
    uint8_t *dS = (uint8_t *)scratch(SLOT_STATES, (size_t)(G * n_seq));
    if (!dS) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemcpyAsync(dS, seqs, (size_t)(G * n_seq), cudaMemcpyHostToDevice, st));
    return icnv_dev_cnv_regions_u8(dS, G, G, n_seq, nullptr, chr_start, chr_len, K, gene_start, gene_stop, n_regions, st);
}

int icnv_predicted_cnv_regions_u8(const uint8_t *states, int64_t G, int64_t C, const int32_t *chr_start,
                                  const int32_t *chr_len, int K, const double *gene_start, const double *gene_stop,
                                  const int32_t *grp_off, const int32_t *grp_idx, int n_grp, uint8_t *consensus,
                                  int64_t *n_regions) {
    ICNV_HOST_PROLOGUE();
    if (!states || !n_regions || G <= 0 || C <= 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_predicted_cnv_regions_u8: bad argument");
    int rc = validate_chr(G, chr_start, chr_len, K);
    if (rc) return rc;
    if ((rc = validate_groups(C, grp_off, grp_idx, n_grp, false))) return rc;
    if ((rc = validate_gene_positions(gene_start, gene_stop, G))) return rc;
    bool singletons = !consensus;       // by = "cell": every group is one cell, its consensus is its own column
    for (int k = 0; k < n_grp && singletons; ++k) singletons = grp_off[k + 1] - grp_off[k] == 1;
    if (singletons) {
        uint8_t *dS = (uint8_t *)scratch(SLOT_STATES, (size_t)(G * C));
        int32_t *d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)n_grp);
        if (!dS || !d_idx) return ICNV_E_NOMEM;
        ICNV_CUDA(cudaMemcpyAsync(dS, states, (size_t)(G * C), cudaMemcpyHostToDevice, st));
        ICNV_CUDA(cudaMemcpyAsync(d_idx, grp_idx, sizeof(int32_t) * (size_t)n_grp, cudaMemcpyHostToDevice, st));
        return icnv_dev_cnv_regions_u8(dS, G, G, n_grp, d_idx, chr_start, chr_len, K, gene_start, gene_stop, n_regions, st);
    }
    uint8_t *d_cons;
    if ((rc = host_consensus(states, G, C, grp_off, grp_idx, n_grp, &d_cons, st))) return rc;
    if (consensus) ICNV_CUDA(cudaMemcpyAsync(consensus, d_cons, (size_t)G * (size_t)n_grp, cudaMemcpyDeviceToHost, st));
    return icnv_dev_cnv_regions_u8(d_cons, G, G, n_grp, nullptr, chr_start, chr_len, K, gene_start, gene_stop, n_regions, st);
}

int icnv_cnv_regions_fetch(int64_t n_regions, int32_t *seq, int32_t *chr, int32_t *first_gene, int32_t *last_gene,
                           int32_t *state, double *start, double *end) {
    ICNV_HOST_PROLOGUE();
    int64_t n = 0;
    const int32_t *d_seq, *d_chr, *d_first, *d_last, *d_state;
    const double *d_start, *d_end;
    int rc = icnv_dev_cnv_regions_records(&n, &d_seq, &d_chr, &d_first, &d_last, &d_state, &d_start, &d_end);
    if (rc) return rc;
    if (n_regions != n)
        return set_error(ICNV_E_BAD_ARG, "icnv_cnv_regions_fetch: %lld records asked for, the last region call produced %lld",
                         (long long)n_regions, (long long)n);
    if (n == 0) return ICNV_OK;
    const size_t bi = sizeof(int32_t) * (size_t)n, bd = sizeof(double) * (size_t)n;
    if (seq) ICNV_CUDA(cudaMemcpyAsync(seq, d_seq, bi, cudaMemcpyDeviceToHost, st));
    if (chr) ICNV_CUDA(cudaMemcpyAsync(chr, d_chr, bi, cudaMemcpyDeviceToHost, st));
    if (first_gene) ICNV_CUDA(cudaMemcpyAsync(first_gene, d_first, bi, cudaMemcpyDeviceToHost, st));
    if (last_gene) ICNV_CUDA(cudaMemcpyAsync(last_gene, d_last, bi, cudaMemcpyDeviceToHost, st));
    if (state) ICNV_CUDA(cudaMemcpyAsync(state, d_state, bi, cudaMemcpyDeviceToHost, st));
    if (start) ICNV_CUDA(cudaMemcpyAsync(start, d_start, bd, cudaMemcpyDeviceToHost, st));
    if (end) ICNV_CUDA(cudaMemcpyAsync(end, d_end, bd, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

/* normalize_counts_by_seq_depth, R/inferCNV_ops.R:3064-3111: x / colSums * normalize_factor; a negative or NaN
 * factor means "median of the column sums" (the reference's NA default). */
int icnv_normalize_counts_by_seq_depth_f64(const double *X, double *Y, int64_t G, int64_t C, double normalize_factor) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || G <= 1 || C <= 0) return set_error(ICNV_E_BAD_ARG, "icnv_normalize_counts_by_seq_depth_f64: bad argument");
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(G * C));
    double *d_sums = (double *)scratch(SLOT_PARTIAL, sizeof(double) * (size_t)C);
    if (!dY || !d_sums) return ICNV_E_NOMEM;
    if ((rc = icnv_dev_column_stats_f64(dX, G, nullptr, C, d_sums, nullptr, st))) return rc;
    if (!(normalize_factor >= 0.0)) {  // median of the C column sums: a C-vector, selected on the host
        std::vector<double> cs((size_t)C);
        ICNV_CUDA(cudaMemcpyAsync(cs.data(), d_sums, sizeof(double) * (size_t)C, cudaMemcpyDeviceToHost, st));
        ICNV_CUDA(cudaStreamSynchronize(st));
        std::sort(cs.begin(), cs.end());
        normalize_factor = (C & 1) ? cs[C / 2] : 0.5 * (cs[C / 2 - 1] + cs[C / 2]);
    }
    if (!std::isfinite(normalize_factor)) return set_error(ICNV_E_NONFINITE, "Error, normalize factor not estimated");
    if ((rc = icnv_dev_scale_columns_f64(dX, dY, G, C, d_sums, normalize_factor, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

/* clear_noise_via_ref_mean_sd, R/inferCNV_ops.R:2302-2346 (noise_logistic = FALSE): mu = mean over all values of
 * the listed cells, s = sd_amplifier * mean over those cells of the per-cell sd; values strictly inside
 * (mu - s, mu + s) become mu.  idx = reference cells, or all observation cells when there are none. */
int icnv_clear_noise_via_ref_mean_sd_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *idx, int64_t n_idx,
                                         double sd_amplifier) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || !idx || G <= 1 || C <= 0 || n_idx <= 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_clear_noise_via_ref_mean_sd_f64: bad argument");
    for (int64_t i = 0; i < n_idx; ++i)
        if (idx[i] < 0 || idx[i] >= C) return set_error(ICNV_E_BAD_ARG, "cell index out of range");
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(G * C));
    int32_t *d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)n_idx);
    double *d_stats = (double *)scratch(SLOT_MEANS, sizeof(double) * 2 * (size_t)n_idx);
    if (!dY || !d_idx || !d_stats) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemcpyAsync(d_idx, idx, sizeof(int32_t) * (size_t)n_idx, cudaMemcpyHostToDevice, st));
    if ((rc = icnv_dev_column_stats_f64(dX, G, d_idx, n_idx, d_stats, d_stats + n_idx, st))) return rc;
    std::vector<double> h(2 * (size_t)n_idx);
    ICNV_CUDA(cudaMemcpyAsync(h.data(), d_stats, sizeof(double) * 2 * (size_t)n_idx, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    double tot = 0.0, sdsum = 0.0;   // n_idx-vectors: combined on the host in list order
    for (int64_t i = 0; i < n_idx; ++i) {
        tot += h[i];
        sdsum += h[n_idx + i];
    }
    const double mu = tot / ((double)G * (double)n_idx);
    const double s = (sdsum / (double)n_idx) * sd_amplifier;
    if ((rc = icnv_dev_clear_noise_f64(dX, dY, G * C, mu - s, mu + s, mu, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

// R's mean(): long-double sum, then one refinement pass (summary.c, real_mean)
static double r_mean_host(const double *x, int64_t n) {
    long double s = 0.0L;
    for (int64_t i = 0; i < n; ++i) s += x[i];
    s /= (long double)n;
    long double t = 0.0L;
    for (int64_t i = 0; i < n; ++i) t += (x[i] - s);
    s += t / (long double)n;
    return (double)s;
}

/* remove_outliers_norm / .remove_outliers_norm, R/inferCNV_ops.R:1969-2056 (run() step 16, prune_outliers): values
 * below lower_bound become lower_bound, values above upper_bound become upper_bound.  Both bounds given (not NaN):
 * hard thresholds.  Otherwise out_method "average_bound" (.get_average_bounds, ops.R:2734-2742): the mean over
 * the cells of each cell's smallest value, and of each cell's largest value.  bounds_out (may be NULL) receives the
 * two bounds used. */
int icnv_remove_outliers_norm_f64(const double *X, double *Y, int64_t G, int64_t C, double lower_bound, double upper_bound,
                                  double *bounds_out) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || G <= 0 || C <= 0)
        return set_error(ICNV_E_BAD_ARG, "Error, something is wrong with the data, either null or no rows or columns");
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(G * C));
    if (!dY) return ICNV_E_NOMEM;
    if (!(lower_bound == lower_bound) || !(upper_bound == upper_bound)) {
        double *d_mm = (double *)scratch(SLOT_PARTIAL, sizeof(double) * 2 * (size_t)C);
        if (!d_mm) return ICNV_E_NOMEM;
        if ((rc = icnv_dev_column_minmax_f64(dX, G, C, d_mm, d_mm + C, st))) return rc;
        std::vector<double> mm(2 * (size_t)C);
        ICNV_CUDA(cudaMemcpyAsync(mm.data(), d_mm, sizeof(double) * 2 * (size_t)C, cudaMemcpyDeviceToHost, st));
        ICNV_CUDA(cudaStreamSynchronize(st));
        lower_bound = r_mean_host(mm.data(), C);
        upper_bound = r_mean_host(mm.data() + C, C);
    }
    if (bounds_out) {
        bounds_out[0] = lower_bound;
        bounds_out[1] = upper_bound;
    }
    if ((rc = icnv_dev_clamp_bounds_f64(dX, dY, G * C, lower_bound, upper_bound, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

// mean over all values of the listed cells (idx == NULL: all C cells) and mean of the per-cell sds
static int host_ref_mean_sd(const double *dX, int64_t G, int64_t C, const int32_t *idx, int64_t n_idx, double *mean,
                            double *mean_sd, cudaStream_t st) {
    const int64_t n = idx ? n_idx : C;
    int32_t *d_idx = nullptr;
    if (idx) {
        for (int64_t i = 0; i < n_idx; ++i)
            if (idx[i] < 0 || idx[i] >= C) return set_error(ICNV_E_BAD_ARG, "cell index out of range");
        d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)n_idx);
        if (!d_idx) return ICNV_E_NOMEM;
        ICNV_CUDA(cudaMemcpyAsync(d_idx, idx, sizeof(int32_t) * (size_t)n_idx, cudaMemcpyHostToDevice, st));
    }
    double *d_stats = (double *)scratch(SLOT_MEANS, sizeof(double) * 2 * (size_t)n);
    if (!d_stats) return ICNV_E_NOMEM;
    int rc = icnv_dev_column_stats_f64(dX, G, d_idx, n, d_stats, d_stats + n, st);
    if (rc) return rc;
    std::vector<double> h(2 * (size_t)n);
    ICNV_CUDA(cudaMemcpyAsync(h.data(), d_stats, sizeof(double) * 2 * (size_t)n, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    double tot = 0.0, sdsum = 0.0;   // n-vectors: combined on the host in list order
    for (int64_t i = 0; i < n; ++i) {
        tot += h[(size_t)i];
        sdsum += h[(size_t)(n + i)];
    }
    *mean = tot / ((double)G * (double)n);
    *mean_sd = sdsum / (double)n;
    return ICNV_OK;
}

/* clear_noise / .clear_noise, R/inferCNV_ops.R:2232-2275 (run() step 22 with a numeric noise_filter): centre = mean
 * over all values of the listed cells (reference cells; idx == NULL / n_idx == 0: all data, ops.R:2243-2247).
 * noise_logistic == 0: values strictly inside (centre - threshold, centre + threshold) become centre;
 * noise_logistic != 0: depress_log_signal_midpt_val(centre, threshold), slope 20 (R/inferCNV_heatmap.R:2783-2810).
 * threshold == 0: the matrix is returned unchanged (ops.R:2236-2238). */
int icnv_clear_noise_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *idx, int64_t n_idx, double threshold,
                         int noise_logistic) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || G <= 1 || C <= 0 || n_idx < 0) return set_error(ICNV_E_BAD_ARG, "icnv_clear_noise_f64: bad argument");
    if (threshold == 0.0) {
        if (Y != X) memcpy(Y, X, sizeof(double) * (size_t)(G * C));
        return ICNV_OK;
    }
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(G * C));
    if (!dY) return ICNV_E_NOMEM;
    double mu, msd;
    if ((rc = host_ref_mean_sd(dX, G, C, n_idx > 0 ? idx : nullptr, n_idx, &mu, &msd, st))) return rc;
    if (noise_logistic)
        rc = icnv_dev_logistic_adj_f64(dX, dY, G * C, mu, threshold, 20.0, st);
    else
        rc = icnv_dev_clear_noise_f64(dX, dY, G * C, mu - threshold, mu + threshold, mu, st);
    if (rc) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

/* clear_noise_via_ref_mean_sd with noise_logistic = TRUE, R/inferCNV_ops.R:2325-2329: the logistic depression around
 * the reference mean with midpoint sd_amplifier * mean(per-cell sd). */
int icnv_clear_noise_via_ref_mean_sd_logistic_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *idx,
                                                  int64_t n_idx, double sd_amplifier) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || !idx || G <= 1 || C <= 0 || n_idx <= 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_clear_noise_via_ref_mean_sd_logistic_f64: bad argument");
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)(G * C));
    if (!dY) return ICNV_E_NOMEM;
    double mu, msd;
    if ((rc = host_ref_mean_sd(dX, G, C, idx, n_idx, &mu, &msd, st))) return rc;
    if ((rc = icnv_dev_logistic_adj_f64(dX, dY, G * C, mu, msd * sd_amplifier, 20.0, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)(G * C), cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

/* assign_HMM_states_to_proxy_expr_vals (HMM.R:1191-1206, m = 6) / i3HMM_assign_HMM_states_to_proxy_expr_vals
 * (i3HMM.R:405-417, m = 3) on a numeric state matrix of n entries. */
int icnv_assign_hmm_states_to_proxy_expr_vals_f64(const double *X, double *Y, int64_t n, int m) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || n <= 0 || (m != 6 && m != 3))
        return set_error(ICNV_E_BAD_ARG, "icnv_assign_hmm_states_to_proxy_expr_vals_f64: bad argument");
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, n, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)n);
    if (!dY) return ICNV_E_NOMEM;
    if ((rc = icnv_dev_proxy_vals_f64(dX, dY, n, m, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

static int host_elementwise(const double *X, double *Y, int64_t n, int op, double param) {
    ICNV_HOST_PROLOGUE();
    if (!X || !Y || n <= 0) return set_error(ICNV_E_BAD_ARG, "element-wise step: bad argument");
    double *dX;
    int rc;
    if ((rc = upload_matrix(X, n, &dX, SLOT_IN, st))) return rc;
    double *dY = (double *)scratch(SLOT_OUT, sizeof(double) * (size_t)n);
    int *d_flag = (int *)scratch(SLOT_MISC, 64);
    if (!dY || !d_flag) return ICNV_E_NOMEM;
    ICNV_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    if ((rc = icnv_dev_elementwise_f64(dX, dY, n, op, param, d_flag, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(Y, dY, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    return check_flag(d_flag, st);
}

int icnv_log2xplus1_f64(const double *X, double *Y, int64_t n) { return host_elementwise(X, Y, n, 0, 0.0); }
int icnv_invert_log2_f64(const double *X, double *Y, int64_t n) { return host_elementwise(X, Y, n, 1, 0.0); }
int icnv_apply_max_threshold_bounds_f64(const double *X, double *Y, int64_t n, double threshold) {
    if (!(threshold > 0.0)) return set_error(ICNV_E_BAD_ARG, "threshold must be positive");
    return host_elementwise(X, Y, n, 2, threshold);
}

int icnv_mean_sd_f64(const double *X, int64_t G, int64_t C, const int32_t *idx, int64_t n_idx, double *mu,
                     double *sigma) {
    ICNV_HOST_PROLOGUE();
    if (!X || !idx || !mu || !sigma || G <= 1 || C <= 0 || n_idx <= 0)
        return set_error(ICNV_E_BAD_ARG, "icnv_mean_sd_f64: bad argument");
    for (int64_t i = 0; i < n_idx; ++i)
        if (idx[i] < 0 || idx[i] >= C) return set_error(ICNV_E_BAD_ARG, "cell index out of range");
    // only the listed columns go to the device (the reference cells are a tenth of the matrix): runs of consecutive
    // indices as one copy each into a compact buffer, the statistics then run over its columns 0 .. n_idx-1
    int rc;
    double *dX = (double *)scratch(SLOT_IN, sizeof(double) * (size_t)G * (size_t)n_idx);
    int32_t *d_idx = (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)n_idx);
    double *d_stats = (double *)scratch(SLOT_MEANS, sizeof(double) * 2 * (size_t)n_idx);
    if (!dX || !d_idx || !d_stats) return ICNV_E_NOMEM;
    std::vector<int32_t> ident((size_t)n_idx);
    for (int64_t i = 0; i < n_idx;) {
        int64_t j = i + 1;
        while (j < n_idx && idx[j] == idx[j - 1] + 1) ++j;
        ICNV_CUDA(cudaMemcpyAsync(dX + G * i, X + G * (int64_t)idx[i], sizeof(double) * (size_t)G * (size_t)(j - i), cudaMemcpyHostToDevice, st));
        for (int64_t k = i; k < j; ++k) ident[(size_t)k] = (int32_t)k;
        i = j;
    }
    idx = ident.data();
    ICNV_CUDA(cudaMemcpyAsync(d_idx, idx, sizeof(int32_t) * (size_t)n_idx, cudaMemcpyHostToDevice, st));
    // per-cell sum and sd on the device; the n_idx-vectors are combined in list order (the same
    // combination the multi-GPU driver applies to all-gathered per-cell statistics: identical bits)
    if ((rc = icnv_dev_column_stats_f64(dX, G, d_idx, n_idx, d_stats, d_stats + n_idx, st))) return rc;
    std::vector<double> h(2 * (size_t)n_idx);
    ICNV_CUDA(cudaMemcpyAsync(h.data(), d_stats, sizeof(double) * 2 * (size_t)n_idx, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    icnv_combine_cell_stats(h.data(), h.data() + n_idx, n_idx, G, mu, sigma);
    return ICNV_OK;
}

int icnv_pairwise_dist_f64(const double *X, int64_t G, int64_t C, const int32_t *cells, int64_t n, double *out) {
    ICNV_HOST_PROLOGUE();
    if (!cells) n = C;
    if (!X || G < 0 || C <= 0 || n < 0 || (n >= 2 && !out)) return set_error(ICNV_E_BAD_ARG, "icnv_pairwise_dist_f64: bad argument");
    for (int64_t i = 0; cells && i < n; ++i)
        if (cells[i] < 0 || cells[i] >= C) return set_error(ICNV_E_BAD_ARG, "cell index out of range");
    if (n < 2) return ICNV_OK;
    const size_t n_out = (size_t)n * (size_t)(n - 1) / 2;
    double *dX = nullptr, *d_out = (double *)scratch(SLOT_OUT, sizeof(double) * n_out);
    int32_t *d_idx = cells ? (int32_t *)scratch(SLOT_IDX, sizeof(int32_t) * (size_t)n) : nullptr;
    int rc;
    if (!d_out || (cells && !d_idx)) return ICNV_E_NOMEM;
    if (G > 0 && (rc = upload_matrix(X, G * C, &dX, SLOT_IN, st))) return rc;
    if (G == 0) dX = d_out;   // never read
    if (cells) ICNV_CUDA(cudaMemcpyAsync(d_idx, cells, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    if ((rc = icnv_dev_pairwise_dist_f64(dX, G, G, d_idx, n, d_out, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

int icnv_pairwise_dist_rows_f64(const double *x, int64_t n, int64_t G, double *out) {
    ICNV_HOST_PROLOGUE();
    if (!x || G < 0 || n < 0 || (n >= 2 && !out)) return set_error(ICNV_E_BAD_ARG, "icnv_pairwise_dist_rows_f64: bad argument");
    if (n < 2) return ICNV_OK;
    const size_t n_out = (size_t)n * (size_t)(n - 1) / 2;
    double *dX = nullptr, *d_out = (double *)scratch(SLOT_OUT, sizeof(double) * n_out);
    int rc;
    if (!d_out) return ICNV_E_NOMEM;
    if (G > 0 && (rc = upload_matrix(x, G * n, &dX, SLOT_IN, st))) return rc;
    if (G == 0) dX = d_out;   // never read
    if ((rc = icnv_dev_pairwise_dist_rows_f64(dX, n, n, G, d_out, st))) return rc;
    ICNV_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, st));
    ICNV_CUDA(cudaStreamSynchronize(st));
    return ICNV_OK;
}

/* mu, sigma (n-1) over all values of n cells from the cells' own (sum, sd): sum of squares about mu of a cell =
 * sd_c^2 (G-1) + G (mean_c - mu)^2.  Plain host arithmetic on 2n numbers, in list order. */
void icnv_combine_cell_stats(const double *sums, const double *sds, int64_t n, int64_t G, double *mu, double *sigma) {
    double tot = 0.0;
    for (int64_t i = 0; i < n; ++i) tot += sums[i];
    const double N = (double)G * (double)n;
    const double m = tot / N;
    double ss = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double d = sums[i] / (double)G - m;
        ss += sds[i] * sds[i] * (double)(G - 1) + (double)G * d * d;
    }
    *mu = m;
    *sigma = std::sqrt(ss / (N - 1.0));
}

}  // extern "C"
