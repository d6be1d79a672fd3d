// emu_stubs.cpp - TEST INFRASTRUCTURE: entry points of the translation units the host emulation does not cover
// (the FP64 pipeline kernels: inline PTX, TMA, mbarriers).  They fail loudly, so an emulated test can never pass by
// silently skipping the real kernel.
#include <stdint.h>

#include "icnv_common.cuh"

#define UNSUPPORTED(name) return icnv::set_error(ICNV_E_UNSUPPORTED, name " is not part of the host emulation")

extern "C" {
int icnv_dev_group_partial_sums_f64(const double *, int64_t, int64_t, const int32_t *, int64_t, int, int, double *, void *) {
    UNSUPPORTED("icnv_dev_group_partial_sums_f64");
}
int icnv_dev_combine_partials_f64(const double *, int64_t, int64_t, int64_t, double *, void *) { UNSUPPORTED("icnv_dev_combine_partials_f64"); }
int icnv_dev_bounds_from_means_f64(const double *, int64_t, int, double *, double *, double *, void *) {
    UNSUPPORTED("icnv_dev_bounds_from_means_f64");
}
ICNV_API int icnv_debug_stats(unsigned long long *, int) { UNSUPPORTED("icnv_debug_stats"); }
int icnv_dev_invlog_finish_f64(double *, int64_t, void *) { UNSUPPORTED("icnv_dev_invlog_finish_f64"); }
int icnv_dev_cell_pipeline_f64(const double *, int64_t, int64_t, const int32_t *, int64_t, double *, int64_t, const int32_t *,
                               const int32_t *, int, int, const double *, const double *, const double *, double, int, int,
                               const double *, const double *, const double *, int, int *, void *) {
    UNSUPPORTED("icnv_dev_cell_pipeline_f64");
}
int icnv_dev_widen_states(const uint8_t *, int32_t *, int64_t, void *) { UNSUPPORTED("icnv_dev_widen_states"); }
int icnv_dev_narrow_states(const int32_t *, uint8_t *, int64_t, void *) { UNSUPPORTED("icnv_dev_narrow_states"); }
int icnv_dev_scatter_group_states(const uint8_t *, int64_t, int64_t, const int32_t *, int32_t *, void *) {
    UNSUPPORTED("icnv_dev_scatter_group_states");
}
int icnv_dev_viterbi_f64(const double *, int64_t, int64_t, const int32_t *, const int32_t *, int, int, const double *,
                         const double *, const double *, const double *, int, uint8_t *, double *, int *, void *) {
    UNSUPPORTED("icnv_dev_viterbi_f64");
}
int icnv_set_hmm_mode(int) { UNSUPPORTED("icnv_set_hmm_mode"); }
int64_t icnv_hmm_rerun_count(void) { return -1; }
int icnv_dev_median_filter_f64(const double *, double *, int64_t, int64_t, const int32_t *, const int32_t *, int, const int32_t *,
                               const int32_t *, int, int, void *) {
    UNSUPPORTED("icnv_dev_median_filter_f64");
}
int icnv_dev_synth_f64(double *, int64_t, int64_t, int64_t, int64_t, const int32_t *, const int32_t *, int, uint64_t, void *) {
    UNSUPPORTED("icnv_dev_synth_f64");
}
}
