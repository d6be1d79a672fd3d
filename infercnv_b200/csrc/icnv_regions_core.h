// icnv_regions_core.h - the per-element decisions of CNV region calling, shared by the kernels of
// icnv_regions.cu and by a host-compiled test (tests/host/regions_core_test.cpp).
//
// Reference semantics (R/inferCNV_HMM.R):
//   .get_state_consensus  :977-988   table(x) lists the distinct states in increasing order and
//                                     order(t, decreasing=TRUE)[1] takes the first of the largest counts, so the
//                                     modal state with ties going to the SMALLEST state; a cell the HMM left at
//                                     -1 (HMM.R:296) counts as the smallest value.
//   .define_cnv_gene_regions :1006-1058  a region starts at the first gene of every chromosome with >= 2 genes
//                                     and wherever the state differs from the previous gene's.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ICNV_HD __host__ __device__ __forceinline__
#else
#define ICNV_HD inline
#endif

namespace icnv {

constexpr int RG_SLOTS = 8;            // slot 0: unassigned (wire value 255 = R's -1); slot v+1: state v, v = 0..6
constexpr unsigned RG_UNASSIGNED = 255u;

// slot of a wire-format state, or -1 for a value outside {0..6, 255}
ICNV_HD int rg_slot(unsigned v) { return v == RG_UNASSIGNED ? 0 : (v <= 6u ? (int)v + 1 : -1); }
ICNV_HD uint8_t rg_state_of_slot(int s) { return s == 0 ? (uint8_t)RG_UNASSIGNED : (uint8_t)(s - 1); }

// one byte-wide counter per slot, packed in 64 bits: + (1 << 8*slot).  Flush before 256 additions.
ICNV_HD uint64_t rg_packed_one(int slot) { return 1ull << (8 * slot); }
ICNV_HD unsigned rg_packed_get(uint64_t acc, int slot) { return (unsigned)((acc >> (8 * slot)) & 0xffull); }

// modal slot, ties to the lowest slot (= smallest state)
ICNV_HD int rg_argmax_first(const uint32_t *c) {
    int best = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 1; s < RG_SLOTS; ++s)
        if (c[s] > c[best]) best = s;
    return best;
}

// does gene g open a region?  chr_* = chromosome id of the gene (-1: chromosome with < 2 genes, never reported),
// s_* = state; *_prev belong to gene g-1 and are ignored for g == 0.
ICNV_HD bool rg_opens_region(int64_t g, int chr_prev, int chr_cur, unsigned s_prev, unsigned s_cur) {
    return chr_cur >= 0 && (g == 0 || chr_prev != chr_cur || s_prev != s_cur);
}

}  // namespace icnv
