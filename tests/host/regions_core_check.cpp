// Host-side check of infercnv_b200/csrc/icnv_regions_core.h (the per-element decisions of the CNV region kernels,
// compiled here as plain C++): packed byte counters against plain counting incl. the flush-at-255 rule, the modal
// slot against table()/order(decreasing=TRUE)[1] semantics (ties -> smallest state, 255 = -1 first), and the
// region-opening predicate against a literal walk of .define_cnv_gene_regions (R/inferCNV_HMM.R:1006-1058).
#include <cstdio>
#include <random>
#include <vector>

#include "icnv_regions_core.h"

using namespace icnv;

int main() {
    std::mt19937_64 rng(7);
    int bad = 0;
    // slots
    for (unsigned v = 0; v < 256; ++v) {
        const int s = rg_slot(v);
        const bool valid = v <= 6 || v == 255;
        if (valid != (s >= 0)) ++bad;
        if (valid && rg_state_of_slot(s) != (uint8_t)v) ++bad;
    }
    if (rg_slot(255) != 0 || rg_slot(0) != 1 || rg_slot(6) != 7) ++bad;
    // packed counters + modal slot
    for (int t = 0; t < 20000; ++t) {
        const int n = 1 + (int)(rng() % 700);
        uint32_t plain[RG_SLOTS] = {0}, tot[RG_SLOTS] = {0};
        uint64_t acc = 0;
        int pending = 0;
        const int skew = (int)(rng() % 8);
        for (int i = 0; i < n; ++i) {
            int s = (int)(rng() % RG_SLOTS);
            if (rng() % 3 == 0) s = skew;
            plain[s]++;
            acc += rg_packed_one(s);
            if (++pending == 255 || i + 1 == n) {
                for (int k = 0; k < RG_SLOTS; ++k) tot[k] += rg_packed_get(acc, k);
                acc = 0;
                pending = 0;
            }
        }
        for (int k = 0; k < RG_SLOTS; ++k)
            if (plain[k] != tot[k]) ++bad;
        if (t % 4 == 0) {   // force ties
            const int a = (int)(rng() % RG_SLOTS), b = (int)(rng() % RG_SLOTS);
            tot[a] = tot[b] = 100000;
        }
        int want = 0;      // first of the largest counts in increasing state order
        for (int k = 0; k < RG_SLOTS; ++k)
            if (tot[k] > tot[want]) want = k;
        if (rg_argmax_first(tot) != want) ++bad;
        for (int k = 0; k < rg_argmax_first(tot); ++k)
            if (tot[k] >= tot[rg_argmax_first(tot)]) ++bad;
    }
    // region-opening predicate vs the literal loop
    for (int t = 0; t < 2000; ++t) {
        const int K = 1 + (int)(rng() % 6);
        std::vector<int> chr_of;
        std::vector<int> lens;
        for (int k = 0; k < K; ++k) {
            const int len = (int)(rng() % 4 == 0 ? 1 : 1 + rng() % 40);
            lens.push_back(len);
            for (int i = 0; i < len; ++i) chr_of.push_back(len >= 2 ? k : -1);
        }
        const int G = (int)chr_of.size();
        std::vector<unsigned> s((size_t)G);
        for (int g = 0; g < G; ++g) s[(size_t)g] = (g > 0 && rng() % 3) ? s[(size_t)g - 1] : (unsigned)(rng() % 7);
        std::vector<char> want((size_t)G, 0);
        int pos = 0;
        for (int k = 0; k < K; ++k) {
            if (lens[(size_t)k] >= 2) {
                want[(size_t)pos] = 1;
                for (int i = 1; i < lens[(size_t)k]; ++i)
                    if (s[(size_t)(pos + i)] != s[(size_t)(pos + i - 1)]) want[(size_t)(pos + i)] = 1;
            }
            pos += lens[(size_t)k];
        }
        for (int g = 0; g < G; ++g) {
            const bool got = rg_opens_region(g, g ? chr_of[(size_t)g - 1] : -1, chr_of[(size_t)g], g ? s[(size_t)g - 1] : 0u,
                                             s[(size_t)g]);
            if (got != (bool)want[(size_t)g]) ++bad;
        }
    }
    printf("mismatches %d\n", bad);
    return bad != 0;
}
