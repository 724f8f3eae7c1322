// indels_hostcheck.cpp -- TEST INFRASTRUCTURE ONLY (built by tests/test_indels_oracle.py).
//
// Runs the per-position core of biostrings_b200/csrc/bsgpu_indels.cuh on the CPU the way a kernel
// will run it -- every (pattern, position) independently, then the best-local-match filter per
// pattern -- so that it can be checked against the reference's _match_pattern_indels before any
// kernel exists.  Nothing in the product links or loads this.
#include <vector>
#include "../../biostrings_b200/csrc/bsgpu_indels.cuh"

// counts[i] = number of matches of pattern i (concat + widths) in S with at most max_nmis edits.
// Patterns not longer than max_nmis are outside the indels algorithm (src/match_pattern.c:98-99):
// counts[i] = -1 for them.
extern "C" int indels_hostcheck(const uint8_t *concat, const int32_t *widths, int32_t npat, const uint8_t *S,
				int64_t slen, int max_nmis, int fixedP, int fixedS, int64_t *counts)
{
	if (max_nmis < 1 || max_nmis > BSGPU_INDELS_MAX_K)
		return -1;
	size_t off = 0;
	for (int32_t i = 0; i < npat; i++) {
		const uint8_t *P = concat + off;
		int plen = widths[i];

		off += (size_t) plen;
		if (plen <= max_nmis) {
			counts[i] = -1;
			continue;
		}
		if (max_nmis < plen - slen) {           // src/match_pattern.c:93-95
			counts[i] = 0;
			continue;
		}
		int16_t first[256];
		std::vector<BsgpuIndelCand> cands;
		BsgpuIndelCand c;

		bsgpu_indels_first_table(P, plen, fixedP, fixedS, first);
		for (int64_t j0 = 0; j0 < slen; j0++)
			if (bsgpu_indels_candidate(P, plen, first, S, slen, j0, max_nmis, fixedP, fixedS, &c))
				cands.push_back(c);
		counts[i] = bsgpu_indels_best_local(cands.data(), (int64_t) cands.size(), nullptr, nullptr);
	}
	return 0;
}
