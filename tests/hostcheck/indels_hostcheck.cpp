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

// The whole engine as bsgpu_count_indels() runs it, minus the device: candidates of every (letter,
// pattern) of a multi-sequence subject laid out like bsgpu_subject's concat buffer (guard, sequences
// separated by one 0x80 byte), emitted in a scrambled order as atomics would, then the host step.
extern "C" int indels_engine_hostcheck(const uint8_t *concat, const int32_t *widths, int32_t npat,
				       const uint8_t *seqs, const int32_t *seq_len, int32_t nseq, int max_nmis,
				       int fixedP, int fixedS, int per_sequence, int32_t *counts)
{
	const int POS_BITS = 34, GUARD = 128;
	std::vector<int64_t> seg_start((size_t) nseq);
	std::vector<uint8_t> buf;
	std::vector<int64_t> pat_off((size_t) npat + 1, 0);

	buf.assign(GUARD, 0x80);
	size_t soff = 0;
	for (int32_t g = 0; g < nseq; g++) {
		seg_start[(size_t) g] = (int64_t) buf.size();
		buf.insert(buf.end(), seqs + soff, seqs + soff + seq_len[g]);
		buf.push_back(0x80);
		soff += (size_t) seq_len[g];
	}
	for (int32_t i = 0; i < npat; i++)
		pat_off[(size_t) i + 1] = pat_off[(size_t) i] + widths[i];
	std::vector<unsigned long long> keys;
	std::vector<uint32_t> vals;
	for (int32_t pid = npat - 1; pid >= 0; pid--) {         // patterns and letters in reverse: any order must do
		const uint8_t *P = concat + pat_off[(size_t) pid];
		int plen = widths[pid];
		int16_t first[256];

		if (plen <= max_nmis)
			return -1;
		bsgpu_indels_first_table(P, plen, fixedP, fixedS, first);
		for (int32_t g = nseq - 1; g >= 0; g--)
			for (int64_t j0 = seq_len[g] - 1; j0 >= 0; j0--) {
				BsgpuIndelCand c;

				if (!bsgpu_indels_candidate(P, plen, first, buf.data() + seg_start[(size_t) g], seq_len[g], j0,
							    max_nmis, fixedP, fixedS, &c))
					continue;
				keys.push_back(((unsigned long long) pid << POS_BITS) |
					       (unsigned long long) (seg_start[(size_t) g] + j0));
				vals.push_back(((uint32_t) c.width << 8) | (uint32_t) c.nedit);
			}
	}
	int64_t ncounts = (int64_t) npat * (per_sequence ? nseq : 1);
	for (int64_t i = 0; i < ncounts; i++)
		counts[i] = 0;
	bsgpu_indels_counts_from_candidates(keys.data(), vals.data(), keys.size(), POS_BITS, seg_start, npat,
					    per_sequence != 0, counts);
	return 0;
}
