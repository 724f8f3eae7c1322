// bsgpu_indels.cuh -- matching with indels on a non-preprocessed dictionary: the per-position core.
//
// Used by bsgpu_count_indels() (bsgpu.cu).  The core is plain host/device C++ and is checked on
// the CPU against the reference's own _match_pattern_indels (tests/hostcheck/indels_hostcheck.cpp,
// tests/test_indels_oracle.py); the engine around it on the device (tests/test_gpu_indels.py).
//
// The reference (src/match_pattern_indels.c:58-107) walks the subject once per pattern; at every
// position j0 whose letter matches some pattern letter it takes the FIRST such pattern position i0
// (byte2offset, src/utils.c:77-97), spends i0 edits on skipping the pattern's first i0 letters and
// aligns the rest of the pattern right behind j0 with a banded edit distance of at most
// k = max.mismatch - i0 (_nedit_for_Ploffset, src/lowlevel_matching.c:132-213: band 2k+1, letters
// outside the subject are mismatches, early bailout; k = 0: a plain mismatch count,
// _nmismatch_at_Pshift :66-89).  A position that stays within budget yields a candidate
// (start, width, nedit); the candidates then pass the sequential "best local match" filter
// (report_provisory_match, :38-56).  The candidate of a position depends on nothing but the pattern
// and the subject letters behind it -- one thread per (pattern, position) -- and the filter touches
// candidates only (bsgpu_indels_best_local below, for the host or one thread per pattern).
#pragma once
#include <stdint.h>

#ifndef BSGPU_HD
#ifdef __CUDACC__
#define BSGPU_HD __host__ __device__ __forceinline__
#else
#define BSGPU_HD inline
#endif
#endif

#define BSGPU_INDELS_MAX_K 15           // band 2k+1 <= 31 cells per row, kept in registers / local memory

struct BsgpuIndelCand {
	int32_t start;          // 1-based start of the candidate match in the subject
	int32_t width;
	int32_t nedit;
};

// the four bytewise match tables on raw code bytes (src/lowlevel_matching.c:26-56)
BSGPU_HD bool bsgpu_indels_match(uint32_t p, uint32_t s, int fixedP, int fixedS)
{
	if (fixedP && fixedS)
		return p == s;
	if (fixedP)
		return ((p & ~s) & 0xFFu) == 0;
	if (fixedS)
		return ((~p & s) & 0xFFu) == 0;
	return (p & s) != 0;
}

// first[y] = first pattern position whose letter matches subject byte y, or -1
// (_init_byte2offset_with_Chars_holder, src/utils.c:77-97); once per pattern
BSGPU_HD void bsgpu_indels_first_table(const uint8_t *P, int plen, int fixedP, int fixedS, int16_t first[256])
{
	for (int y = 0; y < 256; y++) {
		int at = -1;

		for (int n = 0; n < plen; n++)
			if (bsgpu_indels_match(P[n], (uint32_t) y, fixedP, fixedS)) {
				at = n;
				break;
			}
		first[y] = (int16_t) at;
	}
}

// one cell of the banded recurrence (PROPAGATE_NEDIT, src/lowlevel_matching.c:111-121)
BSGPU_HD int bsgpu_indels_cell(const int *curr, const int *prev, int B, int row_length, int mismatch)
{
	int nedit = prev[B] + mismatch;

	if (B - 1 >= 0 && curr[B - 1] + 1 < nedit)
		nedit = curr[B - 1] + 1;
	if (B + 1 < row_length && prev[B + 1] + 1 < nedit)
		nedit = prev[B + 1] + 1;
	return nedit;
}

// _nedit_for_Ploffset (src/lowlevel_matching.c:132-213) for P[0, plen) placed at subject offset
// `off` with at most k >= 1 edits: returns the minimal number of edits (> k: no match), *min_width =
// length of the shortest subject substring starting at `off` that achieves it.
BSGPU_HD int bsgpu_indels_nedit(const uint8_t *P, int plen, const uint8_t *S, int64_t slen, int64_t off,
				int k, int fixedP, int fixedS, int *min_width)
{
	const int INF = 1 << 28;
	int rows[2][2 * BSGPU_INDELS_MAX_K + 1];
	int *curr = rows[0], *prev = rows[1];

	*min_width = 0;
	if (plen == 0)
		return 0;
	const int k_plus1 = k + 1;
	if (k > plen)
		k = plen;
	const int row_length = 2 * k + 1;
	for (int B = 0; B < row_length; B++)
		curr[B] = prev[B] = INF;
	int64_t min_si = off;
	int a, pi, min_nedit;

	// stage 0: the empty pattern prefix against 0..k subject letters
	for (int B = k, b = 0; B < row_length; B++, b++)
		curr[B] = b;
	// stage 1: pattern prefixes shorter than k; the band still widens to the left
	for (a = 1, pi = 0; a < k; a++, pi++) {
		uint32_t pc = P[pi];
		int *t = prev; prev = curr; curr = t;
		int B = k - a;
		int64_t si = min_si;

		curr[B++] = a;
		for (; B < row_length; B++, si++)
			curr[B] = bsgpu_indels_cell(curr, prev, B, row_length,
				si < 0 || si >= slen || !bsgpu_indels_match(pc, S[si], fixedP, fixedS));
	}
	// stage 2: the first full-width row
	{
		uint32_t pc = P[pi];
		int *t = prev; prev = curr; curr = t;
		int64_t si = min_si;

		curr[0] = min_nedit = a;
		for (int B = 1; B < row_length; B++, si++) {
			curr[B] = bsgpu_indels_cell(curr, prev, B, row_length,
				si < 0 || si >= slen || !bsgpu_indels_match(pc, S[si], fixedP, fixedS));
			if (curr[B] < min_nedit) {
				min_nedit = curr[B];
				*min_width = (int) (si - off + 1);
			}
		}
		a++;
		pi++;
	}
	// stage 3: the band slides one subject letter per pattern letter; bail out once every cell
	// of a row is over the budget
	for (; pi < plen; pi++, a++, min_si++) {
		uint32_t pc = P[pi];
		int *t = prev; prev = curr; curr = t;
		int64_t si = min_si;

		min_nedit = a;
		*min_width = 0;
		for (int B = 0; B < row_length; B++, si++) {
			curr[B] = bsgpu_indels_cell(curr, prev, B, row_length,
				si < 0 || si >= slen || !bsgpu_indels_match(pc, S[si], fixedP, fixedS));
			if (curr[B] < min_nedit) {
				min_nedit = curr[B];
				*min_width = (int) (si - off + 1);
			}
		}
		if (min_nedit >= k_plus1)
			break;
	}
	return min_nedit;
}

// The candidate of subject position j0 (0-based) for pattern P, or false.  Requires
// plen > max_nmis (shorter patterns go through the plain mismatch loop, src/match_pattern.c:98-99)
// and max_nmis <= BSGPU_INDELS_MAX_K.
BSGPU_HD bool bsgpu_indels_candidate(const uint8_t *P, int plen, const int16_t first[256], const uint8_t *S,
				     int64_t slen, int64_t j0, int max_nmis, int fixedP, int fixedS,
				     BsgpuIndelCand *out)
{
	int i0 = first[S[j0]];

	if (i0 < 0)
		return false;
	int k = max_nmis - i0;

	if (k < 0)
		return false;
	const uint8_t *P1 = P + i0 + 1;
	int p1len = plen - i0 - 1, nedit1, width1;

	if (k == 0) {
		// _nmismatch_at_Pshift with max_nmis = 0: any mismatch (or a letter outside S) rejects
		nedit1 = 0;
		for (int i = 0; i < p1len; i++) {
			int64_t j = j0 + 1 + i;

			if (j < 0 || j >= slen || !bsgpu_indels_match(P1[i], S[j], fixedP, fixedS)) {
				nedit1 = 1;
				break;
			}
		}
		width1 = p1len;
	} else {
		nedit1 = bsgpu_indels_nedit(P1, p1len, S, slen, j0 + 1, k, fixedP, fixedS, &width1);
	}
	if (nedit1 > k)
		return false;
	out->start = (int32_t) (j0 + 1);
	out->width = width1 + 1;
	out->nedit = nedit1 + i0;
	return true;
}

// report_provisory_match + the final flush (src/match_pattern_indels.c:38-56,103-105) over the
// candidates of ONE pattern in ascending start order: the number of best local matches; their
// (start, width) go to starts/widths when not NULL.
BSGPU_HD int64_t bsgpu_indels_best_local(const BsgpuIndelCand *cands, int64_t n, int32_t *starts, int32_t *widths)
{
	int64_t nout = 0;
	bool have = false;
	int32_t p_start = 0, p_end = 0, p_width = 0, p_nedit = 0;

	for (int64_t i = 0; i < n; i++) {
		int32_t end = cands[i].start + cands[i].width - 1;

		if (have) {
			if (end > p_end) {
				if (starts) starts[nout] = p_start;
				if (widths) widths[nout] = p_width;
				nout++;
			} else if (cands[i].nedit > p_nedit) {
				continue;
			}
		}
		p_start = cands[i].start;
		p_end = end;
		p_width = cands[i].width;
		p_nedit = cands[i].nedit;
		have = true;
	}
	if (have) {
		if (starts) starts[nout] = p_start;
		if (widths) widths[nout] = p_width;
		nout++;
	}
	return nout;
}

#include <algorithm>
#include <vector>

// Host step of the engine: candidates as the kernel emits them -- key = pattern << pos_bits | concat
// index of the candidate's first letter, val = width << 8 | nedit, in any order -- to counts.  The
// filter restarts with every (pattern, sequence): each view / set element is an independent subject
// (src/match_pdict.c:267-293,348-384).  seg_start: concat index of the first letter of every
// sequence, ascending.  counts: M entries, or M x N column-major when per_sequence.
static inline void bsgpu_indels_counts_from_candidates(const unsigned long long *keys, const uint32_t *vals,
		size_t n, int pos_bits, const std::vector<int64_t> &seg_start, int32_t M, bool per_sequence,
		int32_t *counts)
{
	const unsigned long long pos_mask = (1ull << pos_bits) - 1;
	std::vector<size_t> order(n);
	std::vector<BsgpuIndelCand> run;
	int64_t run_pid = -1, run_seg = -1;

	for (size_t i = 0; i < n; i++)
		order[i] = i;
	std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
	auto flush = [&]() {
		if (run.empty())
			return;
		int64_t k = bsgpu_indels_best_local(run.data(), (int64_t) run.size(), nullptr, nullptr);

		counts[per_sequence ? run_pid + run_seg * (int64_t) M : run_pid] += (int32_t) k;
		run.clear();
	};
	for (size_t q = 0; q < n; q++) {
		size_t i = order[q];
		int64_t pid = (int64_t) (keys[i] >> pos_bits), pos = (int64_t) (keys[i] & pos_mask);
		int64_t seg = std::upper_bound(seg_start.begin(), seg_start.end(), pos) - seg_start.begin() - 1;

		if (pid != run_pid || seg != run_seg) {
			flush();
			run_pid = pid;
			run_seg = seg;
		}
		BsgpuIndelCand c = {(int32_t) (pos - seg_start[(size_t) seg] + 1), (int32_t) (vals[i] >> 8),
				    (int32_t) (vals[i] & 0xFFu)};
		run.push_back(c);
	}
	flush();
}
