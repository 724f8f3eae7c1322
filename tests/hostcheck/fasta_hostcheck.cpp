// fasta_hostcheck.cpp -- TEST INFRASTRUCTURE ONLY (built by tests/test_fasta_hostlogic.py).
//
// Runs the chunk walk and the resolve step of biostrings_b200/csrc/bsgpu_fasta.cuh -- the very
// code the two FASTA kernels execute, one chunk per thread -- chunk after chunk on the CPU, so
// that the chunk-boundary logic can be checked against the reference's parser without a GPU.
// Nothing in the product links or loads this.
#include <cstring>
#include "../../biostrings_b200/csrc/bsgpu_fasta.cuh"

// out: at least BSGPU_FASTA_GUARD + n bytes; seg_start / desc_off: at least n + 1 entries.
// totals: kept, invalid, nrec, first_header, first_seqline.
extern "C" int fasta_hostcheck(const uint8_t *text, int64_t n, const int16_t *lkup, uint8_t *out,
			       int64_t *seg_start, int64_t *desc_off, int64_t *totals)
{
	std::vector<uint8_t> padded((size_t) ((n + 15) / 16 * 16 + 16), (uint8_t) '\n');
	if (n > 0)
		memcpy(padded.data(), text, (size_t) n);
	int64_t nchunks = (n + BSGPU_FASTA_CHUNK - 1) / BSGPU_FASTA_CHUNK;
	std::vector<BsgpuFastaSummary> sums((size_t) nchunks);
	std::vector<BsgpuFastaResolved> res;
	BsgpuFastaResolved none = {0, 0, BSGPU_FASTA_SEQ};

	for (int64_t c = 0; c < nchunks; c++) {
		int64_t a = c * BSGPU_FASTA_CHUNK, b = a + BSGPU_FASTA_CHUNK < n ? a + BSGPU_FASTA_CHUNK : n;

		fasta_walk_chunk<false>(padded.data(), n, a, b, lkup, &sums[(size_t) c], none, nullptr, nullptr, nullptr);
	}
	// the between-passes step exactly as bsgpu_subject_fasta() runs it: fold 64 chunks to one group
	// summary, resolve the groups, expand back to one state per chunk
	int64_t ngroups = (nchunks + BSGPU_FASTA_GROUP - 1) / BSGPU_FASTA_GROUP;
	std::vector<BsgpuFastaGroup> groups((size_t) ngroups);
	std::vector<BsgpuFastaResolved> bases;
	for (int64_t g = 0; g < ngroups; g++) {
		int64_t c0 = g * BSGPU_FASTA_GROUP, c1 = c0 + BSGPU_FASTA_GROUP < nchunks ? c0 + BSGPU_FASTA_GROUP : nchunks;

		fasta_fold_group(sums.data(), c0, c1, &groups[(size_t) g]);
	}
	BsgpuFastaTotals t = fasta_resolve_groups(groups, bases);
	res.resize((size_t) nchunks);
	for (int64_t g = 0; g < ngroups; g++) {
		int64_t c0 = g * BSGPU_FASTA_GROUP, c1 = c0 + BSGPU_FASTA_GROUP < nchunks ? c0 + BSGPU_FASTA_GROUP : nchunks;

		fasta_expand_group(sums.data(), c0, c1, bases[(size_t) g], res.data());
	}
	totals[0] = t.kept;
	totals[1] = t.invalid;
	totals[2] = t.nrec;
	totals[3] = t.first_header;
	totals[4] = t.first_seqline;
	if (t.first_seqline >= 0 && (t.first_header < 0 || t.first_seqline < t.first_header))
		return 1;       // sequence data in front of the first record: the caller raises
	for (int64_t c = 0; c < nchunks; c++) {
		int64_t a = c * BSGPU_FASTA_CHUNK, b = a + BSGPU_FASTA_CHUNK < n ? a + BSGPU_FASTA_CHUNK : n;

		fasta_walk_chunk<true>(padded.data(), n, a, b, lkup, nullptr, res[(size_t) c], out, seg_start, desc_off);
	}
	return 0;
}
