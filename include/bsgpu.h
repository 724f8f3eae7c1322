/*
 * bsgpu.h -- C ABI of the B200-native PDict matching path (libbsgpu.so).
 *
 * This is the drop-in boundary for ONE path of Bioconductor/Biostrings: PDict
 * preprocessing + matchPDict / countPDict / whichPDict / vcountPDict / vwhichPDict.
 * Each entry point names the reference interface it replaces (file:line under the
 * Biostrings source tree).  Plain C: pointers and sizes only, no R types, no longjmp
 * across the boundary, no torch types.  A C dispatch layer (the role src/match_pdict.c
 * plays in the reference) turns return codes into R's error() with bsgpu_last_error(),
 * whose text is the reference's own message for the same condition.
 *
 * Conventions
 *   - Sequences are Biostrings-encoded bytes (A=1 C=2 G=4 T=8, IUPAC = OR of those,
 *     '-'=16 '+'=32 '.'=64; R/XStringCodec-class.R:114,129-139), exactly what
 *     Chars_holder.ptr points at in the reference (src/XStringSet_class.c:37-51).
 *   - A set of sequences is (concat bytes, int32 widths[n]); a NULL widths pointer for
 *     head/tail means "no head"/"no tail" (pdict_head == R_NilValue, src/match_pdict.c:28).
 *   - NA is INT32_MIN (R's NA_INTEGER).  Pattern ids in results are 1-based where the
 *     reference's are (which), 0-based array positions otherwise.
 *   - Return value: 0 = OK, negative = error (BSGPU_E*); message in bsgpu_last_error().
 *   - One caller thread at a time per process (the reference entry points are
 *     non-reentrant too: static buffers in src/match_reporting.c:41,262).
 *   - There is no CPU fallback: without a usable CUDA device every compute entry
 *     point fails with BSGPU_ENODEVICE.
 */
#ifndef BSGPU_H
#define BSGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSGPU_NA INT32_MIN

enum {
	BSGPU_OK = 0,
	BSGPU_EINVAL = -1,	/* argument error; text mirrors the reference's error() */
	BSGPU_ENODEVICE = -2,	/* no CUDA device / driver */
	BSGPU_ECUDA = -3,	/* CUDA runtime failure */
	BSGPU_ENOMEM = -4,
	BSGPU_EUNSUPPORTED = -5
};

/* "matches_as" of the reference (src/match_reporting.c:10-31) */
enum {
	BSGPU_MATCHES_AS_NULL = 0,
	BSGPU_MATCHES_AS_WHICH = 1,
	BSGPU_MATCHES_AS_COUNTS = 2,
	BSGPU_MATCHES_AS_ENDS = 4
};

enum { BSGPU_ALGO_ACTREE2 = 0, BSGPU_ALGO_TWOBIT = 1 };

const char *bsgpu_last_error(void);
const char *bsgpu_version(void);

/* Number of CUDA devices (0 if none / no driver).  Never fails. */
int bsgpu_device_count(void);
/* Select the device used by this process (default: BSGPU_DEVICE env var, else 0). */
int bsgpu_set_device(int device);

/* ------------------------------------------------------------------------------------
 * Preprocessing: one Trusted Band + its head and tail = the reference's PDict3Parts
 * (R/PDict-class.R:174-233).  Replaces
 *     build_Twobit(tb, pp_exclude, base_codes)            src/match_pdict_Twobit.c:104
 *     ACtree2_build(tb, pp_exclude, base_codes, ...)      src/match_pdict_ACtree2.c:792
 * and the per-call _new_HeadTail()                        src/match_pdict_utils.c:347.
 * The result is a device-resident index (2-bit TB keys in a bucketed hash table, TB
 * duplicate groups in CSR form = low2high, head/tail bytes) and `high2low_out[M]`, the
 * same vector the reference returns (NA, or 1-based id of the first pattern with the
 * same TB).  `pp_exclude` (may be NULL): patterns with a non-NA entry are skipped
 * entirely, as in the reference.  Errors carry the reference's text: empty TB, TB
 * width > 14 for Twobit, unequal TB widths, non-base letter in the TB.
 * ------------------------------------------------------------------------------------ */
typedef struct bsgpu_pdict3parts bsgpu_pdict3parts;

int bsgpu_pdict3parts_build(int algo, int32_t M,
		const uint8_t *tb_concat, const int32_t *tb_widths,
		const uint8_t *head_concat, const int32_t *head_widths,
		const uint8_t *tail_concat, const int32_t *tail_widths,
		const int32_t *pp_exclude, int32_t *high2low_out,
		bsgpu_pdict3parts **out);
/* Non-preprocessed dictionary: `pattern` is a plain XStringSet (any letters, any widths, no
 * Trusted Band).  Replaces the `pattern` argument of match_XStringSet_XString,
 * match_XStringSet_XStringViews and vmatch_XStringSet_XStringSet (src/match_pdict.c:174,267,519),
 * which run one matchPattern() per pattern; all of the reference's algorithms report what its
 * naive inexact one does (src/match_pattern.c:53-76), and so does this.  The handle is passed to
 * bsgpu_match() / bsgpu_vmatch() (nparts = 1) like a PDict3Parts; with.indels goes through
 * bsgpu_count_indels().  A dictionary of >= 256 base-only patterns of >= 8 letters that is matched
 * exactly (max.mismatch 0, subject taken literally, no open-ended chunk) gets a Trusted-Band
 * index of its own on first use and is answered by the preprocessed engines -- same counts,
 * which-indices and ends as the matchPattern() loop, without its M x L cost (BSGPU_PLAIN_INDEX=0
 * switches this off, BSGPU_PLAIN_INDEX_MIN moves the threshold).
 * Errors: "empty patterns are not supported", "patterns with more than 20000 letters are not
 * supported" (R/match-utils.R:47-52). */
int bsgpu_patterns_build(const uint8_t *concat, const int32_t *widths, int32_t M,
		bsgpu_pdict3parts **out);

/* Replaces the head / tail of an existing index (both may be NULL = absent).  The reference
 * passes pdict_head / pdict_tail with every match call (src/match_pdict.c:153-171) while the
 * Trusted Band is preprocessed once; a C dispatch that built the index in build_Twobit /
 * ACtree2_build attaches the flanks here on the first match call. */
int bsgpu_pdict3parts_set_flanks(bsgpu_pdict3parts *p3,
		const uint8_t *head_concat, const int32_t *head_widths,
		const uint8_t *tail_concat, const int32_t *tail_widths);
/* R/PDict-class.R:219-227: the whole-pattern index is reused with its dups blanked. */
int bsgpu_pdict3parts_blank_dups(bsgpu_pdict3parts *p3);
void bsgpu_pdict3parts_free(bsgpu_pdict3parts *p3);
int32_t bsgpu_pdict3parts_length(const bsgpu_pdict3parts *p3);
int32_t bsgpu_pdict3parts_tb_width(const bsgpu_pdict3parts *p3);

/* ------------------------------------------------------------------------------------
 * Subjects.  A device-resident subject holds the encoded bytes plus a 2-bit-packed
 * code plane and 1-bit validity planes built on upload.  Replaces hold_XRaw(subject) /
 * _hold_XStringSet(subject) / the views loop of src/match_pdict.c:164,240-262,330,399.
 *   xstring: one sequence (DNAString)
 *   views:   XStringViews / MaskedXString gaps: every view is an independent subject
 *            (src/match_pdict.c:246-262); a view outside the subject is an error
 *            ("'subject' has \"out of limits\" views", :253-254)
 *   set:     DNAStringSet for vcountPDict / vwhichPDict
 *   chunk:   bytes [chunk_start, chunk_start+chunk_len) of a longer sequence of
 *            `subject_len` letters; only Trusted-Band ends n (1-based, subject
 *            coordinates) with report_lo < n <= report_hi are reported, in subject
 *            coordinates, and "out of limits" is judged against the whole subject.  This
 *            is how one long subject is sharded over GPUs; the chunk must carry a halo of
 *            width(pattern)-1 letters around the report range (checked at match time).
 * ------------------------------------------------------------------------------------ */
typedef struct bsgpu_subject bsgpu_subject;

int bsgpu_subject_xstring(const uint8_t *s, int64_t length, bsgpu_subject **out);
/* The same from LETTERS (ASCII): DNAString(<character>) with the encoding done on the device.
 * Replaces the copy-with-lookup behind new_XString_from_CHARACTER (src/XString_class.c:191-218
 * -> _copy_CHARSXP_to_Chars_holder :134-159 -> XVector's Ocopy_bytes_from_i1i2_with_lkup; the
 * tables of src/XString_class.c:18-65 and R/XStringCodec-class.R:114-166).  `lkup`: byte -> code, BSGPU_NA = not a letter, as for
 * bsgpu_subject_fasta.  Error: "key %d (char '%c') not in lookup table" for the first such
 * letter, XVector's text. */
int bsgpu_subject_letters(const uint8_t *letters, int64_t length, const int32_t *lkup, int lkup_length,
		bsgpu_subject **out);
int bsgpu_subject_views(const uint8_t *s, int64_t length,
		const int32_t *views_start, const int32_t *views_width, int32_t nviews,
		bsgpu_subject **out);
int bsgpu_subject_set(const uint8_t *concat, const int32_t *widths, int32_t n,
		bsgpu_subject **out);
int bsgpu_subject_chunk(const uint8_t *chunk, int64_t chunk_len, int64_t chunk_start,
		int64_t subject_len, int64_t report_lo, int64_t report_hi,
		bsgpu_subject **out);
/* Tells the library that the resident letters are to be treated as new (e.g. the caller
 * overwrote them in place through its own device pointer, or wants the derivation inside a
 * timed region as bench.py does): the packed 2-bit / validity planes, which are derived
 * lazily by the first scan that needs them, are dropped.  The exact-dictionary scan reads
 * the letters directly and never builds them.  `stream` is unused (kept for ABI stability). */
int bsgpu_subject_repack(bsgpu_subject *subj, void *stream);
void bsgpu_subject_free(bsgpu_subject *subj);
int64_t bsgpu_subject_nletters(const bsgpu_subject *subj);	/* letters that are scanned */

/* ------------------------------------------------------------------------------------
 * Subject ingestion: FASTA text -> resident subject.  Replaces
 *     read_fasta_files(filexp_list, nrec, skip, seek_first_rec, use_names, elementType, lkup)
 *                                                          src/read_fasta_files.c:412
 * for one in-memory file read whole (nrec = -1, skip = 0, seek_first_rec = FALSE; the host
 * layer cuts the text for other values): parse_FASTA_file()'s line rules (:240-352) and
 * translate() (:31-50) run on the device, one thread per 512-byte chunk, and the letters land
 * directly in the subject layout -- the text is uploaded once and never re-read by the host.
 * `lkup`: the reference's integer lookup (byte -> code, BSGPU_NA = not a letter; keys >=
 * lkup_length are not letters), e.g. get_seqtype_conversion_lookup("B", "DNA").
 * The result is a set subject (one sequence per record) usable with bsgpu_vmatch /
 * bsgpu_oligo_frequency; bsgpu_subject_element() makes a single-sequence subject from one
 * record (device-to-device), which is how a chromosome read once stays resident for repeated
 * matchPDict / countPDict calls; bsgpu_subject_download() brings the encoded letters back
 * for callers that want the XStringSet in host memory like the reference returns it.
 * `info` (fields allocated by the library, release with bsgpu_fasta_info_free): widths, and
 * for names() the offset and length of every description text inside `text`.
 * Errors: "\">\" expected at beginning of line %d" (:322-327).  ninvalid != 0 is the
 * reference's warning "ignored %lld invalid one-letter sequence codes" (:388-391).
 * ------------------------------------------------------------------------------------ */
typedef struct bsgpu_fasta_info {
	int32_t nrec;
	int64_t ninvalid;
	int64_t nletters;
	int32_t *widths;	/* [nrec] */
	int64_t *desc_offset;	/* [nrec] byte offset in `text` of the description (after '>') */
	int32_t *desc_length;	/* [nrec] */
} bsgpu_fasta_info;

int bsgpu_subject_fasta(const uint8_t *text, int64_t nbytes, const int32_t *lkup, int lkup_length,
		bsgpu_subject **out, bsgpu_fasta_info *info);
void bsgpu_fasta_info_free(bsgpu_fasta_info *info);
int32_t bsgpu_subject_nsequences(const bsgpu_subject *subj);
int bsgpu_subject_element(const bsgpu_subject *set, int32_t i, bsgpu_subject **out);
/* encoded letters of all sequences back to back (concat_out, may be NULL) and their widths
 * (widths_out[nsequences], may be NULL) */
int bsgpu_subject_download(const bsgpu_subject *subj, uint8_t *concat_out, int32_t *widths_out);

/* ------------------------------------------------------------------------------------
 * Results (library-allocated host memory; release with bsgpu_result_free()).
 *   COUNTS: counts[M]                                   src/match_reporting.c:163-166
 *   WHICH:  which[n_which], ascending 1-based ids        src/match_reporting.c:150-161
 *   ENDS:   CSR ends_offsets[M+1], ends[]; an empty row is the reference's NULL list
 *           element (src/match_reporting.c:192-201); per pattern the ends are in the
 *           reference's order (ascending; view order for views)
 * For vmatch (set subjects):
 *   COUNTS, collapse=0: counts[M*N], column-major (column j = subject j)
 *           collapse=1/2: counts[] (int weights) or dcounts[] (double weights), length
 *           M (collapse=1) or N (collapse=2)             src/match_pdict.c:85-127,386-432
 *   WHICH:  CSR ends_offsets[N+1], ends[] = ascending 1-based pattern ids per subject
 *                                                        src/match_pdict.c:320-346
 * ------------------------------------------------------------------------------------ */
typedef struct bsgpu_result {
	int32_t M;		/* number of patterns */
	int64_t n_counts;
	int32_t *counts;
	double *dcounts;
	int64_t n_which;
	int32_t *which;
	int64_t n_rows;		/* rows of the CSR (M, or N for vwhich) */
	int64_t *ends_offsets;
	int32_t *ends;
} bsgpu_result;

void bsgpu_result_free(bsgpu_result *res);

/* ------------------------------------------------------------------------------------
 * Matching against a device-resident subject.
 *
 * nparts == 1: one PDict3Parts = a TB_PDict; replaces the body of
 *     match_PDict3Parts_XString()       src/match_pdict.c:153-171
 *     match_PDict3Parts_XStringViews()  src/match_pdict.c:225-264
 * i.e. TB scan (_match_Twobit / _match_tbACtree2, src/match_pdict_Twobit.c:177,
 * src/match_pdict_ACtree2.c:1160), flank verification (_match_pdict_all_flanks,
 * src/match_pdict_utils.c:809) and reporting (_MatchBuf_as_SEXP, src/match_reporting.c:232).
 *
 * nparts > 1: the components of an MTB_PDict; the per-component results are merged as
 * .match.MTB_PDict + ByPos_MIndex_combine do (R/matchPDict.R:335-377,
 * src/MIndex_class.c:230-263): per pattern the sorted distinct ends; counts = their
 * number; which = the union.
 * ------------------------------------------------------------------------------------ */
int bsgpu_match(const bsgpu_pdict3parts *const *parts, int nparts,
		const bsgpu_subject *subject,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS,
		int matches_as, bsgpu_result *res);

/* Replaces vmatch_PDict3Parts_XStringSet(), src/match_pdict.c:484-516 (nparts > 1 only
 * for WHICH, as in R/matchPDict.R:481-482).  weight: int32 or double vector of length
 * N (collapse=1) or M (collapse=2); ignored when collapse=0. */
int bsgpu_vmatch(const bsgpu_pdict3parts *const *parts, int nparts,
		const bsgpu_subject *subject_set,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS,
		int collapse, int weight_is_double, const void *weight, int64_t weight_length,
		int matches_as, bsgpu_result *res);

/* ------------------------------------------------------------------------------------
 * countPDict on a subject in HOST memory, upload pipelined with the scan: what
 *     match_PDict3Parts_XString(pptb, head, tail, subject, max, min, fixed,
 *                               "MATCHES_AS_COUNTS", envir)      src/match_pdict.c:153
 * amounts to when `subject` (or one overlap chunk of it) still has to cross PCIe.
 * The letters are copied in slices (BSGPU_HOST_SLICE_MB, default 32) on a copy stream and slice
 * k is scanned on `stream` while slice k + 1 is in flight; the staging buffer, the events and
 * the device count vector persist between calls.  letters/length: the chunk (pinned host
 * memory makes the copies asynchronous); chunk_start / subject_len / report_lo / report_hi as in
 * bsgpu_subject_chunk (0, length, 0, length for a whole XString).  d_counts != NULL: the counts
 * are ADDED to this device vector (the caller zeroes / reduces it); counts_out != NULL: the
 * counts are copied to this host vector and the call returns when they are there; with
 * counts_out == NULL the call only enqueues work and the caller synchronises `stream`.
 * Dictionaries whose engine needs records on the host side (MTB_PDict without the q-gram
 * index, fixedS = FALSE) are uploaded whole and then counted.
 * ------------------------------------------------------------------------------------ */
int bsgpu_count_host(const bsgpu_pdict3parts *const *parts, int nparts,
		const uint8_t *letters, int64_t length, int64_t chunk_start, int64_t subject_len,
		int64_t report_lo, int64_t report_hi,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS,
		int32_t *counts_out, int32_t *d_counts, void *stream);

/*
 * Device-side counting for callers that own the device (bench, multi-GPU sharding):
 * adds the per-pattern match counts of `subject` into d_counts[M], a DEVICE pointer
 * (e.g. a torch tensor's data_ptr, later all-reduced over NCCL), on CUDA stream
 * `stream` (a cudaStream_t passed as void*; NULL = default stream).  Asynchronous
 * when no candidate buffering is needed (full-width Trusted Band); otherwise it
 * synchronises the stream internally.  n_launches_out (may be NULL) receives the number
 * of kernels this call launched.
 */
int bsgpu_count_device(const bsgpu_pdict3parts *const *parts, int nparts,
		const bsgpu_subject *subject,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS,
		int32_t *d_counts, void *stream, int *n_launches_out);

/*
 * One count vector for the GPUs of a box (one process per GPU).  The reference is a single
 * process: its counts are one INTEGER vector the match loop increments (match_reporting.c:
 * _MatchBuf_report_match, src/match_reporting.c:95-116).  With the subject sharded over N GPUs the
 * same vector lives in ONE GPU's memory and every rank's scan kernels increment it directly, by
 * system-scope reductions through NVLink peer memory: the gather of the per-pattern counts is part
 * of the scan launch, there is no separate collective.
 *   owner rank : bsgpu_shared_counts_create(n, &dptr, handle)   cudaMalloc + zero + IPC handle
 *   other ranks: bsgpu_shared_counts_open(handle, &dptr)        maps the owner's vector (peer access)
 * dptr is what bsgpu_count_device / bsgpu_count_host take as d_counts.  The 64 handle bytes travel
 * by whatever channel the processes share (bench.py: a torch.distributed broadcast).
 * bsgpu_shared_counts_close(dptr, is_owner) unmaps / frees.
 */
#define BSGPU_IPC_HANDLE_BYTES 64
int bsgpu_shared_counts_create(int64_t n, int32_t **d_counts_out, unsigned char handle_out[BSGPU_IPC_HANDLE_BYTES]);
int bsgpu_shared_counts_open(const unsigned char handle[BSGPU_IPC_HANDLE_BYTES], int32_t **d_counts_out);
int bsgpu_shared_counts_close(int32_t *d_counts, int is_owner);

/* ------------------------------------------------------------------------------------
 * Host-buffer entry points with the reference's .Call shapes (upload + match + free).
 * ------------------------------------------------------------------------------------ */
/* match_PDict3Parts_XString, src/match_pdict.c:153 */
int bsgpu_match_PDict3Parts_XString(const bsgpu_pdict3parts *p3,
		const uint8_t *subject, int64_t subject_length,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS,
		int matches_as, bsgpu_result *res);
/* match_PDict3Parts_XStringViews, src/match_pdict.c:225 */
int bsgpu_match_PDict3Parts_XStringViews(const bsgpu_pdict3parts *p3,
		const uint8_t *subject, int64_t subject_length,
		const int32_t *views_start, const int32_t *views_width, int32_t nviews,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS,
		int matches_as, bsgpu_result *res);
/* vmatch_PDict3Parts_XStringSet, src/match_pdict.c:484 */
int bsgpu_vmatch_PDict3Parts_XStringSet(const bsgpu_pdict3parts *p3,
		const uint8_t *subject_concat, const int32_t *subject_widths, int32_t nsubject,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS,
		int collapse, int weight_is_double, const void *weight, int64_t weight_length,
		int matches_as, bsgpu_result *res);

/* ByPos_MIndex_combine, src/MIndex_class.c:230: per-row sorted union of CSR results. */
int bsgpu_mindex_combine(const bsgpu_result *const *components, int ncomponents,
		bsgpu_result *res);

/* ------------------------------------------------------------------------------------
 * Matching with indels on a non-preprocessed dictionary (`plain` from
 * bsgpu_patterns_build): the counts behind countPDict / whichPDict / vcountPDict / vwhichPDict
 * with with.indels=TRUE, i.e. the number of "best local matches" _match_pattern_indels reports
 * per pattern (src/match_pattern_indels.c:22-107), every view / set element being an
 * independent subject.  counts_out: M entries, or M x N column-major when per_sequence != 0.
 * The per-position core (csrc/bsgpu_indels.cuh) equals the reference on the CPU
 * (tests/test_indels_oracle.py), the engine equals it on the device (tests/test_gpu_indels.py).
 * Patterns not longer than max_mismatch are counted without indels, as the reference does
 * (src/match_pattern.c:95-96).  Limits: 1 <= max_mismatch <= 15, no subject chunks,
 * min.mismatch = 0 (R/match-utils.R:54-57).
 * ------------------------------------------------------------------------------------ */
int bsgpu_count_indels(const bsgpu_pdict3parts *plain, const bsgpu_subject *subject, int max_mismatch,
		int fixedP, int fixedS, int per_sequence, int32_t *counts_out);

/* ------------------------------------------------------------------------------------
 * oligonucleotideFrequency.  Replaces
 *     XString_oligo_frequency(x, width, step, as_prob, as_array, fast_moving_side,
 *                             with_labels, base_codes)     src/letter_frequency.c:634
 *     XStringSet_oligo_frequency(..., simplify_as, ...)     src/letter_frequency.c:667
 * i.e. the rolling two-bit signature walk of update_int_oligo_freqs /
 * update_double_oligo_freqs (src/letter_frequency.c:183-267): the window of `width` letters
 * starting at 0-based position s of a sequence is counted iff s % step == 0 and all its
 * letters are A/C/G/T.  Entry `offset` of a row is the window whose letters spell offset in
 * base 4 (A=0 C=1 G=2 T=3), first letter most significant (fast.moving.side = "right",
 * invert_twobit_order = 0) or least significant ("left" / as.array, 1).
 *   collapsed = 1: one row of 4^width entries over all sequences of the subject (a
 *                  DNAString, or simplify.as = "collapsed")
 *   collapsed = 0: nseq x 4^width, column-major like allocMatrix() (simplify.as = "matrix";
 *                  "list" is its rows).  nseq * 4^width must fit an R matrix (< 2^31).
 *   as_prob = 0: counts_out (int32); as_prob = 1: freq_out (double), every row divided by
 *                  its sum as normalize_oligo_freqs() does (:286-301; all-zero rows stay 0).
 * Chunk subjects count the windows whose last letter lies in the chunk's report range, with
 * `step` judged in whole-subject coordinates, so the tables of the chunks of a sharded
 * sequence add up to the table of the sequence.  Labels, dim and dimnames
 * (format_oligo_freqs, :414-452) are the host layer's job.
 * Errors: width outside 1..15 ("_new_TwobitEncodingBuffer(): 'buflength' must be >= 1 and
 * <= 15", src/utils.c:101-102), step < 1 (R/letterFrequency.R:26-34).
 * ------------------------------------------------------------------------------------ */
int bsgpu_oligo_frequency(const bsgpu_subject *subject, int width, int step,
		int invert_twobit_order, int collapsed, int as_prob,
		int32_t *counts_out, double *freq_out);
/* Device-side twin for callers that own the device (bench, multi-GPU sharding): adds the
 * collapsed table of `subject` into d_counts[4^width] (DEVICE pointer, 64-bit counters) on
 * `stream`; asynchronous. */
int bsgpu_oligo_count_device(const bsgpu_subject *subject, int width, int step,
		int invert_twobit_order, unsigned long long *d_counts, void *stream);
/* Host-buffer entry points with the reference's .Call shapes (upload + count + free). */
int bsgpu_XString_oligo_frequency(const uint8_t *x, int64_t x_length, int width, int step,
		int invert_twobit_order, int as_prob, int32_t *counts_out, double *freq_out);
int bsgpu_XStringSet_oligo_frequency(const uint8_t *concat, const int32_t *widths, int32_t n,
		int width, int step, int invert_twobit_order, int collapsed, int as_prob,
		int32_t *counts_out, double *freq_out);

/* Human-readable description of the engine the last match/count call used, e.g.
 * "byteseed_hash S=10 seed=16 filter_bytes=19779008 buckets=6593006 ..." or
 * "qgram_hamming q=12 shifts=7 entries=6930000" or "tb_hash w=12 buckets=2097152 + verify". */
const char *bsgpu_last_plan(void);

/* Kernel launch counter (every kernel this library launched since process start). */
int64_t bsgpu_launch_count(void);

/* Opt-in timing of the hot kernels (pack_subject_kernel, scan_byteseed_kernel,
 * match_plain_kernel, scan_tb_kernel, qscan_kernel, oligo_freq_kernel,
 * fasta_summary_kernel, fasta_emit_kernel, encode_letters_kernel) with CUDA events recorded on the
 * launching stream.  bsgpu_kernel_timing(1) starts collecting (at most 8192 launches),
 * bsgpu_kernel_times() waits for them, writes up to max_entries (name, milliseconds) pairs in
 * launch order, forgets them and returns how many it wrote (-1 on error). */
int bsgpu_kernel_timing(int enable);
int bsgpu_kernel_times(const char **names, float *ms, int max_entries);

/* Micro-benchmarks used to establish the memory rooflines on this device (GB/s):
 * a streaming copy and a random 32-byte-sector gather from an L2-resident buffer. */
int bsgpu_probe_peaks(double *copy_gbs, double *l2_gather_gsectors, int64_t gather_table_bytes);

#ifdef __cplusplus
}
#endif
#endif
