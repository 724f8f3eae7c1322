/*
 * ref_harness.c -- plain-C driver around the *reference's own* .Call entry
 * points (TEST INFRASTRUCTURE ONLY; never linked into the product).
 *
 * The reference's hot-path C files are compiled unmodified from
 * /root/reference/src against the mini-R stand-in in oracle/shim/ (see
 * oracle/Makefile); this file builds the R-level objects those entry points
 * expect (PreprocessedTB, DNAStringSet, DNAString, Dups) and unpacks the SEXP
 * results into flat arrays that Python reads through ctypes (oracle/ref.py).
 *
 * Entry points driven (reference file:line):
 *   build_Twobit                      src/match_pdict_Twobit.c:104
 *   ACtree2_build                     src/match_pdict_ACtree2.c:792
 *   match_PDict3Parts_XString         src/match_pdict.c:153
 *   match_PDict3Parts_XStringViews    src/match_pdict.c:225
 *   vmatch_PDict3Parts_XStringSet     src/match_pdict.c:484
 *   ByPos_MIndex_endIndex / _combine  src/MIndex_class.c:134,230
 *   ACtree2_nnodes / _has_all_flinks / _compute_all_flinks
 *                                     src/match_pdict_ACtree2.c:681,972,981
 *   XString_oligo_frequency / XStringSet_oligo_frequency
 *                                     src/letter_frequency.c:634,667
 *   read_fasta_files                  src/read_fasta_files.c:412
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "shim/minir.h"
#include "shim/S4Vectors_interface.h"
#include "shim/XVector_interface.h"

/* reference entry points (prototypes restated; they live in the reference's .c files) */
void _init_bytewise_match_tables(void);
SEXP build_Twobit(SEXP tb, SEXP pp_exclude, SEXP base_codes);
SEXP IntegerBAB_new(SEXP max_nblock);
SEXP ACtree2_nodebuf_max_nblock(void);
SEXP ACtree2_nodeextbuf_max_nblock(void);
SEXP ACtree2_build(SEXP tb, SEXP pp_exclude, SEXP base_codes,
		   SEXP nodebuf_ptr, SEXP nodeextbuf_ptr);
SEXP ACtree2_nnodes(SEXP pptb);
SEXP ACtree2_has_all_flinks(SEXP pptb);
SEXP ACtree2_compute_all_flinks(SEXP pptb);
SEXP match_PDict3Parts_XString(SEXP pptb, SEXP pdict_head, SEXP pdict_tail,
		SEXP subject, SEXP max_mismatch, SEXP min_mismatch, SEXP fixed,
		SEXP matches_as, SEXP envir);
SEXP match_PDict3Parts_XStringViews(SEXP pptb, SEXP pdict_head, SEXP pdict_tail,
		SEXP subject, SEXP views_start, SEXP views_width,
		SEXP max_mismatch, SEXP min_mismatch, SEXP fixed,
		SEXP matches_as, SEXP envir);
SEXP vmatch_PDict3Parts_XStringSet(SEXP pptb, SEXP pdict_head, SEXP pdict_tail,
		SEXP subject, SEXP max_mismatch, SEXP min_mismatch, SEXP fixed,
		SEXP collapse, SEXP weight, SEXP matches_as, SEXP envir);
SEXP match_XStringSet_XString(SEXP pattern, SEXP subject, SEXP max_mismatch, SEXP min_mismatch,
		SEXP with_indels, SEXP fixed, SEXP algorithm, SEXP matches_as, SEXP envir);
SEXP match_XStringSet_XStringViews(SEXP pattern, SEXP subject, SEXP views_start, SEXP views_width,
		SEXP max_mismatch, SEXP min_mismatch, SEXP with_indels, SEXP fixed,
		SEXP algorithm, SEXP matches_as, SEXP envir);
SEXP vmatch_XStringSet_XStringSet(SEXP pattern, SEXP subject, SEXP max_mismatch, SEXP min_mismatch,
		SEXP with_indels, SEXP fixed, SEXP algorithm, SEXP collapse, SEXP weight,
		SEXP matches_as, SEXP envir);
SEXP XString_oligo_frequency(SEXP x, SEXP width, SEXP step, SEXP as_prob, SEXP as_array,
		SEXP fast_moving_side, SEXP with_labels, SEXP base_codes);
SEXP XStringSet_oligo_frequency(SEXP x, SEXP width, SEXP step, SEXP as_prob, SEXP as_array,
		SEXP fast_moving_side, SEXP with_labels, SEXP simplify_as, SEXP base_codes);
SEXP read_fasta_files(SEXP filexp_list, SEXP nrec, SEXP skip, SEXP seek_first_rec,
		SEXP use_names, SEXP elementType, SEXP lkup);
SEXP ByPos_MIndex_endIndex(SEXP x_high2low, SEXP x_ends, SEXP x_width0);
SEXP ByPos_MIndex_combine(SEXP ends_listlist);

/* Out-of-scope Biostrings internals referenced by the compiled files (the
 * non-preprocessed dictionary path and SparseMIndex); never reached here. */
void _copy_CHARSXP_to_Chars_holder(void) { Rf_error("ref_harness: out of scope"); }
void _new_CHARSXP_from_Chars_holder(void) { Rf_error("ref_harness: out of scope"); }
void _get_val_from_env(void) { Rf_error("ref_harness: out of scope"); }

static int initialized = 0;
static SEXP result = NULL;		/* result of the last call (scratch) */
static SEXP stash[64];			/* results kept for ByPos_MIndex_combine */
static int nstash = 0;

#define GUARD_BEGIN \
	minir_errmsg[0] = 0; \
	if (!initialized) { _init_bytewise_match_tables(); initialized = 1; } \
	minir_jmp_armed = 1; \
	if (setjmp(minir_jmpbuf) == 0) {
#define GUARD_END(okval, errval) \
		minir_jmp_armed = 0; \
		minir_set_generation(2); \
		return okval; \
	} \
	minir_jmp_armed = 0; \
	minir_set_generation(2); \
	return errval;

const char *ref_last_error(void) { return minir_errmsg; }
const char *ref_last_warning(void) { return minir_warnmsg; }
void ref_clear_warning(void) { minir_warnmsg[0] = 0; }

/* Frees everything allocated by match calls since the last release
 * (what returning from .Call + gc does in R).  PPTB handles survive. */
void ref_release(void)
{
	result = NULL;
	nstash = 0;
	minir_release_scratch();
}

static SEXP new_set(const char *classname, int n, const unsigned char *concat,
		    const int *widths)
{
	const char **ptrs = malloc((size_t) (n ? n : 1) * sizeof(char *));
	size_t off = 0;
	int i;
	SEXP x;

	for (i = 0; i < n; i++) {
		ptrs[i] = (const char *) concat + off;
		off += (size_t) widths[i];
	}
	x = minir_new_XStringSet(classname, n, ptrs, widths);
	free(ptrs);
	return x;
}

static SEXP new_int(int n, const int *vals)
{
	SEXP x = NEW_INTEGER(n);

	if (n)
		memcpy(INTEGER(x), vals, (size_t) n * sizeof(int));
	return x;
}

static SEXP new_base_codes(void)
{
	static const int codes[4] = {1, 2, 4, 8};	/* A C G T, R/XStringCodec-class.R:114 */

	return new_int(4, codes);
}

/* IRanges Dups(high2low): slots high2low + low2high (list; NULL when no dup). */
static SEXP new_Dups(int n, const int *high2low)
{
	SEXP dups = NEW_OBJECT(MAKE_CLASS("Dups")), l2h = NEW_LIST(n), elt;
	int *cnt = calloc((size_t) (n ? n : 1), sizeof(int)), i, k;

	for (i = 0; i < n; i++)
		if (high2low[i] != NA_INTEGER)
			cnt[high2low[i] - 1]++;
	for (k = 0; k < n; k++)
		if (cnt[k] != 0) {
			elt = NEW_INTEGER(cnt[k]);
			elt->length = 0;	/* used as fill cursor below */
			SET_VECTOR_ELT(l2h, k, elt);
		}
	for (i = 0; i < n; i++)
		if (high2low[i] != NA_INTEGER) {
			elt = VECTOR_ELT(l2h, high2low[i] - 1);
			INTEGER(elt)[elt->length++] = i + 1;
		}
	free(cnt);
	SET_SLOT(dups, Rf_install("high2low"), new_int(n, high2low));
	SET_SLOT(dups, Rf_install("low2high"), l2h);
	return dups;
}

/*
 * Builds a PreprocessedTB object of class 'algo' ("Twobit" | "ACtree2") the
 * way R/PDict-class.R:95-105 (Twobit) and :146-167 (ACtree2) do.
 * 'tb_concat'/'tb_widths' must stay alive as long as the handle is used.
 * Returns NULL on error (see ref_last_error()).
 */
void *ref_pptb_build(const char *algo, int M, const unsigned char *tb_concat,
		     const int *tb_widths, const int *pp_exclude, int *high2low_out)
{
	SEXP tb, excl, codes, C_ans, pptb, h2l, nodebuf, nodeextbuf;

	GUARD_BEGIN
	minir_set_generation(2);
	tb = new_set("DNAStringSet", M, tb_concat, tb_widths);
	excl = pp_exclude != NULL ? new_int(M, pp_exclude) : R_NilValue;
	codes = new_base_codes();
	pptb = NEW_OBJECT(MAKE_CLASS(algo));
	if (strcmp(algo, "Twobit") == 0) {
		C_ans = build_Twobit(tb, excl, codes);
		SET_SLOT(pptb, Rf_install("sign2pos"), VECTOR_ELT(C_ans, 0));
	} else if (strcmp(algo, "ACtree2") == 0) {
		nodebuf = IntegerBAB_new(ACtree2_nodebuf_max_nblock());
		nodeextbuf = IntegerBAB_new(ACtree2_nodeextbuf_max_nblock());
		C_ans = ACtree2_build(tb, excl, codes, nodebuf, nodeextbuf);
		SET_SLOT(pptb, Rf_install("nodebuf_ptr"), nodebuf);
		SET_SLOT(pptb, Rf_install("nodeextbuf_ptr"), nodeextbuf);
	} else {
		Rf_error("ref_harness: unknown algo \"%s\"", algo);
	}
	h2l = VECTOR_ELT(C_ans, 1);
	if (high2low_out != NULL)
		memcpy(high2low_out, INTEGER(h2l), (size_t) M * sizeof(int));
	SET_SLOT(pptb, Rf_install("tb"), tb);
	SET_SLOT(pptb, Rf_install("exclude_dups0"), Rf_ScalarLogical(pp_exclude != NULL));
	SET_SLOT(pptb, Rf_install("dups"), new_Dups(M, INTEGER(h2l)));
	SET_SLOT(pptb, Rf_install("base_codes"), codes);
	minir_persist(pptb);
	minir_release_scratch();	/* build-time scratch + AE buffers */
	GUARD_END(pptb, (minir_release_scratch(), (void *) NULL))
}

/* R/PDict-class.R:219-227: reuse pptb0 with dups blanked to all-NA. */
int ref_pptb_blank_dups(void *pptb_)
{
	SEXP pptb = pptb_;
	int M, i, *na;

	GUARD_BEGIN
	minir_set_generation(2);
	M = get_XVectorList_length(GET_SLOT(pptb, Rf_install("tb")));
	na = malloc((size_t) M * sizeof(int));
	for (i = 0; i < M; i++)
		na[i] = NA_INTEGER;
	SET_SLOT(pptb, Rf_install("dups"), new_Dups(M, na));
	SET_SLOT(pptb, Rf_install("exclude_dups0"), Rf_ScalarLogical(1));
	free(na);
	GUARD_END(0, -1)
}

/* Hooks of the product's dispatch library (absent from the reference library: weak). */
extern SEXP bsgpu_dispatch_detach(SEXP pptb) __attribute__((weak));
extern int bsgpu_dispatch_live_indexes(void) __attribute__((weak));

/* Forget the device index an object carries, as readRDS() would hand the object back. */
int ref_pptb_detach_device(void *pptb)
{
	if (bsgpu_dispatch_detach == NULL)
		return -1;
	GUARD_BEGIN
	minir_set_generation(2);
	bsgpu_dispatch_detach(pptb);
	GUARD_END(0, -1)
}

int ref_dispatch_live_indexes(void)
{
	return bsgpu_dispatch_live_indexes == NULL ? -1 : bsgpu_dispatch_live_indexes();
}

void ref_pptb_free(void *pptb)
{
	minir_free_object(pptb);
	minir_release_scratch();
	result = NULL;
	nstash = 0;
}

int ref_actree2_nnodes(void *pptb)
{
	int n;

	GUARD_BEGIN
	minir_set_generation(2);
	n = INTEGER(ACtree2_nnodes(pptb))[0];
	GUARD_END(n, -1)
}

int ref_actree2_has_all_flinks(void *pptb)
{
	int n;

	GUARD_BEGIN
	minir_set_generation(2);
	n = LOGICAL(ACtree2_has_all_flinks(pptb))[0];
	GUARD_END(n, -1)
}

int ref_actree2_compute_all_flinks(void *pptb)
{
	GUARD_BEGIN
	minir_set_generation(2);
	ACtree2_compute_all_flinks(pptb);
	GUARD_END(0, -1)
}

static const char *mode_name(int mode)
{
	switch (mode) {
	case 0: return "MATCHES_AS_NULL";
	case 1: return "MATCHES_AS_WHICH";
	case 2: return "MATCHES_AS_COUNTS";
	case 4: return "MATCHES_AS_ENDS";
	}
	Rf_error("ref_harness: unsupported mode %d", mode);
}

static SEXP new_fixed(int fixedP, int fixedS)
{
	SEXP f = NEW_LOGICAL(2);

	LOGICAL(f)[0] = fixedP;
	LOGICAL(f)[1] = fixedS;
	return f;
}

static SEXP opt_set(int M, const unsigned char *concat, const int *widths)
{
	/* head(threeparts)/tail(threeparts) are NULL when all widths are 0
	   (R/PDict-class.R:184-203); the caller passes widths == NULL then. */
	return widths != NULL ? new_set("DNAStringSet", M, concat, widths) : R_NilValue;
}

/* match_PDict3Parts_XString, src/match_pdict.c:153 */
int ref_match_xstring(void *pptb, int M,
		const unsigned char *head_concat, const int *head_widths,
		const unsigned char *tail_concat, const int *tail_widths,
		const unsigned char *subject, int subject_length,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS, int mode)
{
	GUARD_BEGIN
	minir_set_generation(2);
	result = match_PDict3Parts_XString(pptb,
			opt_set(M, head_concat, head_widths),
			opt_set(M, tail_concat, tail_widths),
			minir_new_XString("DNAString", (const char *) subject, subject_length),
			Rf_ScalarInteger(max_mismatch), Rf_ScalarInteger(min_mismatch),
			new_fixed(fixedP, fixedS),
			Rf_mkString(mode_name(mode)), R_NilValue);
	GUARD_END(0, -1)
}

/* match_PDict3Parts_XStringViews, src/match_pdict.c:225 */
int ref_match_views(void *pptb, int M,
		const unsigned char *head_concat, const int *head_widths,
		const unsigned char *tail_concat, const int *tail_widths,
		const unsigned char *subject, int subject_length,
		const int *views_start, const int *views_width, int nviews,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS, int mode)
{
	GUARD_BEGIN
	minir_set_generation(2);
	result = match_PDict3Parts_XStringViews(pptb,
			opt_set(M, head_concat, head_widths),
			opt_set(M, tail_concat, tail_widths),
			minir_new_XString("DNAString", (const char *) subject, subject_length),
			new_int(nviews, views_start), new_int(nviews, views_width),
			Rf_ScalarInteger(max_mismatch), Rf_ScalarInteger(min_mismatch),
			new_fixed(fixedP, fixedS),
			Rf_mkString(mode_name(mode)), R_NilValue);
	GUARD_END(0, -1)
}

/* vmatch_PDict3Parts_XStringSet, src/match_pdict.c:484 */
int ref_vmatch_set(void *pptb, int M,
		const unsigned char *head_concat, const int *head_widths,
		const unsigned char *tail_concat, const int *tail_widths,
		const unsigned char *subject_concat, const int *subject_widths, int nsubject,
		int max_mismatch, int min_mismatch, int fixedP, int fixedS,
		int collapse, int weight_is_double, const void *weight, int weight_length,
		int mode)
{
	SEXP w;

	GUARD_BEGIN
	minir_set_generation(2);
	if (weight_is_double) {
		w = NEW_NUMERIC(weight_length);
		memcpy(REAL(w), weight, (size_t) weight_length * sizeof(double));
	} else {
		w = new_int(weight_length, weight);
	}
	result = vmatch_PDict3Parts_XStringSet(pptb,
			opt_set(M, head_concat, head_widths),
			opt_set(M, tail_concat, tail_widths),
			new_set("DNAStringSet", nsubject, subject_concat, subject_widths),
			Rf_ScalarInteger(max_mismatch), Rf_ScalarInteger(min_mismatch),
			new_fixed(fixedP, fixedS),
			Rf_ScalarInteger(collapse), w,
			Rf_mkString(mode_name(mode)), R_NilValue);
	GUARD_END(0, -1)
}

/* ---- the non-preprocessed dictionary path (pattern = plain XStringSet) ----- */

/* match_XStringSet_XString, src/match_pdict.c:174 */
int ref_match_set_xstring(const unsigned char *pat_concat, const int *pat_widths, int M,
		const unsigned char *subject, int subject_length,
		int max_mismatch, int min_mismatch, int with_indels, int fixedP, int fixedS,
		const char *algo, int mode)
{
	GUARD_BEGIN
	minir_set_generation(2);
	result = match_XStringSet_XString(
			new_set("DNAStringSet", M, pat_concat, pat_widths),
			minir_new_XString("DNAString", (const char *) subject, subject_length),
			Rf_ScalarInteger(max_mismatch), Rf_ScalarInteger(min_mismatch),
			Rf_ScalarLogical(with_indels), new_fixed(fixedP, fixedS),
			Rf_mkString(algo), Rf_mkString(mode_name(mode)), R_NilValue);
	GUARD_END(0, -1)
}

/* match_XStringSet_XStringViews, src/match_pdict.c:267 */
int ref_match_set_views(const unsigned char *pat_concat, const int *pat_widths, int M,
		const unsigned char *subject, int subject_length,
		const int *views_start, const int *views_width, int nviews,
		int max_mismatch, int min_mismatch, int with_indels, int fixedP, int fixedS,
		const char *algo, int mode)
{
	GUARD_BEGIN
	minir_set_generation(2);
	result = match_XStringSet_XStringViews(
			new_set("DNAStringSet", M, pat_concat, pat_widths),
			minir_new_XString("DNAString", (const char *) subject, subject_length),
			new_int(nviews, views_start), new_int(nviews, views_width),
			Rf_ScalarInteger(max_mismatch), Rf_ScalarInteger(min_mismatch),
			Rf_ScalarLogical(with_indels), new_fixed(fixedP, fixedS),
			Rf_mkString(algo), Rf_mkString(mode_name(mode)), R_NilValue);
	GUARD_END(0, -1)
}

/* vmatch_XStringSet_XStringSet, src/match_pdict.c:519 */
int ref_vmatch_set_set(const unsigned char *pat_concat, const int *pat_widths, int M,
		const unsigned char *subject_concat, const int *subject_widths, int nsubject,
		int max_mismatch, int min_mismatch, int with_indels, int fixedP, int fixedS,
		const char *algo, int collapse, int weight_is_double, const void *weight,
		int weight_length, int mode)
{
	SEXP w;

	GUARD_BEGIN
	minir_set_generation(2);
	if (weight_is_double) {
		w = NEW_NUMERIC(weight_length);
		memcpy(REAL(w), weight, (size_t) weight_length * sizeof(double));
	} else {
		w = new_int(weight_length, weight);
	}
	result = vmatch_XStringSet_XStringSet(
			new_set("DNAStringSet", M, pat_concat, pat_widths),
			new_set("DNAStringSet", nsubject, subject_concat, subject_widths),
			Rf_ScalarInteger(max_mismatch), Rf_ScalarInteger(min_mismatch),
			Rf_ScalarLogical(with_indels), new_fixed(fixedP, fixedS),
			Rf_mkString(algo), Rf_ScalarInteger(collapse), w,
			Rf_mkString(mode_name(mode)), R_NilValue);
	GUARD_END(0, -1)
}

/* ---- MIndex helpers (R/MIndex-class.R:178-219) ---------------------------- */

/* Keep the current (ENDS) result for a later ref_mindex_combine(). */
int ref_result_stash(void)
{
	if (result == NULL || result->type != VECSXP || nstash >= 64)
		return -1;
	stash[nstash++] = result;
	return nstash;
}

/* ByPos_MIndex_combine(list(stash...)), src/MIndex_class.c:230 */
int ref_mindex_combine(void)
{
	SEXP ll;
	int j;

	GUARD_BEGIN
	minir_set_generation(2);
	ll = NEW_LIST(nstash);
	for (j = 0; j < nstash; j++)
		SET_VECTOR_ELT(ll, j, stash[j]);
	result = ByPos_MIndex_combine(ll);
	nstash = 0;
	GUARD_END(0, -1)
}

/* ByPos_MIndex_endIndex(high2low, <current result>, width0), src/MIndex_class.c:134 */
int ref_mindex_endindex(int M, const int *high2low, const int *width0)
{
	GUARD_BEGIN
	minir_set_generation(2);
	result = ByPos_MIndex_endIndex(
			high2low != NULL ? new_int(M, high2low) : R_NilValue,
			result,
			width0 != NULL ? new_int(M, width0) : R_NilValue);
	GUARD_END(0, -1)
}

/* ---- oligonucleotideFrequency (src/letter_frequency.c:634,667) ---------------- */

static SEXP new_named_base_codes(void)
{
	static const char *const letters[4] = {"A", "C", "G", "T"};
	SEXP codes = new_base_codes(), names = NEW_CHARACTER(4);
	int i;

	for (i = 0; i < 4; i++)
		SET_STRING_ELT(names, i, Rf_mkChar(letters[i]));
	SET_NAMES(codes, names);
	return codes;
}

int ref_oligo_frequency_xstring(const unsigned char *x, int x_length, int width, int step,
		int as_prob, int as_array, int fast_moving_side_is_left, int with_labels)
{
	GUARD_BEGIN
	minir_set_generation(2);
	result = XString_oligo_frequency(
			minir_new_XString("DNAString", (const char *) x, x_length),
			Rf_ScalarInteger(width), Rf_ScalarInteger(step),
			Rf_ScalarLogical(as_prob), Rf_ScalarLogical(as_array),
			Rf_mkString(fast_moving_side_is_left ? "left" : "right"),
			Rf_ScalarLogical(with_labels), new_named_base_codes());
	GUARD_END(0, -1)
}

/* simplify_as: "matrix" | "list" | "collapsed" (R/letterFrequency.R:66-75) */
int ref_oligo_frequency_set(const unsigned char *concat, const int *widths, int n, int width,
		int step, int as_prob, int as_array, int fast_moving_side_is_left, int with_labels,
		const char *simplify_as)
{
	GUARD_BEGIN
	minir_set_generation(2);
	result = XStringSet_oligo_frequency(
			new_set("DNAStringSet", n, concat, widths),
			Rf_ScalarInteger(width), Rf_ScalarInteger(step),
			Rf_ScalarLogical(as_prob), Rf_ScalarLogical(as_array),
			Rf_mkString(fast_moving_side_is_left ? "left" : "right"),
			Rf_ScalarLogical(with_labels), Rf_mkString(simplify_as),
			new_named_base_codes());
	GUARD_END(0, -1)
}

/* ---- read_fasta_files (src/read_fasta_files.c:412) ----------------------------- */

/* One in-memory file; `lkup` is the integer lookup R builds with
 * get_seqtype_conversion_lookup("B", "DNA") (NA for bytes that are not DNA letters). */
int ref_read_fasta(const char *text, long nbytes, const int *lkup, int lkup_length,
		int nrec, int skip, int seek_first_rec, int use_names)
{
	SEXP files, names;

	GUARD_BEGIN
	minir_set_generation(2);
	files = NEW_LIST(1);
	SET_VECTOR_ELT(files, 0, minir_new_filexp(text, nbytes));
	names = NEW_CHARACTER(1);
	SET_STRING_ELT(names, 0, Rf_mkChar("<memory>"));
	SET_NAMES(files, names);
	result = read_fasta_files(files, Rf_ScalarInteger(nrec), Rf_ScalarInteger(skip),
			Rf_ScalarLogical(seek_first_rec), Rf_ScalarLogical(use_names),
			Rf_mkString("DNAString"),
			lkup != NULL ? new_int(lkup_length, lkup) : R_NilValue);
	GUARD_END(0, -1)
}

/* several in-memory files in one call: texts back to back, nbytes[i] each (one record counter over
   the whole list, src/read_fasta_files.c:447-458) */
int ref_read_fasta_multi(const char *texts, const long *nbytes, int nfiles, const int *lkup, int lkup_length,
		int nrec, int skip, int seek_first_rec, int use_names)
{
	SEXP files, names;
	long off = 0;
	int i;

	GUARD_BEGIN
	minir_set_generation(2);
	files = NEW_LIST(nfiles);
	names = NEW_CHARACTER(nfiles);
	for (i = 0; i < nfiles; i++) {
		SET_VECTOR_ELT(files, i, minir_new_filexp(texts + off, nbytes[i]));
		SET_STRING_ELT(names, i, Rf_mkChar("<memory>"));
		off += nbytes[i];
	}
	SET_NAMES(files, names);
	result = read_fasta_files(files, Rf_ScalarInteger(nrec), Rf_ScalarInteger(skip),
			Rf_ScalarLogical(seek_first_rec), Rf_ScalarLogical(use_names),
			Rf_mkString("DNAString"),
			lkup != NULL ? new_int(lkup_length, lkup) : R_NilValue);
	GUARD_END(0, -1)
}

/* XStringSet results: number of elements, widths, the letters back to back, names */
int ref_result_set_length(void) { return get_XVectorList_length(result); }

long ref_result_set_widths(int *out)
{
	XVectorList_holder h = hold_XVectorList(result);
	long total = 0;
	int i;

	for (i = 0; i < h.length; i++)
		total += (out[i] = get_elt_from_XRawList_holder(&h, i).length);
	return total;
}

void ref_result_set_concat(unsigned char *out)
{
	XVectorList_holder h = hold_XVectorList(result);
	int i;

	for (i = 0; i < h.length; i++) {
		Chars_holder e = get_elt_from_XRawList_holder(&h, i);

		memcpy(out, e.ptr, (size_t) e.length);
		out += e.length;
	}
}

long ref_result_set_names(char *buf, long buflen)
{
	SEXP names = get_XVectorList_names(result);
	long i, used = 0;

	if (names == R_NilValue)
		return 0;
	for (i = 0; i < names->length; i++) {
		const char *s = CHAR(STRING_ELT(names, i));
		long n = (long) strlen(s);

		if (used + n + 2 > buflen)
			return -1;
		memcpy(buf + used, s, (size_t) n);
		used += n;
		buf[used++] = '\n';
	}
	buf[used] = 0;
	return names->length;
}

/* ---- result access ---------------------------------------------------------- */

/* dim attribute of the result (or of its first list element): returns ndim, fills out[16] */
int ref_result_dim(int *out)
{
	SEXP x = result, dim;
	int i;

	if (x == NULL)
		return -1;
	if (x->type == VECSXP && x->length > 0)
		x = VECTOR_ELT(x, 0);
	dim = GET_DIM(x);
	if (dim == R_NilValue)
		return 0;
	for (i = 0; i < LENGTH(dim) && i < 16; i++)
		out[i] = INTEGER(dim)[i];
	return LENGTH(dim);
}

/* Labels of the result (or of its first list element) joined with '\n': names() of a vector,
 * else dimnames()[[which]].  Returns the number of labels, 0 when there are none, -1 when
 * buf is too small. */
long ref_result_labels(int which, char *buf, long buflen)
{
	SEXP x = result, labels;
	long i, used = 0;

	if (x == NULL)
		return -1;
	if (x->type == VECSXP && x->length > 0)
		x = VECTOR_ELT(x, 0);
	labels = GET_NAMES(x);
	if (labels == R_NilValue || labels == NULL) {
		SEXP dn = GET_DIMNAMES(x);

		if (dn == R_NilValue || which >= LENGTH(dn))
			return 0;
		labels = VECTOR_ELT(dn, which);
	}
	if (labels == R_NilValue)
		return 0;
	for (i = 0; i < labels->length; i++) {
		const char *s = CHAR(STRING_ELT(labels, i));
		long k = (long) strlen(s);

		if (used + k + 2 > buflen)
			return -1;
		memcpy(buf + used, s, (size_t) k);
		used += k;
		buf[used++] = '\n';
	}
	buf[used] = 0;
	return labels->length;
}

/* type of the first non-NULL element of a list result */
int ref_result_list_elt_type(void)
{
	long i;

	for (i = 0; i < result->length; i++)
		if (VECTOR_ELT(result, i) != R_NilValue)
			return VECTOR_ELT(result, i)->type;
	return NILSXP;
}

void ref_result_list_flat_double(double *out)
{
	long i;

	for (i = 0; i < result->length; i++) {
		SEXP e = VECTOR_ELT(result, i);

		if (e == R_NilValue)
			continue;
		memcpy(out, REAL(e), (size_t) LENGTH(e) * sizeof(double));
		out += LENGTH(e);
	}
}

int ref_result_type(void) { return result == NULL ? -1 : result->type; }
long ref_result_length(void) { return result == NULL ? -1 : result->length; }

void ref_result_copy_int(int *out)
{
	memcpy(out, INTEGER(result), (size_t) result->length * sizeof(int));
}

void ref_result_copy_double(double *out)
{
	memcpy(out, REAL(result), (size_t) result->length * sizeof(double));
}

/* list results: element lengths (-1 = NULL) and total */
long ref_result_list_lengths(int *out)
{
	long i, total = 0;

	for (i = 0; i < result->length; i++) {
		SEXP e = VECTOR_ELT(result, i);

		if (e == R_NilValue) {
			out[i] = -1;
		} else {
			out[i] = LENGTH(e);
			total += LENGTH(e);
		}
	}
	return total;
}

void ref_result_list_flat(int *out)
{
	long i;

	for (i = 0; i < result->length; i++) {
		SEXP e = VECTOR_ELT(result, i);

		if (e == R_NilValue)
			continue;
		memcpy(out, INTEGER(e), (size_t) LENGTH(e) * sizeof(int));
		out += LENGTH(e);
	}
}
