/*
 * match_pdict_gpu.c -- the R-facing C dispatch of the PDict path over libbsgpu.
 *
 * Drop-in for the .Call entry points of Biostrings' src/match_pdict.c,
 * src/match_pdict_Twobit.c, src/match_pdict_ACtree2.c and src/MIndex_class.c: SAME names,
 * SAME arguments, SAME result shapes, so R/PDict-class.R, R/matchPDict.R and
 * src/R_init_Biostrings.c stay untouched.  "Host code stays C": this file only moves data
 * between R objects and the plain-C ABI of include/bsgpu.h; all per-base work happens in
 * the CUDA kernels behind that ABI.
 *
 * It is written against the public C interfaces Biostrings itself uses (Rdefines.h,
 * XVector_interface.h, IRanges_interface.h), so it compiles against real R headers.  In
 * this repository (no R in the image) it is compiled against the stand-in headers of
 * oracle/shim/ and driven by oracle/ref_harness.c -- the same harness that drives the
 * reference's own C -- so tests/test_gpu_dispatch.py compares the SEXP results of both
 * libraries entry point by entry point.
 *
 * Entry points (reference file:line):
 *   build_Twobit                      src/match_pdict_Twobit.c:104
 *   ACtree2_build                     src/match_pdict_ACtree2.c:792
 *   ACtree2_nodebuf_max_nblock / ACtree2_nodeextbuf_max_nblock / IntegerBAB_new
 *                                     src/match_pdict_ACtree2.c:217,320; src/BAB_class.c:9
 *   ACtree2_nnodes / ACtree2_has_all_flinks / ACtree2_compute_all_flinks
 *                                     src/match_pdict_ACtree2.c:681,972,981
 *   match_PDict3Parts_XString         src/match_pdict.c:153
 *   match_PDict3Parts_XStringViews    src/match_pdict.c:225
 *   vmatch_PDict3Parts_XStringSet     src/match_pdict.c:484
 *   ByPos_MIndex_endIndex / ByPos_MIndex_combine    src/MIndex_class.c:134,230
 */
#include <Rdefines.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "XVector_interface.h"
#include "IRanges_interface.h"
#include "S4Vectors_interface.h"
#include "bsgpu.h"

/* ---- small helpers ------------------------------------------------------------------ */

typedef struct { unsigned char *concat; int *widths; int n; } FlatSet;

/* concat + widths of an XStringSet (its elements may be scattered in the XVector pool) */
static FlatSet flatten(SEXP x)
{
	FlatSet f;
	XVectorList_holder h = hold_XVectorList(x);
	size_t total = 0, off = 0;
	int i;

	f.n = get_length_from_XVectorList_holder(&h);
	f.widths = (int *) R_alloc((size_t) (f.n ? f.n : 1), sizeof(int));
	for (i = 0; i < f.n; i++)
		total += (size_t) (f.widths[i] = get_elt_from_XRawList_holder(&h, i).length);
	f.concat = (unsigned char *) R_alloc(total ? total : 1, 1);
	for (i = 0; i < f.n; i++) {
		Chars_holder e = get_elt_from_XRawList_holder(&h, i);

		memcpy(f.concat + off, e.ptr, (size_t) e.length);
		off += (size_t) e.length;
	}
	return f;
}

static int matches_as_code(SEXP matches_as)
{
	const char *s = CHAR(STRING_ELT(matches_as, 0));

	/* src/match_reporting.c:10-31 */
	if (strcmp(s, "MATCHES_AS_NULL") == 0) return BSGPU_MATCHES_AS_NULL;
	if (strcmp(s, "MATCHES_AS_WHICH") == 0) return BSGPU_MATCHES_AS_WHICH;
	if (strcmp(s, "MATCHES_AS_COUNTS") == 0) return BSGPU_MATCHES_AS_COUNTS;
	if (strcmp(s, "MATCHES_AS_ENDS") == 0) return BSGPU_MATCHES_AS_ENDS;
	error("Biostrings internal error in _get_match_storing_code(): "
	      "\"%s\": unknown match storing mode", s);
	return -1;
}

/* ---- the device index lives in the PreprocessedTB object ------------------------------------
 * The Trusted Band is preprocessed once (build_Twobit / ACtree2_build) while head and tail
 * arrive with every match call.  The device index is created at build time and hung on an
 * external pointer the object already owns -- no new slot, the R classes stay as they are:
 *   ACtree2: pptb@nodebuf_ptr@xp   (IntegerBAB_new makes it with a NULL address, src/BAB_class.c:9-22)
 *   Twobit : pptb@sign2pos@shared@xp   (the SharedInteger behind the XInteger, NULL address too)
 * with a C finalizer, so the index is freed when R garbage-collects the PDict, copies of the
 * object share it, and no table keyed by SEXP addresses can go stale.  An object whose pointer
 * is NULL (readRDS, or built by the CPU package) is rebuilt from its own slots on first use:
 * the patterns the build skipped (pp_exclude) are recovered from what the object records --
 * the leaf nodes / sign2pos entries name the Trusted-Band leaders, pptb@dups the duplicates. */

struct dev_index {
	bsgpu_pdict3parts *p3;
	uint64_t flank_sig;		/* which head/tail are attached */
	int dups_blanked;
};

static int live_indexes = 0;
int bsgpu_dispatch_live_indexes(void) { return live_indexes; }

static void dev_index_finalizer(SEXP xp)
{
	struct dev_index *d = (struct dev_index *) R_ExternalPtrAddr(xp);

	if (d == NULL)
		return;
	bsgpu_pdict3parts_free(d->p3);
	free(d);
	R_SetExternalPtrAddr(xp, NULL);
	live_indexes--;
}

static void dev_index_attach(SEXP xp, bsgpu_pdict3parts *p3)
{
	struct dev_index *d = (struct dev_index *) calloc(1, sizeof(*d));

	if (d == NULL) {
		bsgpu_pdict3parts_free(p3);
		error("out of memory");
	}
	dev_index_finalizer(xp);		/* a rebuilt object: drop what it held */
	d->p3 = p3;
	R_SetExternalPtrAddr(xp, d);
	R_RegisterCFinalizerEx(xp, dev_index_finalizer, TRUE);
	live_indexes++;
}

static SEXP bab_xp(SEXP bab) { return GET_SLOT(bab, install("xp")); }
static SEXP xinteger_xp(SEXP x) { return GET_SLOT(GET_SLOT(x, install("shared")), install("xp")); }

static uint64_t fnv(uint64_t h, const void *p, size_t n)
{
	const unsigned char *b = p;
	size_t i;

	for (i = 0; i < n; i++)
		h = (h ^ b[i]) * 0x100000001B3ull;
	return h;
}

/* SEXP slot access, as src/PreprocessedTB_class.c:23-59 does it */
static SEXP pptb_tb(SEXP pptb) { return GET_SLOT(pptb, install("tb")); }
static SEXP pptb_dups(SEXP pptb) { return GET_SLOT(pptb, install("dups")); }

static int algo_of(SEXP pptb)
{
	const char *type = get_classname(pptb);

	if (strcmp(type, "Twobit") == 0) return BSGPU_ALGO_TWOBIT;
	if (strcmp(type, "ACtree2") == 0) return BSGPU_ALGO_ACTREE2;
	error("%s: unsupported Trusted Band type in 'pdict'", type);
	return -1;
}

/* the external pointer of 'pptb' that carries the device index */
static SEXP pptb_xp(SEXP pptb)
{
	if (algo_of(pptb) == BSGPU_ALGO_TWOBIT)
		return xinteger_xp(GET_SLOT(pptb, install("sign2pos")));
	return bab_xp(GET_SLOT(pptb, install("nodebuf_ptr")));
}

#define ISLEAF_BIT (1 << 30)		/* src/match_pdict_ACtree2.c:389-392 */
#define MAX_P_ID (ISLEAF_BIT - 1)

/* Rebuilds the device index of an object that does not carry one.  pp_exclude[i] = NA for the
 * patterns the original build indexed: the Trusted-Band leaders (Twobit: the values of sign2pos;
 * ACtree2: the P_ids of the leaf nodes in the node buffer, two ints per node) and their
 * duplicates (non-NA in pptb@dups); every other pattern was excluded (or is a duplicate of a
 * blanked dups slot, which the C level never reports either). */
static struct dev_index *rebuild_index(SEXP pptb, SEXP xp)
{
	SEXP tb = pptb_tb(pptb), h2l = get_H2LGrouping_high2low(pptb_dups(pptb));
	FlatSet t = flatten(tb);
	int *excl = (int *) R_alloc((size_t) (t.n ? t.n : 1), sizeof(int)), i, algo = algo_of(pptb);
	bsgpu_pdict3parts *p3;

	for (i = 0; i < t.n; i++)
		excl[i] = 1;
	for (i = 0; i < LENGTH(h2l) && i < t.n; i++)
		if (INTEGER(h2l)[i] != NA_INTEGER)
			excl[i] = NA_INTEGER;
	if (algo == BSGPU_ALGO_TWOBIT) {
		SEXP tag = R_ExternalPtrTag(xp);

		for (i = 0; i < LENGTH(tag); i++) {
			int id = INTEGER(tag)[i];

			if (id != NA_INTEGER && id >= 1 && id <= t.n)
				excl[id - 1] = NA_INTEGER;
		}
	} else {
		SEXP blocks = R_ExternalPtrTag(xp), prot = R_ExternalPtrProtected(xp);
		int nblock = INTEGER(prot)[0], last_nelt = INTEGER(prot)[1], b, k;

		for (b = 0; b < nblock; b++) {
			SEXP block = VECTOR_ELT(blocks, b);
			int nnodes = b == nblock - 1 ? last_nelt : LENGTH(block) / 2;

			for (k = 0; k < nnodes; k++) {
				int attribs = INTEGER(block)[2 * k];

				if ((attribs & ISLEAF_BIT) && attribs > 0) {
					int id = attribs & MAX_P_ID;

					if (id >= 1 && id <= t.n)
						excl[id - 1] = NA_INTEGER;
				}
			}
		}
	}
	if (bsgpu_pdict3parts_build(algo, t.n, t.concat, t.widths, NULL, NULL, NULL, NULL,
				    excl, NULL, &p3) != BSGPU_OK)
		error("%s", bsgpu_last_error());
	dev_index_attach(xp, p3);
	return (struct dev_index *) R_ExternalPtrAddr(xp);
}

/* The device index of (pptb, head, tail), ready for matching. */
static bsgpu_pdict3parts *get_p3(SEXP pptb, SEXP pdict_head, SEXP pdict_tail)
{
	SEXP xp = pptb_xp(pptb);
	struct dev_index *e = (struct dev_index *) R_ExternalPtrAddr(xp);
	FlatSet hd = {NULL, NULL, 0}, tl = {NULL, NULL, 0};
	uint64_t sig = 0xCBF29CE484222325ull;
	int i, all_na = 1;
	SEXP h2l;

	if (e == NULL)
		e = rebuild_index(pptb, xp);
	if (pdict_head != R_NilValue) {
		size_t tot = 0;

		hd = flatten(pdict_head);
		for (i = 0; i < hd.n; i++)
			tot += (size_t) hd.widths[i];
		sig = fnv(fnv(sig, hd.widths, (size_t) hd.n * sizeof(int)), hd.concat, tot);
	}
	sig = fnv(sig, "|", 1);
	if (pdict_tail != R_NilValue) {
		size_t tot = 0;

		tl = flatten(pdict_tail);
		for (i = 0; i < tl.n; i++)
			tot += (size_t) tl.widths[i];
		sig = fnv(fnv(sig, tl.widths, (size_t) tl.n * sizeof(int)), tl.concat, tot);
	}
	if (sig == 0)
		sig = 1;
	if (e->flank_sig != sig) {
		if (bsgpu_pdict3parts_set_flanks(e->p3, hd.concat, hd.widths, tl.concat, tl.widths) != BSGPU_OK)
			error("%s", bsgpu_last_error());
		e->flank_sig = sig;
	}
	/* R/PDict-class.R:224-226: a reused whole-pattern pptb0 has its dups blanked at the R
	   level; the C side sees it as low2high being all NULL */
	h2l = get_H2LGrouping_high2low(pptb_dups(pptb));
	for (i = 0; i < LENGTH(h2l); i++)
		if (INTEGER(h2l)[i] != NA_INTEGER) {
			all_na = 0;
			break;
		}
	if (all_na && !e->dups_blanked) {
		if (bsgpu_pdict3parts_blank_dups(e->p3) != BSGPU_OK)
			error("%s", bsgpu_last_error());
		e->dups_blanked = 1;
	}
	return e->p3;
}

/* test hook: forget the device index of an object, as readRDS() would hand it back */
SEXP bsgpu_dispatch_detach(SEXP pptb)
{
	dev_index_finalizer(pptb_xp(pptb));
	return R_NilValue;
}

/* ---- preprocessing entry points -------------------------------------------------------- */

static SEXP build_common(int algo, SEXP tb, SEXP pp_exclude, SEXP first_elt, const char *first_name,
			 bsgpu_pdict3parts **p3_out)
{
	FlatSet t = flatten(tb);
	SEXP ans, ans_names, h2l;
	bsgpu_pdict3parts *p3;

	PROTECT(h2l = NEW_INTEGER(t.n));
	if (bsgpu_pdict3parts_build(algo, t.n, t.concat, t.widths, NULL, NULL, NULL, NULL,
				    pp_exclude == R_NilValue ? NULL : INTEGER(pp_exclude),
				    INTEGER(h2l), &p3) != BSGPU_OK)
		error("%s", bsgpu_last_error());	/* the reference's own message */
	*p3_out = p3;
	PROTECT(ans = NEW_LIST(2));
	PROTECT(ans_names = NEW_CHARACTER(2));
	SET_STRING_ELT(ans_names, 0, mkChar(first_name));
	SET_STRING_ELT(ans_names, 1, mkChar("high2low"));
	SET_NAMES(ans, ans_names);
	SET_VECTOR_ELT(ans, 0, first_elt);
	SET_VECTOR_ELT(ans, 1, h2l);
	UNPROTECT(3);
	return ans;
}

/* --- .Call ENTRY POINT --- (src/match_pdict_Twobit.c:104) */
SEXP build_Twobit(SEXP tb, SEXP pp_exclude, SEXP base_codes)
{
	/* Twobit@sign2pos must stay a valid XInteger of length 4^w: show() prints its length
	   (R/PDict-class.R:79-93,102).  It is filled the way the reference fills it
	   (signature = first letter most significant, src/utils.c:135-142). */
	XVectorList_holder h = hold_XVectorList(tb);
	int n = get_length_from_XVectorList_holder(&h), w = -1, i, d;
	SEXP sign2pos, xint, ans;
	bsgpu_pdict3parts *p3 = NULL;

	for (i = 0; i < n && w < 0; i++)
		if (pp_exclude == R_NilValue || INTEGER(pp_exclude)[i] == NA_INTEGER)
			w = get_elt_from_XRawList_holder(&h, i).length;
	/* argument errors (empty TB, w > 14, ragged, non-base) come from the library, with
	   the reference's text, before any table is allocated */
	PROTECT(ans = build_common(BSGPU_ALGO_TWOBIT, tb, pp_exclude, R_NilValue, "sign2pos", &p3));
	PROTECT(sign2pos = NEW_INTEGER(w < 0 ? 0 : 1 << (2 * w)));
	for (i = 0; i < LENGTH(sign2pos); i++)
		INTEGER(sign2pos)[i] = NA_INTEGER;
	for (i = 0; i < n; i++) {
		Chars_holder e;
		int sig = 0;

		if (pp_exclude != R_NilValue && INTEGER(pp_exclude)[i] != NA_INTEGER)
			continue;
		e = get_elt_from_XRawList_holder(&h, i);
		for (d = 0; d < w; d++) {
			int b = (unsigned char) e.ptr[d];

			sig = (sig << 2) | (b == 1 ? 0 : b == 2 ? 1 : b == 4 ? 2 : 3);
		}
		if (INTEGER(sign2pos)[sig] == NA_INTEGER)
			INTEGER(sign2pos)[sig] = i + 1;
	}
	PROTECT(xint = new_XInteger_from_tag("XInteger", sign2pos));
	dev_index_attach(xinteger_xp(xint), p3);
	SET_VECTOR_ELT(ans, 0, xint);
	UNPROTECT(3);
	(void) base_codes;
	return ans;
}

/* --- .Call ENTRY POINT --- (src/match_pdict_ACtree2.c:792) */
SEXP ACtree2_build(SEXP tb, SEXP pp_exclude, SEXP base_codes, SEXP nodebuf_ptr, SEXP nodeextbuf_ptr)
{
	bsgpu_pdict3parts *p3 = NULL;
	SEXP ans, h2l, xp = bab_xp(nodebuf_ptr), blocks, prot, block;
	int n, i, nleaves = 0;

	(void) base_codes; (void) nodeextbuf_ptr;
	PROTECT(ans = build_common(BSGPU_ALGO_ACTREE2, tb, pp_exclude, R_NilValue, "ACtree", &p3));
	dev_index_attach(xp, p3);
	/* The trie itself is not built, but its leaf nodes are written to the node buffer the way
	   the reference writes them (attribs = ISLEAF_BIT | P_id, src/match_pdict_ACtree2.c:389-392,
	   750-781): they are what names the indexed patterns when the object comes back from disk
	   without its device index (rebuild_index). */
	h2l = VECTOR_ELT(ans, 1);
	n = LENGTH(h2l);
	for (i = 0; i < n; i++)
		if (INTEGER(h2l)[i] == NA_INTEGER && (pp_exclude == R_NilValue || INTEGER(pp_exclude)[i] == NA_INTEGER))
			nleaves++;
	blocks = R_ExternalPtrTag(xp);
	prot = R_ExternalPtrProtected(xp);
	if (nleaves > 0 && LENGTH(blocks) > 0) {
		int k = 0;

		PROTECT(block = NEW_INTEGER(2 * nleaves));
		for (i = 0; i < n; i++)
			if (INTEGER(h2l)[i] == NA_INTEGER && (pp_exclude == R_NilValue || INTEGER(pp_exclude)[i] == NA_INTEGER)) {
				INTEGER(block)[2 * k] = ISLEAF_BIT | (i + 1);
				INTEGER(block)[2 * k + 1] = -1;
				k++;
			}
		SET_ELEMENT(blocks, 0, block);
		UNPROTECT(1);
		INTEGER(prot)[0] = 1;
		INTEGER(prot)[1] = nleaves;
	}
	UNPROTECT(1);
	return ans;
}

/* The node buffers of the CPU tree are not needed: the index lives on the device.  The
 * accessors stay so that the R object remains constructible and printable. */
SEXP ACtree2_nodebuf_max_nblock(void) { return ScalarInteger(1024); }
SEXP ACtree2_nodeextbuf_max_nblock(void) { return ScalarInteger(1024); }

SEXP IntegerBAB_new(SEXP max_nblock)
{
	SEXP blocks, prot, xp, classdef, ans;

	PROTECT(blocks = NEW_LIST(INTEGER(max_nblock)[0]));
	PROTECT(prot = NEW_INTEGER(2));
	INTEGER(prot)[0] = INTEGER(prot)[1] = 0;
	PROTECT(xp = R_MakeExternalPtr(NULL, blocks, prot));
	PROTECT(classdef = MAKE_CLASS("IntegerBAB"));
	PROTECT(ans = NEW_OBJECT(classdef));
	SET_SLOT(ans, mkChar("xp"), xp);
	UNPROTECT(5);
	return ans;
}

SEXP ACtree2_nnodes(SEXP pptb)
{
	/* no CPU tree: report the number of Trusted-Band letters indexed, an upper bound of
	   the node count of the reference's trie */
	struct dev_index *e = (struct dev_index *) R_ExternalPtrAddr(pptb_xp(pptb));

	return ScalarInteger(e ? bsgpu_pdict3parts_length(e->p3) * bsgpu_pdict3parts_tb_width(e->p3) + 1
			       : NA_INTEGER);
}

SEXP ACtree2_has_all_flinks(SEXP pptb) { (void) pptb; return ScalarLogical(1); }
SEXP ACtree2_compute_all_flinks(SEXP pptb) { (void) pptb; return R_NilValue; }
void _init_bytewise_match_tables(void) {}	/* the match tables are bit tests in the kernels */

/* ---- results ------------------------------------------------------------------------------ */

static SEXP csr_as_list(const bsgpu_result *r)
{
	SEXP ans, elt;
	int64_t i;

	/* new_LIST_from_IntAEAE(..., 1): NULL, not integer(0), for empty rows
	   (src/match_reporting.c:192-201) */
	PROTECT(ans = NEW_LIST((int) r->n_rows));
	for (i = 0; i < r->n_rows; i++) {
		int64_t n = r->ends_offsets[i + 1] - r->ends_offsets[i];

		if (n == 0)
			continue;
		PROTECT(elt = NEW_INTEGER((int) n));
		memcpy(INTEGER(elt), r->ends + r->ends_offsets[i], (size_t) n * sizeof(int));
		SET_VECTOR_ELT(ans, (int) i, elt);
		UNPROTECT(1);
	}
	UNPROTECT(1);
	return ans;
}

static SEXP result_as_SEXP(bsgpu_result *r, int ms_code)
{
	SEXP ans;

	switch (ms_code) {
	case BSGPU_MATCHES_AS_NULL:
		ans = R_NilValue;
		break;
	case BSGPU_MATCHES_AS_WHICH:	/* _MatchBuf_which_asINTEGER, src/match_reporting.c:150 */
		PROTECT(ans = NEW_INTEGER((int) r->n_which));
		memcpy(INTEGER(ans), r->which, (size_t) r->n_which * sizeof(int));
		UNPROTECT(1);
		break;
	case BSGPU_MATCHES_AS_COUNTS:	/* _MatchBuf_counts_asINTEGER, src/match_reporting.c:163 */
		PROTECT(ans = NEW_INTEGER((int) r->n_counts));
		memcpy(INTEGER(ans), r->counts, (size_t) r->n_counts * sizeof(int));
		UNPROTECT(1);
		break;
	default:			/* _MatchBuf_ends_asLIST, src/match_reporting.c:192 */
		ans = csr_as_list(r);
	}
	bsgpu_result_free(r);
	return ans;
}

/* ---- matching entry points ------------------------------------------------------------------ */

/* --- .Call ENTRY POINT --- (src/match_pdict.c:153) */
SEXP match_PDict3Parts_XString(SEXP pptb, SEXP pdict_head, SEXP pdict_tail,
		SEXP subject, SEXP max_mismatch, SEXP min_mismatch, SEXP fixed,
		SEXP matches_as, SEXP envir)
{
	Chars_holder S = hold_XRaw(subject);
	int ms_code = matches_as_code(matches_as);
	bsgpu_result r;

	(void) envir;
	if (bsgpu_match_PDict3Parts_XString(get_p3(pptb, pdict_head, pdict_tail),
			(const uint8_t *) S.ptr, S.length,
			INTEGER(max_mismatch)[0], INTEGER(min_mismatch)[0],
			LOGICAL(fixed)[0], LOGICAL(fixed)[1], ms_code, &r) != BSGPU_OK)
		error("%s", bsgpu_last_error());
	return result_as_SEXP(&r, ms_code);
}

/* --- .Call ENTRY POINT --- (src/match_pdict.c:225) */
SEXP match_PDict3Parts_XStringViews(SEXP pptb, SEXP pdict_head, SEXP pdict_tail,
		SEXP subject, SEXP views_start, SEXP views_width,
		SEXP max_mismatch, SEXP min_mismatch, SEXP fixed,
		SEXP matches_as, SEXP envir)
{
	Chars_holder S = hold_XRaw(subject);
	int ms_code = matches_as_code(matches_as);
	bsgpu_result r;

	(void) envir;
	if (bsgpu_match_PDict3Parts_XStringViews(get_p3(pptb, pdict_head, pdict_tail),
			(const uint8_t *) S.ptr, S.length,
			INTEGER(views_start), INTEGER(views_width), LENGTH(views_start),
			INTEGER(max_mismatch)[0], INTEGER(min_mismatch)[0],
			LOGICAL(fixed)[0], LOGICAL(fixed)[1], ms_code, &r) != BSGPU_OK)
		error("%s", bsgpu_last_error());
	return result_as_SEXP(&r, ms_code);
}

/* vmatch on a sequence set for either kind of dictionary handle */
static SEXP vmatch_common(bsgpu_pdict3parts *p3, int M, const char *entry, SEXP subject,
		SEXP max_mismatch, SEXP min_mismatch, SEXP fixed,
		SEXP collapse, SEXP weight, SEXP matches_as)
{
	FlatSet s = flatten(subject);
	int ms_code = matches_as_code(matches_as), coll = INTEGER(collapse)[0];
	bsgpu_result r;
	SEXP ans;

	if (ms_code == BSGPU_MATCHES_AS_NULL)
		error("%s() does not support 'matches_as=\"%s\"' yet, sorry", entry,
		      CHAR(STRING_ELT(matches_as, 0)));
	if (ms_code != BSGPU_MATCHES_AS_WHICH && ms_code != BSGPU_MATCHES_AS_COUNTS)
		error("vmatchPDict() is not supported yet, sorry");
	if (ms_code == BSGPU_MATCHES_AS_COUNTS && coll != 0 && coll != 1 && coll != 2)
		error("'collapse' must be FALSE, 1 or 2");
	if (bsgpu_vmatch_PDict3Parts_XStringSet(p3,
			s.concat, s.widths, s.n,
			INTEGER(max_mismatch)[0], INTEGER(min_mismatch)[0],
			LOGICAL(fixed)[0], LOGICAL(fixed)[1],
			ms_code == BSGPU_MATCHES_AS_COUNTS ? coll : 0, !IS_INTEGER(weight),
			IS_INTEGER(weight) ? (const void *) INTEGER(weight) : (const void *) REAL(weight),
			LENGTH(weight), ms_code, &r) != BSGPU_OK)
		error("%s", bsgpu_last_error());
	if (ms_code == BSGPU_MATCHES_AS_WHICH) {
		/* list of N integer vectors, integer(0) when empty (src/match_pdict.c:333-343) */
		int64_t j;

		PROTECT(ans = NEW_LIST(s.n));
		for (j = 0; j < s.n; j++) {
			int64_t n = r.ends_offsets[j + 1] - r.ends_offsets[j];
			SEXP elt = PROTECT(NEW_INTEGER((int) n));

			memcpy(INTEGER(elt), r.ends + r.ends_offsets[j], (size_t) n * sizeof(int));
			SET_VECTOR_ELT(ans, (int) j, elt);
			UNPROTECT(1);
		}
		UNPROTECT(1);
	} else if (coll == 0) {
		PROTECT(ans = allocMatrix(INTSXP, M, s.n));
		memcpy(INTEGER(ans), r.counts, (size_t) M * (size_t) s.n * sizeof(int));
		UNPROTECT(1);
	} else if (r.dcounts != NULL) {
		PROTECT(ans = NEW_NUMERIC((int) r.n_counts));
		memcpy(REAL(ans), r.dcounts, (size_t) r.n_counts * sizeof(double));
		UNPROTECT(1);
	} else {
		PROTECT(ans = NEW_INTEGER((int) r.n_counts));
		memcpy(INTEGER(ans), r.counts, (size_t) r.n_counts * sizeof(int));
		UNPROTECT(1);
	}
	bsgpu_result_free(&r);
	return ans;
}

/* --- .Call ENTRY POINT --- (src/match_pdict.c:484) */
SEXP vmatch_PDict3Parts_XStringSet(SEXP pptb, SEXP pdict_head, SEXP pdict_tail,
		SEXP subject, SEXP max_mismatch, SEXP min_mismatch, SEXP fixed,
		SEXP collapse, SEXP weight, SEXP matches_as, SEXP envir)
{
	(void) envir;
	return vmatch_common(get_p3(pptb, pdict_head, pdict_tail),
			get_XVectorList_length(pptb_tb(pptb)), "vmatch_PDict3Parts_XStringSet",
			subject, max_mismatch, min_mismatch, fixed, collapse, weight, matches_as);
}

/* ---- non-preprocessed dictionaries: 'pattern' is a plain XStringSet --------------------------
 * The reference runs one matchPattern() per pattern with the algorithm R selected; all of them
 * report the same matches, so 'algorithm' only needs to be a known name here.  with.indels
 * (algorithm "indels", counts / which only, R/matchPDict.R:171-191) goes to bsgpu_count_indels. */

static int wants_indels(SEXP with_indels, SEXP algorithm)
{
	return LOGICAL(with_indels)[0] || strcmp(CHAR(STRING_ELT(algorithm, 0)), "indels") == 0;
}

static bsgpu_pdict3parts *plain_handle(SEXP pattern, SEXP with_indels, SEXP algorithm)
{
	static const char *known[] = {"naive-exact", "naive-inexact", "boyer-moore", "shift-or", "indels", NULL};
	const char *algo = CHAR(STRING_ELT(algorithm, 0));
	FlatSet p = flatten(pattern);
	bsgpu_pdict3parts *p3 = NULL;
	int k;

	(void) with_indels;
	for (k = 0; known[k] != NULL && strcmp(known[k], algo) != 0; k++)
		;
	if (known[k] == NULL)
		error("\"%s\": unknown algorithm", algo);
	if (bsgpu_patterns_build(p.concat, p.widths, p.n, &p3) != BSGPU_OK)
		error("%s", bsgpu_last_error());
	return p3;
}

/* with.indels = TRUE: the counts behind countPDict / whichPDict (src/match_pdict.c:174-199 with
 * algo "indels" -> _match_pattern_indels, src/match_pattern_indels.c:58-107).  Frees p3 and the
 * subject handle.  per_sequence: M x N column-major, else M. */
static int *indels_counts(bsgpu_pdict3parts *p3, bsgpu_subject *subj, int rc_subj, int M, int N,
			  SEXP max_mismatch, SEXP fixed, int per_sequence, int ms_code)
{
	int *counts = (int *) R_alloc((size_t) M * (size_t) (per_sequence ? N : 1) + 1, sizeof(int)), rc = rc_subj;

	if (rc == BSGPU_OK && ms_code != BSGPU_MATCHES_AS_WHICH && ms_code != BSGPU_MATCHES_AS_COUNTS) {
		bsgpu_subject_free(subj);
		bsgpu_pdict3parts_free(p3);
		error("at the moment, within the matchPDict family, only countPDict(), whichPDict(), "
		      "vcountPDict() and vwhichPDict() support indels");
	}
	if (rc == BSGPU_OK)
		rc = bsgpu_count_indels(p3, subj, INTEGER(max_mismatch)[0], LOGICAL(fixed)[0], LOGICAL(fixed)[1],
					per_sequence, counts);
	if (subj != NULL)
		bsgpu_subject_free(subj);
	bsgpu_pdict3parts_free(p3);
	if (rc != BSGPU_OK)
		error("%s", bsgpu_last_error());
	return counts;
}

static SEXP counts_or_which(const int *counts, int M, int ms_code)
{
	SEXP ans;
	int i, n = 0;

	if (ms_code == BSGPU_MATCHES_AS_COUNTS) {
		PROTECT(ans = NEW_INTEGER(M));
		memcpy(INTEGER(ans), counts, (size_t) M * sizeof(int));
		UNPROTECT(1);
		return ans;
	}
	for (i = 0; i < M; i++)
		n += counts[i] != 0;
	PROTECT(ans = NEW_INTEGER(n));
	for (i = 0, n = 0; i < M; i++)
		if (counts[i] != 0)
			INTEGER(ans)[n++] = i + 1;	/* sort(ids) + 1, src/match_reporting.c:150-161 */
	UNPROTECT(1);
	return ans;
}

/* --- .Call ENTRY POINT --- (src/match_pdict.c:174) */
SEXP match_XStringSet_XString(SEXP pattern, SEXP subject,
		SEXP max_mismatch, SEXP min_mismatch, SEXP with_indels, SEXP fixed,
		SEXP algorithm, SEXP matches_as, SEXP envir)
{
	Chars_holder S = hold_XRaw(subject);
	int ms_code = matches_as_code(matches_as), rc;
	bsgpu_pdict3parts *p3 = plain_handle(pattern, with_indels, algorithm);
	bsgpu_result r;

	(void) envir;
	if (wants_indels(with_indels, algorithm)) {
		bsgpu_subject *subj = NULL;
		int M = get_XVectorList_length(pattern);

		rc = bsgpu_subject_xstring((const uint8_t *) S.ptr, S.length, &subj);
		return counts_or_which(indels_counts(p3, subj, rc, M, 1, max_mismatch, fixed, 0, ms_code), M, ms_code);
	}
	rc = bsgpu_match_PDict3Parts_XString(p3, (const uint8_t *) S.ptr, S.length,
			INTEGER(max_mismatch)[0], INTEGER(min_mismatch)[0],
			LOGICAL(fixed)[0], LOGICAL(fixed)[1], ms_code, &r);
	bsgpu_pdict3parts_free(p3);
	if (rc != BSGPU_OK)
		error("%s", bsgpu_last_error());
	return result_as_SEXP(&r, ms_code);
}

/* --- .Call ENTRY POINT --- (src/match_pdict.c:267) */
SEXP match_XStringSet_XStringViews(SEXP pattern, SEXP subject, SEXP views_start, SEXP views_width,
		SEXP max_mismatch, SEXP min_mismatch, SEXP with_indels, SEXP fixed,
		SEXP algorithm, SEXP matches_as, SEXP envir)
{
	Chars_holder S = hold_XRaw(subject);
	int ms_code = matches_as_code(matches_as), rc;
	bsgpu_pdict3parts *p3 = plain_handle(pattern, with_indels, algorithm);
	bsgpu_result r;

	(void) envir;
	if (wants_indels(with_indels, algorithm)) {
		bsgpu_subject *subj = NULL;
		int M = get_XVectorList_length(pattern);

		rc = bsgpu_subject_views((const uint8_t *) S.ptr, S.length, INTEGER(views_start), INTEGER(views_width),
					 LENGTH(views_start), &subj);
		return counts_or_which(indels_counts(p3, subj, rc, M, 1, max_mismatch, fixed, 0, ms_code), M, ms_code);
	}
	rc = bsgpu_match_PDict3Parts_XStringViews(p3, (const uint8_t *) S.ptr, S.length,
			INTEGER(views_start), INTEGER(views_width), LENGTH(views_start),
			INTEGER(max_mismatch)[0], INTEGER(min_mismatch)[0],
			LOGICAL(fixed)[0], LOGICAL(fixed)[1], ms_code, &r);
	bsgpu_pdict3parts_free(p3);
	if (rc != BSGPU_OK)
		error("%s", bsgpu_last_error());
	return result_as_SEXP(&r, ms_code);
}

/* --- .Call ENTRY POINT --- (src/match_pdict.c:519) */
SEXP vmatch_XStringSet_XStringSet(SEXP pattern, SEXP subject,
		SEXP max_mismatch, SEXP min_mismatch, SEXP with_indels, SEXP fixed,
		SEXP algorithm, SEXP collapse, SEXP weight, SEXP matches_as, SEXP envir)
{
	bsgpu_pdict3parts *p3 = plain_handle(pattern, with_indels, algorithm);
	SEXP ans;

	(void) envir;
	if (wants_indels(with_indels, algorithm)) {
		FlatSet s = flatten(subject);
		bsgpu_subject *subj = NULL;
		int M = get_XVectorList_length(pattern), N = s.n, ms_code = matches_as_code(matches_as);
		int coll = ms_code == BSGPU_MATCHES_AS_COUNTS ? INTEGER(collapse)[0] : 0, i, j;
		int rc = bsgpu_subject_set(s.concat, s.widths, s.n, &subj);
		const int *mat = indels_counts(p3, subj, rc, M, N, max_mismatch, fixed, 1, ms_code);

		if (ms_code == BSGPU_MATCHES_AS_WHICH) {
			PROTECT(ans = NEW_LIST(N));
			for (j = 0; j < N; j++)
				SET_VECTOR_ELT(ans, j, counts_or_which(mat + (size_t) j * M, M, BSGPU_MATCHES_AS_WHICH));
		} else if (coll == 0) {
			PROTECT(ans = allocMatrix(INTSXP, M, N));
			memcpy(INTEGER(ans), mat, (size_t) M * (size_t) N * sizeof(int));
		} else if (IS_INTEGER(weight)) {
			/* src/match_pdict.c:85-127: subjects outer, patterns inner */
			PROTECT(ans = NEW_INTEGER(coll == 1 ? M : N));
			memset(INTEGER(ans), 0, (size_t) LENGTH(ans) * sizeof(int));
			for (j = 0; j < N; j++)
				for (i = 0; i < M; i++) {
					if (coll == 1)
						INTEGER(ans)[i] += mat[(size_t) j * M + i] * INTEGER(weight)[j];
					else
						INTEGER(ans)[j] += mat[(size_t) j * M + i] * INTEGER(weight)[i];
				}
		} else {
			PROTECT(ans = NEW_NUMERIC(coll == 1 ? M : N));
			memset(REAL(ans), 0, (size_t) LENGTH(ans) * sizeof(double));
			for (j = 0; j < N; j++)
				for (i = 0; i < M; i++) {
					if (coll == 1)
						REAL(ans)[i] += mat[(size_t) j * M + i] * REAL(weight)[j];
					else
						REAL(ans)[j] += mat[(size_t) j * M + i] * REAL(weight)[i];
				}
		}
		UNPROTECT(1);
		return ans;
	}
	/* an error() inside vmatch_common() leaks p3 until the process ends (longjmp), as any C
	 * allocation would in an R entry point that errors */
	ans = vmatch_common(p3, get_XVectorList_length(pattern), "vmatch_XStringSet_XStringSet",
			    subject, max_mismatch, min_mismatch, fixed, collapse, weight, matches_as);
	bsgpu_pdict3parts_free(p3);
	return ans;
}

/* ---- MIndex helpers (host side; same code shape as the reference's, they are O(hits)) ------- */

/* --- .Call ENTRY POINT --- (src/MIndex_class.c:134) */
SEXP ByPos_MIndex_endIndex(SEXP x_high2low, SEXP x_ends, SEXP x_width0)
{
	SEXP ans, ans_elt;
	int i, j, low;

	PROTECT(ans = duplicate(x_ends));
	for (i = 0; i < LENGTH(ans); i++) {
		if (x_high2low != R_NilValue && LENGTH(x_high2low) != 0
		 && (low = INTEGER(x_high2low)[i]) != NA_INTEGER) {
			PROTECT(ans_elt = duplicate(VECTOR_ELT(ans, low - 1)));
			SET_VECTOR_ELT(ans, i, ans_elt);
			UNPROTECT(1);
			continue;
		}
		if (x_width0 == R_NilValue)
			continue;
		ans_elt = VECTOR_ELT(ans, i);
		if (!IS_INTEGER(ans_elt))
			continue;
		for (j = 0; j < LENGTH(ans_elt); j++)
			INTEGER(ans_elt)[j] += 1 - INTEGER(x_width0)[i];
	}
	UNPROTECT(1);
	return ans;
}

static int cmp_int(const void *a, const void *b)
{
	int x = *(const int *) a, y = *(const int *) b;

	return (x > y) - (x < y);
}

/* --- .Call ENTRY POINT --- (src/MIndex_class.c:230) */
SEXP ByPos_MIndex_combine(SEXP ends_listlist)
{
	int NTB = LENGTH(ends_listlist), ans_length, i, j, n, k, u;
	SEXP ans, ans_elt, ends;
	int *buf = NULL, cap = 0;

	if (NTB == 0)
		error("nothing to combine");
	ans_length = LENGTH(VECTOR_ELT(ends_listlist, 0));
	for (j = 1; j < NTB; j++)
		if (LENGTH(VECTOR_ELT(ends_listlist, j)) != ans_length)
			error("cannot combine MIndex objects of different lengths");
	PROTECT(ans = NEW_LIST(ans_length));
	for (i = 0; i < ans_length; i++) {
		n = 0;
		for (j = 0; j < NTB; j++) {
			ends = VECTOR_ELT(VECTOR_ELT(ends_listlist, j), i);
			if (ends == R_NilValue)
				continue;
			if (n + LENGTH(ends) > cap) {
				cap = 2 * (n + LENGTH(ends));
				buf = (int *) realloc(buf, (size_t) cap * sizeof(int));
			}
			memcpy(buf + n, INTEGER(ends), (size_t) LENGTH(ends) * sizeof(int));
			n += LENGTH(ends);
		}
		if (n == 0)
			continue;
		qsort(buf, (size_t) n, sizeof(int), cmp_int);
		for (k = 1, u = 0; k < n; k++)
			if (buf[k] != buf[u])
				buf[++u] = buf[k];
		PROTECT(ans_elt = NEW_INTEGER(u + 1));
		memcpy(INTEGER(ans_elt), buf, (size_t) (u + 1) * sizeof(int));
		SET_VECTOR_ELT(ans, i, ans_elt);
		UNPROTECT(1);
	}
	free(buf);
	UNPROTECT(1);
	return ans;
}
