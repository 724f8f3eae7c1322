/*
 * letter_frequency_gpu.c -- the R-facing C dispatch of oligonucleotideFrequency() over libbsgpu.
 *
 * Drop-in for two .Call entry points of Biostrings' src/letter_frequency.c: SAME names, SAME
 * arguments, SAME result shapes (integer or double vector with names, N x 4^width matrix with
 * column names, 4 x ... x 4 array with dimnames, list of vectors), so R/letterFrequency.R:557-637
 * and src/R_init_Biostrings.c stay untouched.  The counting is done by oligo_freq_kernel behind
 * bsgpu_XString_oligo_frequency / bsgpu_XStringSet_oligo_frequency (include/bsgpu.h); this file
 * only moves data between R objects and that ABI and attaches the labels.
 *
 * Compiled here against the stand-in headers of oracle/shim/ and driven by
 * oracle/ref_harness.c next to the reference's own C (tests/test_gpu_dispatch.py).
 *
 * Entry points (reference file:line):
 *   XString_oligo_frequency           src/letter_frequency.c:634
 *   XStringSet_oligo_frequency        src/letter_frequency.c:667
 */
#include <Rdefines.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "XVector_interface.h"
#include "bsgpu.h"

/* 4^width words in table order (mk_all_oligos, src/letter_frequency.c:368-412) */
static SEXP all_oligos(int width, SEXP base_letters, int invert_twobit_order)
{
	int n = 1 << (2 * width), i, j, sign;
	char buf[16];
	SEXP ans;

	PROTECT(ans = NEW_CHARACTER(n));
	buf[width] = 0;
	for (i = 0; i < n; i++) {
		sign = i;
		for (j = 0; j < width; j++) {
			int at = invert_twobit_order ? j : width - 1 - j;

			buf[at] = CHAR(STRING_ELT(base_letters, sign & 3))[0];
			sign >>= 2;
		}
		SET_STRING_ELT(ans, i, mkChar(buf));
	}
	UNPROTECT(1);
	return ans;
}

/* names or dim + dimnames of one table (format_oligo_freqs, src/letter_frequency.c:414-430) */
static void format_table(SEXP x, int width, SEXP base_labels, int invert_twobit_order, int as_array)
{
	SEXP tmp;
	int i;

	if (as_array) {
		PROTECT(tmp = NEW_INTEGER(width));
		for (i = 0; i < width; i++)
			INTEGER(tmp)[i] = 4;
		SET_DIM(x, tmp);
		UNPROTECT(1);
		if (base_labels == R_NilValue)
			return;
		PROTECT(tmp = NEW_LIST(width));
		for (i = 0; i < width; i++)
			SET_VECTOR_ELT(tmp, i, duplicate(base_labels));
		SET_DIMNAMES(x, tmp);
		UNPROTECT(1);
		return;
	}
	if (base_labels == R_NilValue)
		return;
	PROTECT(tmp = all_oligos(width, base_labels, invert_twobit_order));
	SET_NAMES(x, tmp);
	UNPROTECT(1);
}

static SEXP new_table(int n, int as_integer)
{
	return as_integer ? NEW_INTEGER(n) : NEW_NUMERIC(n);
}

static void check_width(int width)
{
	/* src/utils.c:101-102, raised before anything is allocated */
	if (width < 1 || width > 15)
		error("_new_TwobitEncodingBuffer(): 'buflength' must be >= 1 and <= 15");
}

SEXP XString_oligo_frequency(SEXP x, SEXP width, SEXP step, SEXP as_prob, SEXP as_array,
		SEXP fast_moving_side, SEXP with_labels, SEXP base_codes)
{
	int width0 = INTEGER(width)[0], step0 = INTEGER(step)[0];
	int as_integer = !LOGICAL(as_prob)[0], as_array0 = LOGICAL(as_array)[0];
	int invert = strcmp(CHAR(STRING_ELT(fast_moving_side, 0)), "right") != 0;
	SEXP base_labels = LOGICAL(with_labels)[0] ? GET_NAMES(base_codes) : R_NilValue;
	Chars_holder X;
	SEXP ans;

	if (LENGTH(base_codes) != 4)
		error("_new_TwobitEncodingBuffer(): 'base_codes' must be of length 4");
	check_width(width0);
	X = hold_XRaw(x);
	PROTECT(ans = new_table(1 << (2 * width0), as_integer));
	if (bsgpu_XString_oligo_frequency((const uint8_t *) X.ptr, X.length, width0, step0, invert,
			!as_integer, as_integer ? INTEGER(ans) : NULL,
			as_integer ? NULL : REAL(ans)) != BSGPU_OK)
		error("%s", bsgpu_last_error());
	format_table(ans, width0, base_labels, invert, as_array0);
	UNPROTECT(1);
	return ans;
}

SEXP XStringSet_oligo_frequency(SEXP x, SEXP width, SEXP step, SEXP as_prob, SEXP as_array,
		SEXP fast_moving_side, SEXP with_labels, SEXP simplify_as, SEXP base_codes)
{
	int width0 = INTEGER(width)[0], step0 = INTEGER(step)[0];
	int as_integer = !LOGICAL(as_prob)[0], as_array0 = LOGICAL(as_array)[0];
	int invert = strcmp(CHAR(STRING_ELT(fast_moving_side, 0)), "right") != 0;
	const char *simplify_as0 = CHAR(STRING_ELT(simplify_as, 0));
	SEXP base_labels = LOGICAL(with_labels)[0] ? GET_NAMES(base_codes) : R_NilValue;
	int collapsed = strcmp(simplify_as0, "collapsed") == 0;
	int as_matrix = strcmp(simplify_as0, "matrix") == 0;
	XVectorList_holder h;
	unsigned char *concat;
	int *widths, n, ncol, i, j;
	size_t total = 0, off = 0;
	SEXP ans, tmp, elt;

	if (LENGTH(base_codes) != 4)
		error("_new_TwobitEncodingBuffer(): 'base_codes' must be of length 4");
	check_width(width0);
	ncol = 1 << (2 * width0);
	/* concat + widths of the set (its elements may be scattered in the XVector pool) */
	h = hold_XVectorList(x);
	n = get_length_from_XVectorList_holder(&h);
	widths = (int *) R_alloc((size_t) (n ? n : 1), sizeof(int));
	for (i = 0; i < n; i++)
		total += (size_t) (widths[i] = get_elt_from_XRawList_holder(&h, i).length);
	concat = (unsigned char *) R_alloc(total ? total : 1, 1);
	for (i = 0; i < n; i++) {
		Chars_holder e = get_elt_from_XRawList_holder(&h, i);

		memcpy(concat + off, e.ptr, (size_t) e.length);
		off += (size_t) e.length;
	}
	if (collapsed) {
		PROTECT(ans = new_table(ncol, as_integer));
		memset(as_integer ? (void *) INTEGER(ans) : (void *) REAL(ans), 0,
		       (size_t) ncol * (as_integer ? sizeof(int) : sizeof(double)));
	} else {
		PROTECT(ans = as_integer ? allocMatrix(INTSXP, n, ncol) : allocMatrix(REALSXP, n, ncol));
	}
	if (bsgpu_XStringSet_oligo_frequency(concat, widths, n, width0, step0, invert, collapsed,
			!as_integer, as_integer ? INTEGER(ans) : NULL,
			as_integer ? NULL : REAL(ans)) != BSGPU_OK)
		error("%s", bsgpu_last_error());
	if (collapsed) {
		format_table(ans, width0, base_labels, invert, as_array0);
		UNPROTECT(1);
		return ans;
	}
	if (as_matrix) {
		/* set_oligo_freqs_colnames(), src/letter_frequency.c:432-446 */
		if (base_labels != R_NilValue) {
			PROTECT(tmp = NEW_LIST(2));
			SET_VECTOR_ELT(tmp, 0, R_NilValue);
			SET_VECTOR_ELT(tmp, 1, all_oligos(width0, base_labels, invert));
			SET_DIMNAMES(ans, tmp);
			UNPROTECT(1);
		}
		UNPROTECT(1);
		return ans;
	}
	/* "list": the rows of the matrix, each formatted like a single table */
	PROTECT(tmp = NEW_LIST(n));
	for (i = 0; i < n; i++) {
		PROTECT(elt = new_table(ncol, as_integer));
		for (j = 0; j < ncol; j++) {
			if (as_integer)
				INTEGER(elt)[j] = INTEGER(ans)[i + (size_t) j * n];
			else
				REAL(elt)[j] = REAL(ans)[i + (size_t) j * n];
		}
		format_table(elt, width0, base_labels, invert, as_array0);
		SET_VECTOR_ELT(tmp, i, elt);
		UNPROTECT(1);
	}
	UNPROTECT(2);
	return tmp;
}
