/*
 * read_fasta_files_gpu.c -- the R-facing C dispatch of readDNAStringSet(format="fasta") over
 * libbsgpu.
 *
 * Drop-in for one .Call entry point of Biostrings' src/read_fasta_files.c: SAME name, SAME
 * arguments, SAME result (an XStringSet allocated with alloc_XRawList, names = the description
 * lines when use_names), SAME error and warning texts, so R/XStringSet-io.R stays untouched.
 * The text of every file is read through XVector's filexp_gets() into one buffer, cut to the
 * records nrec / skip / seek_first_rec select (the record counter runs across the files, as in
 * the reference), and handed to bsgpu_subject_fasta(): the line rules and the letter
 * translation run on the device (fasta_summary_kernel / fasta_emit_kernel), the letters come
 * back with bsgpu_subject_download().
 *
 * Compiled here against the stand-in headers of oracle/shim/ and driven by oracle/ref_harness.c
 * next to the reference's own C (tests/test_gpu_dispatch.py).
 *
 * Entry point (reference file:line):
 *   read_fasta_files                  src/read_fasta_files.c:412
 */
#include <Rdefines.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "XVector_interface.h"
#include "bsgpu.h"

typedef struct { char *p; size_t n, cap; } Buf;

static void buf_append(Buf *b, const void *src, size_t n)
{
	if (b->n + n + 1 > b->cap) {
		size_t cap = b->cap ? b->cap : 1 << 16;
		char *p;

		while (cap < b->n + n + 1)
			cap *= 2;
		p = realloc(b->p, cap);
		if (p == NULL) {
			free(b->p);
			error("read_fasta_files(): out of memory");
		}
		b->p = p;
		b->cap = cap;
	}
	if (n)
		memcpy(b->p + b->n, src, n);
	b->n += n;
}

static void buf_reserve(Buf *b, size_t extra)
{
	size_t n = b->n;

	buf_append(b, NULL, 0);
	if (b->n + extra + 1 > b->cap) {
		char *p = realloc(b->p, 2 * (b->n + extra + 1));

		if (p == NULL) {
			free(b->p);
			error("read_fasta_files(): out of memory");
		}
		b->p = p;
		b->cap = 2 * (b->n + extra + 1);
	}
	b->n = n;
}

/* the whole file behind `filexp`, from its current position */
static void slurp(SEXP filexp, Buf *text)
{
	static char iobuf[200002];
	int EOL_in_buf, rc;

	text->n = 0;
	while ((rc = filexp_gets(filexp, iobuf, (int) sizeof(iobuf), &EOL_in_buf)) != 0) {
		if (rc == -1) {
			free(text->p);
			error("read error while reading characters from a FASTA file");
		}
		buf_append(text, iobuf, strlen(iobuf));
	}
}

/* offset of the k-th (0-based) line that starts with '>' at or after `from`, or n */
static size_t nth_header(const char *t, size_t n, size_t from, long k)
{
	size_t i = from;

	while (i < n) {
		if (t[i] == '>' && (i == 0 || t[i - 1] == '\n')) {
			if (k == 0)
				return i;
			k--;
		}
		{
			const char *nl = memchr(t + i, '\n', n - i);

			if (nl == NULL)
				return n;
			i = (size_t) (nl - t) + 1;
		}
	}
	return n;
}

SEXP read_fasta_files(SEXP filexp_list, SEXP nrec, SEXP skip, SEXP seek_first_rec,
		SEXP use_names, SEXP elementType, SEXP lkup)
{
	int nrec0 = INTEGER(nrec)[0], skip0 = INTEGER(skip)[0];
	int seek0 = LOGICAL(seek_first_rec)[0], use_names0 = LOGICAL(use_names)[0];
	long recno = 0;			/* records seen so far, over all files */
	Buf text = {NULL, 0, 0}, letters = {NULL, 0, 0}, widths = {NULL, 0, 0}, descs = {NULL, 0, 0};
	Buf desc_lens = {NULL, 0, 0};
	char classname[40];
	int lkup_default[256], i, r, n_total = 0;
	const int *lk;
	int lk_len;
	SEXP width, names, ans;
	XVectorList_holder holder;
	size_t off = 0, doff = 0;

	if (lkup == R_NilValue) {	/* no translation: every byte is kept as it is */
		for (i = 0; i < 256; i++)
			lkup_default[i] = i;
		lk = lkup_default;
		lk_len = 256;
	} else {
		lk = INTEGER(lkup);
		lk_len = LENGTH(lkup);
	}
	for (i = 0; i < LENGTH(filexp_list); i++) {
		const char *filename = CHAR(STRING_ELT(GET_NAMES(filexp_list), i));
		bsgpu_subject *subj = NULL;
		bsgpu_fasta_info info;
		size_t first = 0, lo, hi;
		long nheaders = 0, want_from, want_to;
		int rc;

		slurp(VECTOR_ELT(filexp_list, i), &text);
		if (seek0) {		/* src/read_fasta_files.c:276-283 */
			first = nth_header(text.p, text.n, 0, 0);
			if (first == text.n && recno == 0 && i == LENGTH(filexp_list) - 1) {
				free(text.p); free(letters.p); free(widths.p); free(descs.p); free(desc_lens.p);
				error("reading FASTA file %s: no FASTA record found", filename);
			}
		}
		{			/* how many records the file holds */
			size_t p = first;

			while ((p = nth_header(text.p, text.n, p, 0)) < text.n) {
				nheaders++;
				p++;
			}
		}
		/* records [want_from, want_to) of this file are loaded (:297-313); the record
		   counter runs across the files */
		want_from = skip0 > recno ? skip0 - recno : 0;
		want_to = nrec0 >= 0 ? (long) skip0 + nrec0 - recno : nheaders;
		if (want_to > nheaders)
			want_to = nheaders;
		if (want_to <= want_from) {
			lo = hi = first;
		} else {
			lo = want_from > 0 ? nth_header(text.p, text.n, first, want_from) : first;
			hi = want_to < nheaders ? nth_header(text.p, text.n, first, want_to) : text.n;
		}
		recno += nheaders;
		rc = bsgpu_subject_fasta((const uint8_t *) text.p + lo, (int64_t) (hi - lo), lk, lk_len, &subj, &info);
		if (rc != BSGPU_OK) {
			free(text.p); free(letters.p); free(widths.p); free(descs.p); free(desc_lens.p);
			if (rc == BSGPU_EINVAL)
				error("reading FASTA file %s: %s", filename, bsgpu_last_error());
			error("%s", bsgpu_last_error());
		}
		if (info.ninvalid != 0)
			warning("reading FASTA file %s: ignored %lld invalid one-letter sequence codes",
				filename, (long long) info.ninvalid);
		/* letters, widths and description texts of this file behind those of the others */
		buf_reserve(&letters, (size_t) info.nletters);
		rc = bsgpu_subject_download(subj, (uint8_t *) letters.p + letters.n, NULL);
		bsgpu_subject_free(subj);
		if (rc != BSGPU_OK) {
			bsgpu_fasta_info_free(&info);
			free(text.p); free(letters.p); free(widths.p); free(descs.p); free(desc_lens.p);
			error("%s", bsgpu_last_error());
		}
		letters.n += (size_t) info.nletters;
		buf_append(&widths, info.widths, (size_t) info.nrec * sizeof(int32_t));
		for (r = 0; r < info.nrec; r++)
			buf_append(&descs, text.p + lo + info.desc_offset[r], (size_t) info.desc_length[r]);
		buf_append(&desc_lens, info.desc_length, (size_t) info.nrec * sizeof(int32_t));
		n_total += info.nrec;
		bsgpu_fasta_info_free(&info);
	}
	/* allocate the XStringSet like _alloc_XStringSet() does (src/XStringSet_class.c:78-89) */
	PROTECT(width = NEW_INTEGER(n_total));
	if (n_total)
		memcpy(INTEGER(width), widths.p, (size_t) n_total * sizeof(int));
	if (use_names0) {
		PROTECT(names = NEW_CHARACTER(n_total));
		for (r = 0; r < n_total; r++) {
			int len = ((const int32_t *) desc_lens.p)[r];
			char *tmp = R_alloc((size_t) len + 1, 1);

			memcpy(tmp, descs.p + doff, (size_t) len);
			tmp[len] = 0;
			SET_STRING_ELT(names, r, mkChar(tmp));
			doff += (size_t) len;
		}
		SET_NAMES(width, names);
		UNPROTECT(1);
	}
	snprintf(classname, sizeof(classname), "%sSet", CHAR(STRING_ELT(elementType, 0)));
	PROTECT(ans = alloc_XRawList(classname, CHAR(STRING_ELT(elementType, 0)), width));
	holder = hold_XVectorList(ans);
	for (r = 0; r < n_total; r++) {
		Chars_holder e = get_elt_from_XRawList_holder(&holder, r);

		if (e.length)
			memcpy((char *) e.ptr, letters.p + off, (size_t) e.length);
		off += (size_t) e.length;
	}
	free(text.p); free(letters.p); free(widths.p); free(descs.p); free(desc_lens.p);
	UNPROTECT(2);
	return ans;
}
