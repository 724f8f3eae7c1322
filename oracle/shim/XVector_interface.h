/* minir stand-in for XVector's C interface (TEST INFRASTRUCTURE ONLY). */
#ifndef MINIR_XVECTOR_INTERFACE_H
#define MINIR_XVECTOR_INTERFACE_H
#include "XVector_defines.h"

Chars_holder hold_XRaw(SEXP x);
XVectorList_holder hold_XVectorList(SEXP x);
int get_length_from_XVectorList_holder(const XVectorList_holder *x_holder);
Chars_holder get_elt_from_XRawList_holder(const XVectorList_holder *x_holder, int i);
XVectorList_holder get_linear_subset_from_XVectorList_holder(
		const XVectorList_holder *x_holder, int offset, int length);
int get_XVectorList_length(SEXP x);
SEXP get_XVectorList_width(SEXP x);
void set_XVectorList_names(SEXP x, SEXP names);
SEXP get_XVector_tag(SEXP x);
SEXP new_XInteger_from_tag(const char *classname, SEXP tag);
SEXP new_XRaw_from_tag(const char *classname, SEXP tag);
SEXP alloc_XRawList(const char *classname, const char *element_type, SEXP width);
SEXP alloc_XRaw(const char *classname, int length);
SEXP get_XVectorList_names(SEXP x);

/* io_utils: an in-memory "file" behind an external pointer (see minir_new_filexp) */
int filexp_gets(SEXP filexp, char *buf, int buf_size, int *EOL_in_buf);
void filexp_seek(SEXP filexp, long long int offset, int whence);
long long int filexp_tell(SEXP filexp);
int filexp_puts(SEXP filexp, const char *s);
int delete_trailing_LF_or_CRLF(const char *buf, int buf_len);

static inline int translate_byte(char byte, const int *lkup, int lkup_length)
{
	int key = (unsigned char) byte;

	return key >= lkup_length ? NA_INTEGER : lkup[key];
}

void Ocopy_bytes_from_i1i2_with_lkup(int i1, int i2, char *dest, int dest_nbytes,
		const char *src, int src_nbytes, const int *lkup, int lkup_length);
void Ocopy_bytes_to_i1i2_with_lkup(int i1, int i2,
		char *dest, int dest_nbytes,
		const char *src, int src_nbytes,
		const int *lkup, int lkup_length);

#endif
