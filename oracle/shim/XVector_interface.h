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
void Ocopy_bytes_to_i1i2_with_lkup(int i1, int i2,
		char *dest, int dest_nbytes,
		const char *src, int src_nbytes,
		const int *lkup, int lkup_length);

#endif
