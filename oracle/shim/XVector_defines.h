/* minir stand-in for XVector's C typedefs (TEST INFRASTRUCTURE ONLY). */
#ifndef MINIR_XVECTOR_DEFINES_H
#define MINIR_XVECTOR_DEFINES_H
#include <Rdefines.h>

typedef struct chars_holder {
	const char *ptr;
	int length;
} Chars_holder;

/* In the shim a sequence set is (n, per-element pointer, per-element width). */
typedef struct xvectorlist_holder {
	const char *classname;
	int length;
	const char *const *ptrs;
	const int *widths;
} XVectorList_holder;

#endif
