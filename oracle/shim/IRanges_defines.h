/* minir stand-in for IRanges' C typedefs (TEST INFRASTRUCTURE ONLY). */
#ifndef MINIR_IRANGES_DEFINES_H
#define MINIR_IRANGES_DEFINES_H
#include "S4Vectors_defines.h"

typedef struct iranges_holder {
	const char *classname;
	int is_constant_width;
	int length;
	const int *width;
	const int *start;
	const int *end;
	int SEXP_offset;
	SEXP names;
} IRanges_holder;

#endif
