/* minir stand-in for IRanges' C interface (TEST INFRASTRUCTURE ONLY). */
#ifndef MINIR_IRANGES_INTERFACE_H
#define MINIR_IRANGES_INTERFACE_H
#include "IRanges_defines.h"

SEXP get_H2LGrouping_high2low(SEXP x);
SEXP get_H2LGrouping_low2high(SEXP x);
SEXP new_IRanges(const char *classname, SEXP start, SEXP width, SEXP names);

#endif
