/*
 * minir stand-in for S4Vectors' C interface (TEST INFRASTRUCTURE ONLY).
 * Semantics restated from the call sites in /root/reference/src (S4Vectors is
 * not vendored there): auto-extending int buffers and a few helpers.
 */
#ifndef MINIR_S4VECTORS_INTERFACE_H
#define MINIR_S4VECTORS_INTERFACE_H
#include "S4Vectors_defines.h"


IntAE *new_IntAE(size_t buflength, size_t nelt, int val);
size_t IntAE_get_nelt(const IntAE *ae);
size_t IntAE_set_nelt(IntAE *ae, size_t nelt);
void IntAE_set_val(const IntAE *ae, int val);
void IntAE_insert_at(IntAE *ae, size_t at, int val);
void IntAE_append(IntAE *ae, const int *newvals, size_t nnewval);
void IntAE_shift(const IntAE *ae, size_t offset, int shift);
void IntAE_qsort(const IntAE *ae, size_t offset, int desc);
void IntAE_uniq(IntAE *ae, size_t offset);
SEXP new_INTEGER_from_IntAE(const IntAE *ae);
IntAE *new_IntAE_from_CHARACTER(SEXP x, int keyshift);

LLongAE *new_LLongAE(size_t buflength, size_t nelt, long long val);
size_t LLongAE_get_nelt(const LLongAE *ae);
void LLongAE_insert_at(LLongAE *ae, size_t at, long long val);
CharAEAE *new_CharAEAE(size_t buflength, size_t nelt);
size_t CharAEAE_get_nelt(const CharAEAE *aeae);
void CharAEAE_append_string(CharAEAE *aeae, const char *string);
SEXP new_CHARACTER_from_CharAEAE(const CharAEAE *aeae);
void list_as_data_frame(SEXP x, int nrow);

IntAEAE *new_IntAEAE(size_t buflength, size_t nelt);
size_t IntAEAE_get_nelt(const IntAEAE *aeae);
void IntAEAE_sum_and_shift(const IntAEAE *aeae1, const IntAEAE *aeae2, int shift);
SEXP new_LIST_from_IntAEAE(const IntAEAE *aeae, int mode);
SEXP IntAEAE_toEnvir(const IntAEAE *aeae, SEXP envir, int keyshift);

void sort_int_array(int *x, size_t nelt, int desc);
void get_order_of_int_array(const int *x, int nelt, int desc, int *out, int out_shift);

const char *get_classname(SEXP x);
const char *get_List_elementType(SEXP x);
void set_List_elementType(SEXP x, const char *type);

#endif
