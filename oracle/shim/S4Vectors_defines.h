/* minir stand-in for S4Vectors' C typedefs (TEST INFRASTRUCTURE ONLY). */
#ifndef MINIR_S4VECTORS_DEFINES_H
#define MINIR_S4VECTORS_DEFINES_H
#include <Rdefines.h>

typedef struct int_ae { size_t _buflength; size_t _nelt; int *elts; } IntAE;
typedef struct int_aeae { size_t _buflength; size_t _nelt; IntAE **elts; } IntAEAE;
typedef struct llong_ae { size_t _buflength; size_t _nelt; long long *elts; } LLongAE;
/* strings only (src/read_fasta_files.c keeps the description lines in one) */
typedef struct char_aeae { size_t _buflength; size_t _nelt; char **elts; } CharAEAE;

#endif
