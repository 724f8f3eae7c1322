/* minir stand-in (test infrastructure only): nothing on the PDict path needs Rdynload. */
#ifndef MINIR_RDYNLOAD_H
#define MINIR_RDYNLOAD_H
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct minir_dllinfo DllInfo;
#endif
