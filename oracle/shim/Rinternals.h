/* mini-R stand-in: everything lives in Rdefines.h */
#include <Rdefines.h>
