/* Stand-in for Qode's PyC_types.h (the real header lives in the external Qode
 * library, which is not under /root/reference).  Only used to compile the
 * reference's general-XRCC/H_contractions.c, where it lies, into oracle/_ref/
 * as the parity checker.  Test infrastructure, not product code. */
#ifndef XR_ORACLE_PYC_TYPES_H
#define XR_ORACLE_PYC_TYPES_H
#include <stdint.h>
typedef int64_t PyInt;
typedef int64_t BigInt;
typedef double  PyFloat;
typedef double  Double;
#endif
