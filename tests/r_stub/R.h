/* Declarations-only stand-in for R's C API (R is not installed in the build image): just enough of
 * R.h / Rinternals.h / R_ext/Rdynload.h, with the signatures of R 4.x, for `gcc -fsyntax-only` to
 * type-check r/epimcmc_shim.c (tests/test_capi.py).  Nothing here is ever linked or run. */
#ifndef R_STUB_R_H
#define R_STUB_R_H
#include <stddef.h>
void Rf_error(const char *, ...) __attribute__((noreturn));
char *R_alloc(size_t, int);
void R_CheckUserInterrupt(void);
#endif
