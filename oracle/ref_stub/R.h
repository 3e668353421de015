/* Minimal stand-in for <R.h> so the reference's deSolve-ABI RHS files
 * (/root/reference/R/models/model-age*.cpp, which only use error()) compile
 * without an R installation.  Used ONLY by oracle/Makefile to build
 * oracle/_ref/*.so as a behavioural cross-check of the restated RHS. */
#ifndef ORACLE_R_STUB_H
#define ORACLE_R_STUB_H
#include <stdio.h>
#include <stdlib.h>
#define error(...) do { fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); abort(); } while (0)
#endif
