/* MOCK of <caml/fail.h> (see mlvalues.h in this directory). */
#ifndef MOCK_CAML_FAIL_H
#define MOCK_CAML_FAIL_H
void caml_invalid_argument(const char* msg) __attribute__((noreturn));
void caml_failwith(const char* msg) __attribute__((noreturn));
void caml_raise_out_of_memory(void) __attribute__((noreturn));
#endif
