/* MOCK of <caml/memory.h> (see mlvalues.h in this directory). */
#ifndef MOCK_CAML_MEMORY_H
#define MOCK_CAML_MEMORY_H
#include "mlvalues.h"
#define CAMLparam1(a) value* caml__roots1[] = {&(a)}; (void)caml__roots1
#define CAMLparam2(a, b) value* caml__roots2[] = {&(a), &(b)}; (void)caml__roots2
#define CAMLparam3(a, b, c) value* caml__roots3[] = {&(a), &(b), &(c)}; (void)caml__roots3
#define CAMLparam4(a, b, c, d) value* caml__roots4[] = {&(a), &(b), &(c), &(d)}; (void)caml__roots4
#define CAMLparam5(a, b, c, d, e) value* caml__roots5[] = {&(a), &(b), &(c), &(d), &(e)}; (void)caml__roots5
#define CAMLxparam1(a) value* caml__xroots1[] = {&(a)}; (void)caml__xroots1
#define CAMLlocal1(a) value a = Val_unit
#define CAMLlocal2(a, b) value a = Val_unit, b = Val_unit
#define CAMLreturn(x) return (x)
#endif
