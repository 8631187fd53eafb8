/* MOCK of <caml/alloc.h> (see mlvalues.h in this directory). */
#ifndef MOCK_CAML_ALLOC_H
#define MOCK_CAML_ALLOC_H
#include "mlvalues.h"
value caml_alloc_tuple(uintnat n);
value caml_copy_double(double d);
#endif
