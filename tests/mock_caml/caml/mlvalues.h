/* MOCK of the OCaml runtime's <caml/mlvalues.h> — written for tests/test_ocaml_stubs_typecheck.py only.
 * The image has no OCaml toolchain; these few declarations (names and shapes as documented in the
 * OCaml manual, "Interfacing C with OCaml") let gcc type-check the stub file's calls into
 * include/hnsw_b200.h.  They are not the real headers and nothing links against them. */
#ifndef MOCK_CAML_MLVALUES_H
#define MOCK_CAML_MLVALUES_H
#include <stdint.h>
#include <stddef.h>
typedef intptr_t intnat;
typedef uintptr_t uintnat;
typedef intnat value;
#define CAMLprim
#define CAMLextern extern
#define Long_val(v) ((intnat)(v) >> 1)
#define Int_val(v) ((int)Long_val(v))
#define Val_long(x) ((value)(((uintnat)(intnat)(x) << 1) + 1))
#define Val_int(x) Val_long(x)
#define Val_unit Val_long(0)
#define Field(v, i) (((value*)(v))[i])
void caml_modify(value* fp, value v);
#define Store_field(b, i, v) caml_modify(&Field((b), (i)), (v))
#define Wosize_val(v) ((uintnat)(((uintnat*)(v))[-1] >> 10))
#endif
