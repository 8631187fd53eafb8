/* MOCK of <caml/bigarray.h> (see mlvalues.h in this directory). */
#ifndef MOCK_CAML_BIGARRAY_H
#define MOCK_CAML_BIGARRAY_H
#include "mlvalues.h"
#include "custom.h"
struct caml_ba_proxy;
struct caml_ba_array {
  void* data;
  intnat num_dims;
  intnat flags;
  struct caml_ba_proxy* proxy;
  intnat dim[1];
};
#define Caml_ba_array_val(v) ((struct caml_ba_array*)Data_custom_val(v))
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
uintnat caml_ba_byte_size(struct caml_ba_array* b);
#endif
