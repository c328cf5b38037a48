/* The drop-in boundary is a C ABI: this translation unit is compiled as strict C99 (not C++) against
 * include/egb200.h and linked with libegb200.so by tests/test_bindings.py. It makes only host-side calls
 * (no GPU needed): version string, last-error string, device enumeration, program parse + shape inference. */
#include <stdio.h>
#include <string.h>

#include "egb200.h"

int main(void) {
  const char* v = egb_version();
  int count = -1;
  int status;
  if (!v || !strstr(v, "sm_100a")) {
    fprintf(stderr, "unexpected version string\n");
    return 1;
  }
  status = egb_device_count(&count);
  if (status != EGB_OK && egb_last_error() == NULL) {
    fprintf(stderr, "failing call left no error message\n");
    return 2;
  }
  {
    /* a malformed program must come back as EGB_ERR_PARSER with a message, not crash */
    egb_program* prog = NULL;
    const char* text = "egbprog 1 f32 0 tensors 0 targets nonsense";
    status = egb_program_parse(text, strlen(text), &prog);
    if (status != EGB_ERR_PARSER || strlen(egb_last_error()) == 0) {
      fprintf(stderr, "parser error convention broken: %d\n", status);
      return 3;
    }
  }
  {
    egb_program* prog = NULL;
    const char* text = "egbprog 1 f32 0 tensors 0 targets 0 end";
    status = egb_program_parse(text, strlen(text), &prog);
    if (status != EGB_OK) {
      fprintf(stderr, "empty program rejected: %s\n", egb_last_error());
      return 4;
    }
    egb_program_free(prog);
  }
  printf("ok %s devices=%d\n", v, count);
  return 0;
}
