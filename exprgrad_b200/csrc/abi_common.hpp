// Shared by the C-ABI translation units: opaque handle layouts and the exception -> status bridge.
#pragma once
#include <string>

#include "../../include/egb200.h"
#include "egb_internal.hpp"

namespace egb {
void set_last_error(const std::string& s);
}

struct egb_context {
  egb::Context c;
};
struct egb_buffer {
  egb_context* ctx;
  size_t size;
  void* ptr;
};

#define EGB_TRY try {
#define EGB_CATCH                                  \
  }                                                \
  catch (const egb::Error& e) {                    \
    egb::set_last_error(e.what());                 \
    return e.code;                                 \
  }                                                \
  catch (const std::exception& e) {                \
    egb::set_last_error(e.what());                 \
    return EGB_ERR_RUNTIME;                        \
  }                                                \
  return EGB_OK;
