// Handle layouts of the program/model C-ABI group.
#pragma once
#include "abi_common.hpp"
#include "runtime.hpp"

struct egb_program {
  std::shared_ptr<egb::Program> p;
};
struct egb_model {
  std::unique_ptr<egb::Model> m;
  egb_context* ctx;
};
