// Forced-include shim used ONLY to compile the UNMODIFIED reference `_ext`
// sources (read in place from /root/reference) against torch 2.11 / CUDA 12.9.
// It adds the includes the old sources got transitively and the CHECK_EQ macro
// that newer c10 no longer exports.  No reference code is copied.
#pragma once
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#ifndef CHECK_EQ
#define CHECK_EQ(a, b) TORCH_CHECK((a) == (b))
#endif
