// Forced-include shim (nvcc -include) for building the reference's rasterize_cuda_kernel.cu, written for torch 1.6, against
// torch 2.x WITHOUT touching its source: the dispatch macros there are given `tensor.type()`, which no longer converts to a
// ScalarType.  Every header the source includes is pulled in here first (include guards make the source's own #includes
// no-ops), then `type` is renamed for the source's own text only.
#include <ATen/ATen.h>
#include <iostream>
#include <cuda.h>
#include <cuda_runtime.h>
#define type scalar_type
