// Compatibility stub, test infrastructure only (see oracle/README.md).
// The reference's PytorchEMD/cuda/emd_kernel.cu:17 includes <THC/THC.h>, which was
// removed from PyTorch >= 1.11.  This stub supplies the three names that file uses
// so the UNMODIFIED reference source compiles against torch 2.x.
#pragma once
#include <c10/cuda/CUDAException.h>
#include <c10/util/Exception.h>
#ifndef THCudaCheck
#define THCudaCheck(x) C10_CUDA_CHECK(x)
#endif
#ifndef CHECK_EQ
#define CHECK_EQ(a, b) TORCH_CHECK((a) == (b), #a " != " #b)
#endif
