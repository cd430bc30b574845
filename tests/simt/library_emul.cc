// TEST INFRASTRUCTURE - NOT A PRODUCT PATH, never shipped, never loaded by the package.
// The library's own three translation units (kernels, host orchestration + C ABI, decoder) compiled for the host
// against the SIMT emulator and the CUDA runtime stand-ins of this directory, so that the CPU test tier can drive the
// real C ABI - streaming calls, carry buffers, the three-stream orchestration, deferred / speculative resolution of
// time-sharded spans, the split-form blocks, the decoder - and compare it with the oracle where no GPU exists.
// It is 10^4-10^5 times slower than one CPU core running the reference, and it is not a fallback: the package binds
// gr_air_modes_b200/libairmodes_b200.so only (gr_air_modes_b200/_lib.py) and refuses to work without a device.
// tests/test_library_simt.py builds this file into a temporary directory with plain g++.
#include "simt_emul.h"

#include "../../gr_air_modes_b200/csrc/amb_kernels.cu"
#include "../../gr_air_modes_b200/csrc/amb_api.cu"
#include "../../gr_air_modes_b200/csrc/amb_decode.cu"
