// Build-time check of the run-time specialised filter kernel: compiles filter_comb_e.cuh with
// its default parameters (the BASELINE cfg2 plan) so that errors and the register / shared
// memory footprint (ptxas -v, build/filter_comb_e_check.ptxas.txt) are visible without a GPU.
// The object is not linked into the library; the shipped kernel is built by NVRTC from the
// same source.
#include "filter_comb_e.cuh"
