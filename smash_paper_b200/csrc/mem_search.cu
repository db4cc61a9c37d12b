// mem_search.cu -- K2: longSA::MEM / findMEM (longSA.cpp:395-490), placeholder until the faithful
// kernel lands; smash_ctx_create refuses MEM mode while launch_mem_search returns -1.
#include "kernels.cuh"
