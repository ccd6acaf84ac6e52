// Translation unit that instantiates the kernels of k1_tc_wide.cuh and defines its launch function; split out of net.cu so that
// the library's kernel families compile in parallel (make -j).
#define BANN_K1_TCW_IMPL 1
#include "k1_tc_wide.cuh"
