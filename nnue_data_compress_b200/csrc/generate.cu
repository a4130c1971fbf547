// generate.cu -- placeholder until the device-side generator lands
#include "../../include/nnuepack.h"
extern "C" {
int nnp_generate_bin_dev(void*, size_t, uint32_t, uint64_t) { return NNP_ERR_BAD_ARG; }
}
