// plain.cu -- placeholder until the .plain kernels land (next commit)
#include "../../include/nnuepack.h"
extern "C" {
int nnp_plain_to_binpack(const void*, size_t, void*, size_t, size_t*) { return NNP_ERR_BAD_ARG; }
int nnp_binpack_to_plain(const void*, size_t, void*, size_t, size_t*) { return NNP_ERR_BAD_ARG; }
int nnp_bin_to_plain(const void*, size_t, void*, size_t, size_t*) { return NNP_ERR_BAD_ARG; }
int nnp_plain_to_bin(const void*, size_t, void*, size_t, size_t*) { return NNP_ERR_BAD_ARG; }
int nnp_plain_to_binpack_dev(const void*, size_t, void*, size_t, size_t*) { return NNP_ERR_BAD_ARG; }
int nnp_binpack_to_plain_dev(const void*, size_t, void*, size_t, size_t*) { return NNP_ERR_BAD_ARG; }
int nnp_bin_to_plain_dev(const void*, size_t, void*, size_t, size_t*) { return NNP_ERR_BAD_ARG; }
int nnp_plain_to_bin_dev(const void*, size_t, void*, size_t, size_t*) { return NNP_ERR_BAD_ARG; }
}
