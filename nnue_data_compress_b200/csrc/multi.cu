// multi.cu -- the two headline drivers over all GPUs of the process (placeholder, filled in below)
#include "../../include/nnuepack.h"
