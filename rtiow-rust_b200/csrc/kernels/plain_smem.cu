#define RTIOW_PLAIN_SMEM true
#define RTIOW_PLAIN_NAME pick_plain_smem
#include "plain_impl.cuh"
