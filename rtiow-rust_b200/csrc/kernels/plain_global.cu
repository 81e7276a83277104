#define RTIOW_PLAIN_SMEM false
#define RTIOW_PLAIN_NAME pick_plain_global
#include "plain_impl.cuh"
