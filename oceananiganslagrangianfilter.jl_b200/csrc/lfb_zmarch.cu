// lfb_zmarch.cu -- placeholder until the z-marching stage kernel lands.
#include "lfb_internal.h"

extern "C" int lfb_zmarch_supported(const LfbStageParams *) { return 0; }
extern "C" cudaError_t lfb_launch_stage_zmarch(const LfbStageParams *, cudaStream_t) { return cudaErrorNotSupported; }
