// constants.cuh -- forwarding header of the GBD-PCG drop-in set; everything lives in gpu_pcg.cuh.
#pragma once
#include "gpu_pcg.cuh"
