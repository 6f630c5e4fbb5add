// CPU stand-in for <cuda_runtime.h> (see cuda_emu.h)
#pragma once
#include "cuda_emu.h"
