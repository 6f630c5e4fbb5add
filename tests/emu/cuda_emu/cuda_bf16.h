// CPU stand-in for <cuda_bf16.h> (see cuda_emu.h)
#pragma once
#include "cuda_emu.h"
