// CPU stand-in for <cuda.h> (see tcgen05_model.h)
#pragma once
#include "tcgen05_model.h"
