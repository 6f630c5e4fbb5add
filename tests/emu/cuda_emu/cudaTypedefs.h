// CPU stand-in for <cudaTypedefs.h> (see tcgen05_model.h)
#pragma once
#include "tcgen05_model.h"
