// Bindings for the sm_100a attention kernels (attention_sm100.cu).
#include <torch/extension.h>
void register_attention(pybind11::module_& m) { (void)m; }
