// Bindings for the peer-memory collective kernels (comm.cu).
#include <torch/extension.h>
void register_comm(pybind11::module_& m) { (void)m; }
