// td_sc_k8.cu -- instantiations of the persistent solver for 8 wavenumber(s) per thread (split for parallel compilation)
#include "td_sc_impl.cuh"

namespace mct {
void *td_sc_entry_k8(int UR) { return pick_tr<8>(UR); }
}  // namespace mct
