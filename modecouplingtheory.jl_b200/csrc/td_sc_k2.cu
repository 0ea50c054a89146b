// td_sc_k2.cu -- instantiations of the persistent solver for 2 wavenumber(s) per thread (split for parallel compilation)
#include "td_sc_impl.cuh"

namespace mct {
void *td_sc_entry_k2(int UR) { return pick_tr<2>(UR); }
}  // namespace mct
