// td_sc_k4.cu -- instantiations of the persistent solver for 4 wavenumber(s) per thread (split for parallel compilation)
#include "td_sc_impl.cuh"

namespace mct {
void *td_sc_entry_k4(int UR) { return pick_tr<4>(UR); }
}  // namespace mct
