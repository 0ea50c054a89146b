// td_sc_k1.cu -- instantiations of the persistent solver for 1 wavenumber(s) per thread (split for parallel compilation)
#include "td_sc_impl.cuh"

namespace mct {
void *td_sc_entry_k1(int UR) { return pick_tr<1>(UR); }
}  // namespace mct
