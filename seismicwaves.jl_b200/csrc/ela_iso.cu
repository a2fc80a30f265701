// ela_iso.cu -- 2D elastic isotropic P-SV (placeholder until the kernels land; fails loudly).
#include "common.cuh"
#include "kernels.h"
#include "engine.h"
namespace swb {
void ela_step(const swb_ela_step_args &, bool) { throw Error(SWB_ERR_STATE, "elastic kernels not built yet"); }
void ela_correlate(const swb_ela_correlate_args &) { throw Error(SWB_ERR_STATE, "elastic kernels not built yet"); }
SimBase *make_elastic_iso(const swb_sim_desc &) { throw Error(SWB_ERR_STATE, "elastic engine not built yet"); }
} // namespace swb
