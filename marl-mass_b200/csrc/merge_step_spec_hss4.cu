// merge_step_spec_hss4.cu - the step kernel specialised for all-CAV envs under the HSS shield AND built for four CTAs per
// SM (see "Other builds" at the top of merge_step.cu): small grids of the benchmark configurations.
#include "marl_mass_b200.h"
#define MM_VARIANT4 1
#define MM_SPEC_SHIELD 1
#include "merge_step.cu"
