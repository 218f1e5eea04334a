// merge_step_spec_hss.cu - the step kernel specialised for all-CAV envs under the HSS shield (see "Other builds" at the
// top of merge_step.cu): BASELINE configs[1].
#include "marl_mass_b200.h"
#define MM_SPEC_SHIELD 1   /* MM_SHIELD_HSS */
#include "merge_step.cu"
