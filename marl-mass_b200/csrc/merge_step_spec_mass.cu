// merge_step_spec_mass.cu - the step kernel specialised for all-CAV envs under the MASS shield (see "Other builds" at the
// top of merge_step.cu): BASELINE configs[2..4].
#include "marl_mass_b200.h"
#define MM_SPEC_SHIELD 2   /* MM_SHIELD_MASS */
#include "merge_step.cu"
