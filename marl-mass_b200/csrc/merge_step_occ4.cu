// merge_step_occ4.cu - the step kernel built for four CTAs per SM (see MM_VARIANT4 at the top of merge_step.cu).
#define MM_VARIANT4 1
#include "merge_step.cu"
