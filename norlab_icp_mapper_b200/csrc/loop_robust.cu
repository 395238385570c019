// The persistent loop kernel once more, with RobustOutlierFilter compiled in (its in-kernel scale selects, loop_exact_median /
// loop_window_median, and the M-estimator weights of the classify and finish phases): launch_icp_loop_robust.  Kept out of the
// default instantiations because the extra live state costs the headline kernel (Trimmed chains) 2 % in spills; see loop.cu.
#define B200ICP_LOOP_ROBUST 1
#include "loop.cu"
