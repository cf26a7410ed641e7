// Shared-memory TT finite-difference kernels, one chain per stencil point (see pcb_tt_chain.inc).
#define PCB_TT_CHAIN_PART 2
#include "pcb_tt_chain.inc"
