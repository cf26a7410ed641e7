// Shared-memory TT value kernels (see pcb_tt_chain.inc).
#define PCB_TT_CHAIN_PART 1
#include "pcb_tt_chain.inc"
