// placeholder: fused forward layer (filled in next)
#include "spmm_core.cuh"
extern "C" int ngcf_fused_tile_rows(void) { return 64; }
extern "C" int ngcf_fused_tile_entries(void) { return 2048; }
