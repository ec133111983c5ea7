// Multi-GPU merge.  Placeholder: the single-GPU path is brought up first.
#include "bwtm_merge.cuh"

using namespace bwtm;

struct bwtm_comm { int rank, world; };

extern "C"
{

int bwtm_comm_unique_id(uint8_t*) { set_error("multi-GPU merge is not built yet"); return BWTM_ERR_COMM; }
int bwtm_comm_create(const uint8_t*, int, int, bwtm_comm**) { set_error("multi-GPU merge is not built yet"); return BWTM_ERR_COMM; }
int bwtm_comm_destroy(bwtm_comm*) { return BWTM_OK; }
int bwtm_merge_distributed(bwtm_comm*, bwtm_index*, bwtm_index*, const bwtm_merge_options*, bwtm_index**, bwtm_timings*)
{
  set_error("multi-GPU merge is not built yet"); return BWTM_ERR_COMM;
}

} // extern "C"
