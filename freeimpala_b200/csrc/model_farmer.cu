// FarmerLstm learner step (placeholder while the LSTM kernels land).
#include "learner.cuh"
namespace fi {
int farmer_alloc(fi_learner*, Player*) { return set_error(FI_ERR_STATE, "FarmerLstm step not available in this build"); }
void farmer_free(Player*) {}
int farmer_forward_backward(fi_learner*, Player*, const float*, int, int, int) { return set_error(FI_ERR_STATE, "FarmerLstm step not available in this build"); }
int farmer_infer_alloc(fi_learner*, Player*, size_t, size_t) { return set_error(FI_ERR_STATE, "FarmerLstm step not available in this build"); }
int farmer_infer(fi_learner*, Player*, const float*, const float*, const float*, size_t, size_t, float*, cudaStream_t) { return set_error(FI_ERR_STATE, "FarmerLstm step not available in this build"); }
}  // namespace fi
