// bf16 tcgen05 path -- placeholder until the sm_100a kernels land (create() with BCAD_PREC_BF16 is refused).
#include "common.cuh"
#include "model.h"

namespace bcad {

int tensor_path_supported(const Model&) {
    set_error("precision=BF16: the tcgen05 path is not built into this libbcad yet; use BCAD_PREC_FP32");
    return BCAD_ERR_INVALID;
}
int tensor_path_commit(Model&) { return BCAD_ERR_STATE; }
int tensor_forward_chunk(Model&, const float*, int, cudaStream_t) { return BCAD_ERR_STATE; }
int tensor_explain_chunk(Model&, int, const int32_t*, int, float*, cudaStream_t) { return BCAD_ERR_STATE; }
int tensor_get_activation(Model&, int, int, int, float*, cudaStream_t) { return BCAD_ERR_STATE; }
void tensor_path_destroy(Model&) {}

}  // namespace bcad
