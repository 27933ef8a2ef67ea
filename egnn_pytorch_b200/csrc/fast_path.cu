#include "fast_path.h"
namespace egnn {
int fast_packed_bytes(const EgnnLayerDesc&, size_t*) { return EGNN_ERR_UNSUPPORTED; }
int fast_pack_weights(const EgnnLayerDesc&, const EgnnLayerWeights&, void*, size_t, cudaStream_t) { return EGNN_ERR_UNSUPPORTED; }
int fast_workspace_bytes(const EgnnLayerDesc&, size_t*) { return EGNN_ERR_UNSUPPORTED; }
int fast_forward(const EgnnLayerDesc&, const EgnnLayerWeights&, const void*, const EgnnLayerIO&, void*, size_t, cudaStream_t) { return EGNN_ERR_UNSUPPORTED; }
}
