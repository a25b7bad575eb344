// kge_umma.cu -- tcgen05 (5th-gen tensor core) engine for the bilinear contractions.
// Placeholder until the 3xTF32 UMMA kernels land: reports "not supported" so the fp32 tile
// engine (kge_tiles.cu) runs.
#include "kge_common.cuh"
namespace kge {
bool umma_supported(const StepParams&) { return false; }
int umma_score(const LaunchCtx&, const StepParams&, const StepWs&, char*, size_t) { return KGE_ERR_UNSUPPORTED; }
int umma_grad(const LaunchCtx&, const StepParams&, const StepWs&, bool, char*, size_t) { return KGE_ERR_UNSUPPORTED; }
}  // namespace kge
