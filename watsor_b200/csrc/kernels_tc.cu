// kernels_tc.cu -- placeholder until the tcgen05 GEMM lands (precision 1 is refused loudly).
#include "kernels_tc.cuh"

int tc_prepare_weights(const std::vector<wb_layer>&, const std::vector<wb_tensor_entry>&, const float*, TcWeights*,
                       std::string* err) {
  *err = "the bf16 tcgen05 path is not built into this library";
  return 1;
}
void tc_free_weights(TcWeights*) {}
int tc_launch_gemm(const LaunchCtx&, const TcWeights&, int, int, const wb_layer&, const __nv_bfloat16*, const float*,
                   const float*, __nv_bfloat16*, float*, float*, int, int, std::string* err) {
  *err = "the bf16 tcgen05 path is not built into this library";
  return 1;
}
