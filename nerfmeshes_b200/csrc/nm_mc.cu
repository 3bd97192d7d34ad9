// placeholder until the marching-cubes kernels land (replaced in a later commit)
#include "nm_common.h"
namespace nm {
int mc_count(const float*, int, int, int, float, void**, size_t*, int64_t*, cudaStream_t, int64_t*) {
  set_error("marching cubes not built yet");
  return -1;
}
int mc_emit(const float*, int, int, int, float, float, void*, float*, float*, int32_t*, cudaStream_t, int64_t*) {
  set_error("marching cubes not built yet");
  return -1;
}
}  // namespace nm
