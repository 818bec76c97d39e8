#include "sfem_internal.h"
namespace sfem {
bool canonical_supported(const sfem_pattern_s*) { return false; }
int canonical_launches(const sfem_pattern_s*) { return 0; }
int assemble_canonical_ST(const sfem_pattern_s*, const double*, const double*, const uint8_t*, double*, double*,
                          cudaStream_t) {
  set_error("canonical kernels not built");
  return 2;
}
}  // namespace sfem
