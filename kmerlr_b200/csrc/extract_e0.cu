// extract_kernel<0>: one instantiation per translation unit (parallel builds)
#define KL_EXTRACT_KERNEL_IMPL
#include "extract_kernel.cuh"

namespace kl {
namespace xk {
template void launch_extract<0>(const XParams &P);
}
}  // namespace kl
