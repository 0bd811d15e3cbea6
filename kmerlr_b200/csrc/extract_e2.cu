// extract_kernel<2>: one instantiation per translation unit (parallel builds)
#define KL_EXTRACT_KERNEL_IMPL
#include "extract_kernel.cuh"

namespace kl {
namespace xk {
template void launch_extract<2>(const XParams &P);
}
}  // namespace kl
