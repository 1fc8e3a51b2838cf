// placeholder, replaced by the tcgen05 engine
#include "conv_common.cuh"
namespace rcfd {
bool conv_tc_supported(const ConvKP&, int) { return false; }
int conv_tc_launch(const ConvKP&, cudaStream_t) { set_error("tcgen05 engine not built"); return RCFD_EUNSUPPORTED; }
}
