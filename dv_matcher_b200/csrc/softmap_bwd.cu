// placeholder: soft-map backward lands after the forward is parity-green
#include "softmap.cuh"
using namespace dvm;
extern "C" size_t dvm_softmap_bwd_workspace_bytes(int, int, int, int) { return 256; }
extern "C" int dvm_softmap_bwd(const float*, const float*, int, int, int, int, float, int,
                               const int32_t*, const float*, const float*, const float*, const float*, const float*,
                               float*, float*, void*, size_t, void*) {
    set_error("dvm_softmap_bwd: not implemented yet");
    return DVM_ERR_UNSUPPORTED;
}
