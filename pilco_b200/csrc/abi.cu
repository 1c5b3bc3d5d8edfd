// abi.cu -- version / status strings of the C ABI.
#include "common.cuh"

extern "C" {

int pilco_version(void) { return PILCO_ABI_VERSION; }

const char* pilco_status_string(int status) {
    switch (status) {
        case PILCO_OK: return "ok";
        case PILCO_ERR_NULL: return "null pointer argument";
        case PILCO_ERR_DIM: return "invalid dimension (limits: D<=16, E<=16, ldk>=pad64(n))";
        case PILCO_ERR_WORKSPACE: return "workspace too small";
        case PILCO_ERR_ALIGN: return "pointer not 16-byte aligned";
        case PILCO_ERR_LAUNCH: return "CUDA launch failed";
        case PILCO_ERR_UNSUPPORTED: return "unsupported policy/reward kind";
        default: return "unknown status";
    }
}

}  // extern "C"
