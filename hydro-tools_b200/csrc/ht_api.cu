// ht_api.cu -- version / error strings of the C-ABI (include/hydrotools_b200.h).
#include "ht_common.cuh"

extern "C" int ht_version(void) { return 100; }

extern "C" const char *ht_error_string(int code)
{
    switch (code) {
    case HT_OK: return "ok";
    case HT_ERR_ARG: return "invalid argument";
    case HT_ERR_ALIGN: return "pointer or pitch alignment (16 B base, pitch % 16 B) not met";
    case HT_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case HT_ERR_WORKSPACE: return "workspace too small";
    default: break;
    }
    if (code > 0)
        return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}
