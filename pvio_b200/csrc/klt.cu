// placeholder; replaced below by the real KLT kernels
#include "api_internal.h"
namespace pvio {
int klt_track_impl(Handle *h, const uint8_t *, const uint8_t *, int, int, int, const float *, float *, uint8_t *, float *, int, int, int, double) {
    return fail(h, PVIO_B200_EINVAL, "klt: not built yet");
}
void klt_free(Handle *) {}
}
