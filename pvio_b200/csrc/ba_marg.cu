// placeholder; replaced below by the real marginaliser
#include "api_internal.h"
namespace pvio {
int marginalize_impl(Handle *h, const pvio_b200_window *, const pvio_b200_state *, int, double *, double *, double *, double *) {
    return fail(h, PVIO_B200_EINVAL, "marginalize: not built yet");
}
}
