// Error state and version of libagcn_b200.so.
#include "common.cuh"
#include <string.h>

namespace agcn {

char* error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace agcn

extern "C" int agcn_version(void) { return 100; }

extern "C" const char* agcn_last_error_string(void) { return agcn::error_buffer(); }
