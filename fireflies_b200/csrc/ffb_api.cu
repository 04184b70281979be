// Library-level entry points of libffb200.so.
#include "ffb_common.cuh"

namespace ffb {
char* last_error_buf() {
    static thread_local char buf[256] = "";
    return buf;
}
}  // namespace ffb

extern "C" int ffb_version(void) { return FFB_VERSION; }
extern "C" const char* ffb_last_error_string(void) { return ffb::last_error_buf(); }
