// Internals shared by the translation units of libpvgpu.so (not part of the C ABI).
#pragma once
#include <string>

namespace pvgpu {
// Sets the calling thread's pvgpu_last_error() message and returns `code`.
int fail(int code, const char *fmt, ...);
// Message of the calling thread (worker threads hand theirs to the thread that called the C ABI).
const std::string &last_error_string();
}  // namespace pvgpu
