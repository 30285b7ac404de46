// csrc/pt_error.h -- thread-local error text shared by the host and CUDA translation units of libptgpu.so
#pragma once
#include <cstdarg>
#include <cstdio>
namespace ptgpu {
char* error_buffer();  // 512 bytes, thread-local
inline int set_error(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(error_buffer(), 512, fmt, ap); va_end(ap);
  return code;
}
}  // namespace ptgpu
