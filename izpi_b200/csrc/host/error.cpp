// error.cpp -- thread-local error string behind izpi_last_error() (include/izpi_cuda.h).
#include <string>

#include "../../../include/izpi_cuda.h"

namespace izpi {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace izpi

extern "C" const char* izpi_last_error(void) { return izpi::g_last_error.c_str(); }
extern "C" int izpi_version(void) { return 100; }
