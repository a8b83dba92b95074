// Library-wide state of libshineon_b200.so: error text, launch counter, version.
#include "common.cuh"

namespace shineon {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};
thread_local cudaError_t g_launch_error = cudaSuccess;
}  // namespace shineon

extern "C" int shineon_version(void) { return 100; }
extern "C" const char* shineon_last_error(void) { return shineon::g_last_error.c_str(); }
extern "C" uint64_t shineon_launch_count(void) { return shineon::g_launch_count.load(); }
