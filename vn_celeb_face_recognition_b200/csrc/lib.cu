// Library-level state of libvnfr_b200.so: error string, version, launch counter.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";
long long g_vnfr_launches = 0;

void vnfr_set_error(const char* file, int line, const char* msg) {
  const char* base = strrchr(file, '/');
  snprintf(g_err, sizeof(g_err), "%s:%d: %s", base ? base + 1 : file, line, msg);
}

extern "C" const char* vnfr_last_error(void) { return g_err; }
extern "C" int vnfr_version(void) { return 100; }
extern "C" long long vnfr_launch_count(void) { return g_vnfr_launches; }
// Launches replayed through a CUDA graph captured from vnfr_run_ops are not seen by the per-call counter: the host
// adds them here so that vnfr_launch_count() keeps counting kernels actually launched.
extern "C" int vnfr_count_launches(long long n) { g_vnfr_launches += n; return VNFR_OK; }
