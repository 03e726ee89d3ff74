// C-ABI glue: error reporting, version, device query.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void scnb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* scnb_last_error(void) { return g_err; }

extern "C" int scnb_version(void) { return 100; }

// Returns the compute capability (major*10+minor) of the current device, or a negative error.
extern "C" int scnb_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { scnb_set_error("device_arch: no CUDA device"); return SCNB_ERR_CUDA; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

// Programmatic dependent launch switch (common.cuh): SCNB_PDL=1 enables it.  Measured on B200 inside a captured
// graph it changes the step time by < 1 % (graph launches already hide the launch latency), so it is off by default.
int scnb_pdl_enabled(void) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SCNB_PDL");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on;
}
