// Error plumbing and device queries shared by all entry points of libsrvp_b200.so.
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {

static std::mutex g_err_mutex;
static char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  std::lock_guard<std::mutex> lock(g_err_mutex);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int num_sms_cached() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

}  // namespace srvp

extern "C" const char* srvp_last_error(void) { return srvp::g_err; }
extern "C" int srvp_version(void) { return 100; }
extern "C" unsigned long long srvp_launch_count(void) { return srvp::g_launches; }
extern "C" int srvp_num_sms(void) { return srvp::num_sms_cached(); }
