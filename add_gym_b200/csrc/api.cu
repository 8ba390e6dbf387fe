// Library-wide state of libaddk.so: last error text, version, launch counter.
#include "common.cuh"
#include "addk.h"
#include <string.h>

static thread_local char g_err[256] = "";
std::atomic<long long> g_addk_launches{0};

void addk_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

extern "C" const char* addk_last_error(void) { return g_err; }
extern "C" int addk_version(void) { return 100; }
extern "C" long long addk_launch_count(int reset) {
  long long v = g_addk_launches.load();
  if (reset) g_addk_launches.store(0);
  return v;
}
