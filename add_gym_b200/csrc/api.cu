// Library-wide state of libaddk.so: last error text, version, launch counter.
#include "common.cuh"
#include "addk.h"
#include <string.h>
#include <stdlib.h>
#include "switches.h"

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

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
const AddkSwitches& addk_switches() {
  static const AddkSwitches sw = [] {
    AddkSwitches s;
    s.tc_pair = env_int("ADDK_TC_PAIR", 1);
    s.tc_pair_flags = env_int("ADDK_TC_PAIR_FLAGS", 2);
    s.h3_persistent = env_int("ADDK_H3_PERSISTENT", 1);
    s.h3_pair = env_int("ADDK_H3_PAIR", 1);
    s.h3_chunk_kb = env_int("ADDK_H3_CHUNK_KB", 8);
    if (s.h3_chunk_kb < 1) s.h3_chunk_kb = 1;
    const char* e = getenv("ADDK_H3_COMP");
    s.h3_comp = e ? (float)atof(e) : 1.7e-8f;
    s.bf16_persistent = env_int("ADDK_BF16_PERSISTENT", 1);
    s.bf16_drop_f32 = env_int("ADDK_BF16_DROP_F32", 1);
    s.h3_planes_only = env_int("ADDK_H3_PLANES_ONLY", 1);
    s.h3_amax_hooks = env_int("ADDK_H3_AMAX_HOOKS", 1);
    s.h3_fused_planes = env_int("ADDK_H3_FUSED_PLANES", 0);
    s.h3_colpart = env_int("ADDK_H3_COLPART", 1);
    s.h3_relu_bits = env_int("ADDK_H3_RELU_BITS", 1);
    s.fused_tail = env_int("ADDK_FUSED_TAIL", 1);
    return s;
  }();
  return sw;
}

// ---- debug / test exports -------------------------------------------------------------------------------------------
int g_addk_last_gemm_kernel = 0;
long long* g_addk_stamps = nullptr;
extern "C" int addk_debug_last_gemm_kernel(void) { return g_addk_last_gemm_kernel; }
extern "C" int addk_debug_set_stamp_buffer(long long* device_buffer16) { g_addk_stamps = device_buffer16; return ADDK_OK; }
extern "C" int addk_build_flags(void) {
  int f = 0;
#ifdef ADDK_LEGACY_KERNELS
  f |= 1;
#endif
  return f;
}
