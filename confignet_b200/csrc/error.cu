#include "common.cuh"
static thread_local char g_err[512] = "";
void cn_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* cn_last_error(void) { return g_err; }
extern "C" int cn_version(void) { return 1; }
unsigned long long g_cn_launches = 0;
extern "C" long long cn_launch_count(int reset) {
  long long v = (long long)g_cn_launches;
  if (reset) g_cn_launches = 0;
  return v;
}
extern "C" long long cn_launch_count_add(long long d) { g_cn_launches += (unsigned long long)d; return (long long)g_cn_launches; }
// Parameter epoch: bumped whenever a registered parameter buffer may have changed (optimizer / EMA kernels,
// set_weights, registration).  conv.cu keeps the tensor-core stage images of registered weights until then.
unsigned long long g_cn_weight_epoch = 1;
extern "C" int cn_weights_changed(void) { ++g_cn_weight_epoch; return CN_OK; }
