// api_common.cu - error text, version, device probe, NCCL bootstrap id (see include/b2nav.h).
#include <cstring>

#include <nccl.h>

#include "common.cuh"

namespace b2n
{
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
} // namespace b2n

extern "C" {

const char *b2n_last_error(void) { return b2n::g_err; }

int b2n_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int b2n_version(char *buf, size_t cap)
{
  static const char v[] = "libb2nav 0.1 sm_100a";
  if (buf && cap) {
    std::strncpy(buf, v, cap - 1);
    buf[cap - 1] = '\0';
  }
  return (int)sizeof(v);
}

int b2n_comm_unique_id(void *out128)
{
  B2N_REQUIRE(out128, B2N_ERR_INVALID_ARGUMENT, "null argument");
  ncclUniqueId id;
  ncclResult_t r = ncclGetUniqueId(&id);
  B2N_REQUIRE(r == ncclSuccess, B2N_ERR_COMM, "ncclGetUniqueId: %s", ncclGetErrorString(r));
  std::memcpy(out128, &id, sizeof(id));
  return B2N_OK;
}

} // extern "C"
