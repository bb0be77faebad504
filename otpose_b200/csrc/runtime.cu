// Library-level entry points: version, last-error string, device probe.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace otp {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace otp

namespace otp {
namespace {
constexpr int kMaxRecords = 1 << 19;   // a >= 2 s roofline pass records ~170 scopes per step
struct Record {
  int id;
  cudaEvent_t a, b;
};
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_profile_on{0};
std::mutex g_mu;
std::vector<Record> g_records;       // used records of the current session
std::vector<Record> g_pool;          // recycled event pairs
const char *kNames[K_COUNT] = {"final_preds", "mdcn_fwd", "fusion_sum", "fusion_stack", "add_pos_embd",
                               "upsample_linear", "pyramid_conv1x1", "conv2d", "block_front", "block_fold",
                               "block_apply", "block_back", "pack", "tc_block_front", "tc_block_apply",
                               "tc_block_back", "tc_offset_mask_dcn", "mdcn_bwd", "final_layer_sum", "conv2d_bwd", "flow_encoder", "rsb_block", "window_assemble"};
}  // namespace

LaunchScope::LaunchScope(int id, cudaStream_t st, int nlaunch) : slot_(-1), st_(st) {
  g_launches.fetch_add((unsigned long long)nlaunch, std::memory_order_relaxed);
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if ((int)g_records.size() >= kMaxRecords) return;
  Record r;
  r.id = id;
  if (!g_pool.empty()) {
    r = g_pool.back();
    r.id = id;
    g_pool.pop_back();
  } else if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
    return;
  }
  cudaEventRecord(r.a, st);
  g_records.push_back(r);
  slot_ = (int)g_records.size() - 1;
}

LaunchScope::~LaunchScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot_ < (int)g_records.size()) cudaEventRecord(g_records[slot_].b, st_);
}
}  // namespace otp

extern "C" unsigned long long otp_launch_count(void) { return otp::g_launches.load(); }

extern "C" int otp_profile_num_kernels(void) { return otp::K_COUNT; }

extern "C" const char *otp_profile_kernel_name(int id) {
  return (id >= 0 && id < otp::K_COUNT) ? otp::kNames[id] : "";
}

extern "C" int otp_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(otp::g_mu);
  for (auto &r : otp::g_records) otp::g_pool.push_back(r);
  otp::g_records.clear();
  otp::g_profile_on.store(on ? 1 : 0);
  return OTP_OK;
}

extern "C" int otp_profile_read(float *total_ms, int *launches, int n) {
  OTP_REQUIRE(total_ms != nullptr && launches != nullptr && n >= otp::K_COUNT);
  std::lock_guard<std::mutex> lk(otp::g_mu);
  for (int i = 0; i < n; ++i) {
    total_ms[i] = 0.f;
    launches[i] = 0;
  }
  for (auto &r : otp::g_records) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) {
      otp::set_error("otp_profile_read: %s", cudaGetErrorString(cudaGetLastError()));
      return OTP_ERR_CUDA;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    total_ms[r.id] += ms;
    launches[r.id] += 1;
  }
  return OTP_OK;
}

extern "C" const char *otp_version(void) { return "otpose_b200 0.1 (sm_100a)"; }

extern "C" const char *otp_last_error(void) { return otp::g_err; }

extern "C" int otp_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    otp::set_error("otp_device_is_sm100: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  return major == 10 ? 1 : 0;
}
