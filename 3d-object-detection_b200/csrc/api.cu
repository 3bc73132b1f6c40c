// api.cu -- version / error plumbing of libpp_b200.so.
#include "common.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace pp {
thread_local int g_last_cuda_error = 0;
int g_opt_pfn_tensor_cores = 1;
int g_opt_pfn_tc_timing = 0;
int g_opt_pad_reserve_sms = 0;  // SMs the persistent padding pass of pp_input_path leaves to the other stream lane
int g_opt_encode_bulk = 1;       // 1: zero stream of the dense targets as cp.async.bulk copies, 0: float4 store loop
int g_opt_loss_tma = 1;          // 0: generic tile kernel for the classification loss (testing)
int read_tc_prof(long long* out64);
int g_opt_pfn_tc_debug = 0;   // development knob: bit0 skip conversion, bit1 skip MMA, bit2 skip epilogue reads

// ---- launch counter + optional per-kernel CUDA-event timing -----------------------------------
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static thread_local cudaEvent_t t_pending = nullptr;

void prof_begin(const char* name, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  t_pending = nullptr;
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r{name, nullptr, nullptr};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  t_pending = r.b;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}

void prof_end(cudaStream_t st) {
  if (t_pending != nullptr) cudaEventRecord(t_pending, st);
  t_pending = nullptr;
}
}  // namespace pp

extern "C" {

int pp_version(void) { return PP_B200_VERSION; }

int pp_last_cuda_error(void) { return pp::g_last_cuda_error; }

int pp_set_option(const char* key, int value) {
  if (key == nullptr) return PP_ERR_INVALID_ARG;
  if (strcmp(key, "pfn_tensor_cores") == 0) { pp::g_opt_pfn_tensor_cores = value; return PP_OK; }   // 0 CUDA cores, 1 fp16 split (+TF32 fallback), 2 TF32 split
  if (strcmp(key, "loss_tma") == 0) { pp::g_opt_loss_tma = value; return PP_OK; }
  if (strcmp(key, "encode_bulk") == 0) { pp::g_opt_encode_bulk = value; return PP_OK; }
  return PP_ERR_INVALID_ARG;
}

#ifdef PP_DEBUG
/* Development entry points (include/pp_b200_debug.h): only in a library built with -DPP_DEBUG (build.py --debug). */
int pp_debug_set(const char* key, int value) {
  if (key == nullptr) return PP_ERR_INVALID_ARG;
  if (strcmp(key, "pad_reserve_sms") == 0) { pp::g_opt_pad_reserve_sms = value < 0 ? 0 : value; return PP_OK; }
  if (strcmp(key, "pfn_tc_timing") == 0) { pp::g_opt_pfn_tc_timing = value; return PP_OK; }
  if (strcmp(key, "pfn_tc_debug") == 0) { pp::g_opt_pfn_tc_debug = value; return PP_OK; }
  return PP_ERR_INVALID_ARG;
}

/* per-role wait cycles of CTA 0 of the last tensor-core PFN launch (128 int64) */
int pp_debug_tc_timing(int64_t* out64) { return out64 ? pp::read_tc_prof((long long*)out64) : PP_ERR_INVALID_ARG; }
#endif

int64_t pp_launch_count(void) { return (int64_t)pp::g_launches.load(); }

int pp_profile_enable(int on) {
  pp::g_prof_on.store(on ? 1 : 0);
  return PP_OK;
}

// Synchronises the device, then writes "name launches total_ms\n" per kernel (launch order of first
// appearance) into buf and clears the records.  Returns the number of bytes needed (like snprintf).
int64_t pp_profile_report(char* buf, int64_t buf_bytes) {
  using namespace pp;
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::vector<std::string> order;
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (r.a && r.b && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto it = agg.find(r.name);
      if (it == agg.end()) { order.push_back(r.name); agg[r.name] = {1, (double)ms}; }
      else { it->second.first += 1; it->second.second += ms; }
    }
    if (r.a) cudaEventDestroy(r.a);
    if (r.b) cudaEventDestroy(r.b);
  }
  g_prof.clear();
  std::string out;
  char line[256];
  for (auto& n : order) {
    snprintf(line, sizeof line, "%s %lld %.6f\n", n.c_str(), agg[n].first, agg[n].second);
    out += line;
  }
  if (buf != nullptr && buf_bytes > 0) {
    const size_t k = out.size() < (size_t)(buf_bytes - 1) ? out.size() : (size_t)(buf_bytes - 1);
    memcpy(buf, out.data(), k);
    buf[k] = 0;
  }
  return (int64_t)out.size() + 1;
}

const char* pp_error_string(int code) {
  switch (code) {
    case PP_OK: return "ok";
    case PP_ERR_INVALID_ARG: return "invalid argument";
    case PP_ERR_WORKSPACE: return "workspace too small or misaligned";
    case PP_ERR_CUDA: return "CUDA runtime error (see pp_last_cuda_error)";
    case PP_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

}  // extern "C"
