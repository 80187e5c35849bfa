// Run-time compiled closures: the reference's field methods take Rust closures
// (`apply_to_all_values(|v| ..., option)`, dense_number_grid_2d.rs:155-195; a model's `State::update`
// rule over a grid).  A closure cannot cross a C ABI, so the caller hands over its BODY as a CUDA C
// expression; it is compiled for sm_100a with NVRTC into a kernel of this library's own shape, loaded
// through the driver API and cached per (device, source).  libnvrtc / libcuda are opened with dlopen
// on first use, so the library itself keeps loading on a machine without them.
#pragma once
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace kg {
namespace jit {

struct Api {
  void* h_nvrtc = nullptr;
  void* h_cuda = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**) = nullptr;
  std::string why;  // empty = usable
};

inline void* open_first(const std::vector<const char*>& names) {
  for (const char* n : names)
    if (void* h = dlopen(n, RTLD_NOW | RTLD_LOCAL)) return h;
  return nullptr;
}

inline Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    a.h_nvrtc = open_first({"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                            "/usr/local/cuda/lib64/libnvrtc.so"});
    a.h_cuda = open_first({"libcuda.so.1", "libcuda.so"});
    if (!a.h_nvrtc) { a.why = "libnvrtc not found (run-time compiled closures need the CUDA toolkit's NVRTC)"; return; }
    if (!a.h_cuda) { a.why = "libcuda not found (no NVIDIA driver)"; return; }
#define KG_SYM(h, field, name)                                              \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));            \
  if (!a.field && a.why.empty()) a.why = std::string("missing symbol ") + name
    KG_SYM(a.h_nvrtc, CreateProgram, "nvrtcCreateProgram");
    KG_SYM(a.h_nvrtc, CompileProgram, "nvrtcCompileProgram");
    KG_SYM(a.h_nvrtc, GetCUBINSize, "nvrtcGetCUBINSize");
    KG_SYM(a.h_nvrtc, GetCUBIN, "nvrtcGetCUBIN");
    KG_SYM(a.h_nvrtc, GetProgramLogSize, "nvrtcGetProgramLogSize");
    KG_SYM(a.h_nvrtc, GetProgramLog, "nvrtcGetProgramLog");
    KG_SYM(a.h_nvrtc, DestroyProgram, "nvrtcDestroyProgram");
    KG_SYM(a.h_cuda, ModuleLoadData, "cuModuleLoadData");
    KG_SYM(a.h_cuda, ModuleGetFunction, "cuModuleGetFunction");
    KG_SYM(a.h_cuda, LaunchKernel, "cuLaunchKernel");
#undef KG_SYM
  });
  return a;
}

// Compile `src` (one extern "C" kernel called `entry`) for sm_100a and return its function handle for
// the CURRENT device (the caller has done cudaSetDevice + touched the runtime).  Cached.
inline int get_kernel(int device, const std::string& src, const char* entry, CUfunction* out) {
  Api& a = api();
  if (!a.why.empty()) return fail(KG_E_CUDA, "%s", a.why.c_str());
  static std::mutex mu;
  static std::map<std::pair<int, std::string>, CUfunction> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(device, src);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return KG_OK;
  }
  nvrtcProgram prog;
  if (a.CreateProgram(&prog, src.c_str(), "kg_closure.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
    return fail(KG_E_CUDA, "nvrtcCreateProgram failed");
  // the closure's f32 arithmetic follows the library's own rules: no contraction, IEEE division and sqrt
  const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--prec-div=true", "--prec-sqrt=true",
                        "--std=c++17"};
  const nvrtcResult rc = a.CompileProgram(prog, 5, opts);
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0;
    a.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) a.GetProgramLog(prog, &log[0]);
    a.DestroyProgram(&prog);
    if (log.size() > 900) log.resize(900);
    return fail(KG_E_INVALID, "the closure does not compile: %s", log.c_str());
  }
  size_t n = 0;
  a.GetCUBINSize(prog, &n);
  std::vector<char> cubin(n);
  a.GetCUBIN(prog, cubin.data());
  a.DestroyProgram(&prog);
  CUmodule mod;
  CUresult cr = a.ModuleLoadData(&mod, cubin.data());
  if (cr != CUDA_SUCCESS) return fail(KG_E_CUDA, "cuModuleLoadData failed (%d)", (int)cr);
  CUfunction fn;
  cr = a.ModuleGetFunction(&fn, mod, entry);
  if (cr != CUDA_SUCCESS) return fail(KG_E_CUDA, "cuModuleGetFunction failed (%d)", (int)cr);
  cache[key] = fn;
  *out = fn;
  return KG_OK;
}

inline int launch(CUfunction fn, unsigned grid, unsigned block, cudaStream_t stream, void** args) {
  const CUresult cr = api().LaunchKernel(fn, grid, 1, 1, block, 1, 1, 0, (CUstream)stream, args, nullptr);
  if (cr != CUDA_SUCCESS) return fail(KG_E_CUDA, "cuLaunchKernel failed (%d)", (int)cr);
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  return KG_OK;
}

}  // namespace jit
}  // namespace kg
