// presto_b200 — run-time compilation of generated kernels (see jit_runtime.h).
#include "common.cuh"
#include "jit_runtime.h"
#include <nvrtc.h>
#include <dlfcn.h>
#include <map>
#include <mutex>

namespace pg {

DriverApi& DriverApi::get()
{
    static DriverApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.moduleLoadDataEx = (decltype(api.moduleLoadDataEx)) dlsym(h, "cuModuleLoadDataEx");
        api.moduleLoadData = (decltype(api.moduleLoadData)) dlsym(h, "cuModuleLoadData");
        api.moduleGetFunction = (decltype(api.moduleGetFunction)) dlsym(h, "cuModuleGetFunction");
        api.moduleUnload = (decltype(api.moduleUnload)) dlsym(h, "cuModuleUnload");
        api.launchKernel = (decltype(api.launchKernel)) dlsym(h, "cuLaunchKernel");
        api.funcSetAttribute = (decltype(api.funcSetAttribute)) dlsym(h, "cuFuncSetAttribute");
        api.funcGetAttribute = (decltype(api.funcGetAttribute)) dlsym(h, "cuFuncGetAttribute");
        api.occupancyMaxActiveBlocks = (decltype(api.occupancyMaxActiveBlocks)) dlsym(h, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
        api.ok = api.moduleLoadDataEx && api.moduleLoadData && api.moduleGetFunction && api.moduleUnload && api.launchKernel && api.funcSetAttribute
                 && api.funcGetAttribute && api.occupancyMaxActiveBlocks;
    });
    return api;
}

namespace {

struct NvrtcApi {
    nvrtcResult (*createProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*compileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*getCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*getCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*getProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*getProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*destroyProgram)(nvrtcProgram*) = nullptr;
    bool ok = false;
    static NvrtcApi& get()
    {
        static NvrtcApi api;
        static std::once_flag once;
        std::call_once(once, [] {
            void* h = nullptr;
            for (const char* name : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"}) { h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
            if (!h) return;
            api.createProgram = (decltype(api.createProgram)) dlsym(h, "nvrtcCreateProgram");
            api.compileProgram = (decltype(api.compileProgram)) dlsym(h, "nvrtcCompileProgram");
            api.getCUBINSize = (decltype(api.getCUBINSize)) dlsym(h, "nvrtcGetCUBINSize");
            api.getCUBIN = (decltype(api.getCUBIN)) dlsym(h, "nvrtcGetCUBIN");
            api.getProgramLogSize = (decltype(api.getProgramLogSize)) dlsym(h, "nvrtcGetProgramLogSize");
            api.getProgramLog = (decltype(api.getProgramLog)) dlsym(h, "nvrtcGetProgramLog");
            api.destroyProgram = (decltype(api.destroyProgram)) dlsym(h, "nvrtcDestroyProgram");
            api.ok = api.createProgram && api.compileProgram && api.getCUBINSize && api.getCUBIN && api.getProgramLogSize && api.getProgramLog && api.destroyProgram;
        });
        return api;
    }
};

// device headers embedded at build time: { include name, text }
struct EmbeddedSource { const char* name; const unsigned char* text; };
#include "embedded_sources.inc"

}  // namespace

std::vector<char> compileCudaToCubin(const std::string& source, const std::string& name, std::string* log)
{
    NvrtcApi& rtc = NvrtcApi::get();
    if (!rtc.ok) throw Error(PG_ERR_UNSUPPORTED, "libnvrtc.so.12 cannot be loaded: fused pipeline kernels are unavailable");
    std::vector<const char*> names, texts;
    for (const EmbeddedSource* e = kEmbeddedSources; e->name; e++) { names.push_back(e->name); texts.push_back((const char*) e->text); }
    nvrtcProgram prog = nullptr;
    if (rtc.createProgram(&prog, source.c_str(), name.c_str(), (int) names.size(), texts.data(), names.data()) != NVRTC_SUCCESS)
        throw Error(PG_ERR_INVALID_STATE, "nvrtcCreateProgram failed");
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=false", "-default-device"};
    const nvrtcResult r = rtc.compileProgram(prog, 5, opts);
    size_t logSize = 0;
    rtc.getProgramLogSize(prog, &logSize);
    std::string text(logSize, '\0');
    if (logSize > 1) rtc.getProgramLog(prog, &text[0]);
    if (log) *log = text;
    if (r != NVRTC_SUCCESS) { rtc.destroyProgram(&prog); throw Error(PG_ERR_INVALID_STATE, "NVRTC could not compile a generated pipeline kernel:\n" + text); }
    size_t size = 0;
    rtc.getCUBINSize(prog, &size);
    std::vector<char> cubin(size);
    rtc.getCUBIN(prog, cubin.data());
    rtc.destroyProgram(&prog);
    if (const char* dir = getenv("PG_JIT_DUMP_DIR")) {     // keep the generated source + cubin (cuobjdump -sass / --dump-resource-usage evidence)
        static int serial = 0;
        const std::string base = std::string(dir) + "/" + name + "." + std::to_string(serial++);
        if (FILE* f = fopen((base + ".cu").c_str(), "w")) { fwrite(source.data(), 1, source.size(), f); fclose(f); }
        if (FILE* f = fopen((base + ".cubin").c_str(), "wb")) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
    }
    return cubin;
}

const JitKernel* getOrCompileKernel(int device, const std::string& source, const char* entry, std::string* error)
{
    struct Module { CUmodule module = nullptr; std::map<std::string, JitKernel> kernels; };
    static std::mutex mu;
    static std::map<std::pair<int, std::string>, Module> cache;
    static std::map<std::pair<int, std::string>, std::string> failures;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(device, source);
    auto f = failures.find(key);
    if (f != failures.end()) { if (error) *error = f->second; return nullptr; }
    try {
        DriverApi& api = DriverApi::get();
        if (!api.ok) throw Error(PG_ERR_UNSUPPORTED, "libcuda.so.1 cannot be loaded");
        auto it = cache.find(key);
        if (it == cache.end()) {
            const std::vector<char> cubin = compileCudaToCubin(source, "fused_pipeline.cu");
            Module m;
            if (api.moduleLoadData(&m.module, cubin.data()) != CUDA_SUCCESS) throw Error(PG_ERR_CUDA, "cuModuleLoadData of a generated pipeline kernel failed");
            it = cache.emplace(key, m).first;
        }
        Module& m = it->second;
        auto k = m.kernels.find(entry);
        if (k != m.kernels.end()) return &k->second;
        JitKernel jk; jk.module = m.module;
        if (api.moduleGetFunction(&jk.function, m.module, entry) != CUDA_SUCCESS) throw Error(PG_ERR_CUDA, std::string("generated module lacks the entry point ") + entry);
        api.funcGetAttribute(&jk.registers, CU_FUNC_ATTRIBUTE_NUM_REGS, jk.function);
        return &(m.kernels[entry] = jk);
    }
    catch (const std::exception& e) {
        failures[key] = e.what();
        if (error) *error = e.what();
        return nullptr;
    }
}

CUmodule getOrLoadPtxModule(int device, const std::string& ptx, std::string* error)
{
    static std::mutex mu;
    static std::map<std::pair<int, std::string>, CUmodule> cache;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(device, ptx);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    DriverApi& api = DriverApi::get();
    if (!api.ok) { if (error) *error = "libcuda.so.1 cannot be loaded"; return nullptr; }
    char errLog[4096]; errLog[0] = 0;
    CUjit_option opts[2] = {CU_JIT_ERROR_LOG_BUFFER, CU_JIT_ERROR_LOG_BUFFER_SIZE_BYTES};
    void* vals[2] = {errLog, (void*) (uintptr_t) sizeof(errLog)};
    CUmodule module = nullptr;
    const CUresult r = api.moduleLoadDataEx(&module, ptx.c_str(), 2, opts, vals);
    if (r != CUDA_SUCCESS) { if (error) *error = std::string("cuModuleLoadDataEx failed: ") + errLog; return nullptr; }
    cache[key] = module;
    return module;
}

}  // namespace pg

// Host-side entry for tests: compile a generated source without a GPU (NVRTC only), returning the cubin size or the compiler log.
extern "C" int32_t pg_jit_compile_check(const char* source, char* logBuffer, size_t capacity, int64_t* cubinBytes)
{
    try {
        std::string log;
        const std::vector<char> cubin = pg::compileCudaToCubin(source, "check.cu", &log);
        if (cubinBytes) *cubinBytes = (int64_t) cubin.size();
        if (logBuffer && capacity) { snprintf(logBuffer, capacity, "%s", log.c_str()); }
    }
    catch (const pg::Error& e) {
        pg::setLastError(e.what());
        if (logBuffer && capacity) snprintf(logBuffer, capacity, "%s", e.what());
        return e.code;
    }
    return PG_OK;
}
