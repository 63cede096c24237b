// presto_b200 — run-time compilation of generated kernels.
//
// The reference compiles every query's expressions (and join / hash strategies) to JVM bytecode at run time
// (MB/sql/gen/PageFunctionCompiler.java:195-861, ExpressionCompiler.java:144-189, JoinCompiler.java:640).  Here:
//   * PTX produced by fp_codegen.cpp goes to the driver's JIT (cuModuleLoadDataEx);
//   * CUDA C++ produced by fused_codegen.cpp (a pipeline fragment's row program on top of fused_agg.cuh) goes through NVRTC to a
//     sm_100a cubin; the device headers it includes are embedded in libprestogpu.so at build time (embedded_sources.inc).
// libcuda and libnvrtc are dlopen'ed so that the library still loads (and the codegen can be tested) on a box without a driver.
#pragma once
#include <cuda.h>
#include <string>
#include <vector>

namespace pg {

struct DriverApi {
    CUresult (*moduleLoadDataEx)(CUmodule*, const void*, unsigned int, CUjit_option*, void**) = nullptr;
    CUresult (*moduleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*moduleUnload)(CUmodule) = nullptr;
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*funcGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*occupancyMaxActiveBlocks)(int*, CUfunction, int, size_t) = nullptr;
    bool ok = false;
    static DriverApi& get();
};

// CUDA C++ → cubin for sm_100a.  `source` may #include the embedded device headers by their file names ("fused_agg.cuh", …).
// Throws pg::Error (PG_ERR_UNSUPPORTED when NVRTC cannot be loaded, PG_ERR_INVALID_STATE with the compiler log on a compile error).
std::vector<char> compileCudaToCubin(const std::string& source, const std::string& name, std::string* log = nullptr);

// Process-wide cache of compiled fused kernels: operators of the same plan shape (every task of a query, every benchmark step)
// share one module per device.  Returns nullptr + *error when the source does not compile or the module cannot be loaded.
struct JitKernel { CUmodule module = nullptr; CUfunction function = nullptr; int registers = 0; };
const JitKernel* getOrCompileKernel(int device, const std::string& source, const char* entry, std::string* error);

// Process-wide cache of driver-JIT modules (PTX from fp_codegen.cpp): cuModuleLoadDataEx costs ~0.3 ms per module, and every query of
// the same plan shape generates the same PTX.  Returns nullptr + *error when the PTX does not assemble.  Modules live until the
// process exits (like the fused pipeline modules).
CUmodule getOrLoadPtxModule(int device, const std::string& ptx, std::string* error);

}  // namespace pg
