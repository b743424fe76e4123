// Library-level entry points of the C ABI: error string, version, layer stream, launch counter.
#include <atomic>
#include <cuda_runtime.h>

#include "../../include/unet_b200.h"
#include "host_common.h"

static cudaStream_t g_layer_stream = nullptr;  // legacy default stream, like the reference
static std::atomic<unsigned long long> g_launches{0};

cudaStream_t ub_layer_stream() { return g_layer_stream; }

extern "C" {
const char* ub_last_error(void) { return ub_host_last_error(); }
const char* ub_version(void) { return "unet_b200 0.1 (sm_100a: tcgen05 implicit-GEMM conv, NHWC bf16 training path)"; }
int ub_set_stream(void* stream) {
    g_layer_stream = reinterpret_cast<cudaStream_t>(stream);
    return UB_OK;
}
unsigned long long ub_launch_count(void) { return g_launches.load(); }
void ub_count_launches(unsigned long long n) { g_launches += n; }
}
