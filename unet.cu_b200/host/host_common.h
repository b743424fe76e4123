// Shared between the host translation units of libunet_b200.
#pragma once
#include <cuda_runtime.h>

const char* ub_host_last_error();
void ub_host_set_error(const char* s);
// stream used by the layer operators (ub_set_stream)
cudaStream_t ub_layer_stream();
