// cz_nn.cuh — interface between the search engine and the network runtime (cz_nn.cu).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "../../include/cczero_b200.h"

namespace cznn {

struct NnRuntime;

// bytes of device workspace the runtime needs for batches up to max_batch positions
size_t nn_workspace_bytes(int filters, int blocks, int value_fc, int max_batch, int n_nets);
// returns nullptr and sets cz_last_error on failure
NnRuntime* nn_create(int device, int filters, int blocks, int value_fc, int max_batch, void* workspace, size_t bytes,
                     void* stream, int fp32_skip_mode, int n_nets, int in_planes /* 14 or 28 */);
void nn_destroy(NnRuntime*);
int nn_set_weights(NnRuntime*, int net, const cz_tensor_desc* descs, int n);
bool nn_ready(const NnRuntime*);
// boards_dev: [batch][96] packed boards ([batch][2][96] = board, history board when in_planes = 28); policy_dev [batch][2086] f32 softmax; value_dev [batch] f32
int nn_forward_boards(NnRuntime*, int net, const uint8_t* boards_dev, int batch, float* policy_dev, float* value_dev);
int nn_forward_planes(NnRuntime*, int net, const float* planes_dev, int batch, float* policy_dev, float* value_dev);
uint64_t nn_launches(const NnRuntime*);
void nn_profile(NnRuntime*, bool on);
int nn_profile_read(NnRuntime*, double* ms, uint64_t* launches, double* flops);

}  // namespace cznn
