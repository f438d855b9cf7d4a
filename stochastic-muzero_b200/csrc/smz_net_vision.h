// smz_net_vision.h — vision (ResNet-v2) model family: fp32 CUDA-core convolutions, MLP heads of the simulation step on
// the tensor cores (fp32-grade chain of smz_net_tc32.cu); see smz_net_vision.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "smz_common.cuh"

struct SmzVisionImage;

uint64_t smz_vision_blob_floats(int A, int S, int H, int L);
int smz_vision_create(int A, int S, int H, int L, int max_trees, SmzVisionImage** out, char* err, size_t err_len);
void smz_vision_destroy(SmzVisionImage* im);
int smz_vision_pack(SmzVisionImage* im, const float* blob_dev, cudaStream_t s, char* err, size_t err_len);
// representation (obs float[n][3][98][98]) -> hidden slot 0, then Prediction -> out_policy / out_value
void smz_vision_root(SmzVisionImage* im, const SmzArena& a, int n_trees, const float* obs, cudaStream_t s);
// one simulation's network step; returns the number of kernels launched (convolution stage + tensor-core heads, or the
// single all-CUDA-core kernel with SMZ_VISION_CC=1)
bool smz_vision_has_tc(const SmzVisionImage* im);      // the head chains run on the tensor cores (PDL-chained step)
int smz_vision_sim(SmzVisionImage* im, const SmzArena& a, int n_trees, int sim, bool pdl, cudaStream_t s);
// which: 0 repr(obs) 1 pred 2 adyn 3 apred 4 dyn; hidden rows are float[n][SMZ_VISION_SP]
int smz_vision_eval(SmzVisionImage* im, int which, int n_rows, const float* in, const int* idx, float* hidden_out,
                    float* policy_out, float* value_out, float* reward_out, int policy_stride, cudaStream_t s);
