// smz_net_tc32.h — fp32-grade network step on tcgen05 (fp16 hi/lo split operands, three products); see smz_net_tc32.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "smz_kernels.h"

struct SmzTc32Image;

// nprod 3: fp32-grade (hi/lo split, three products; SMZ_NET_TC32); nprod 1: plain fp16 operands (SMZ_NET_F16)
int smz_tc32_create(const SmzNetShape& sh, int nprod, SmzTc32Image** out, char* err, size_t err_len);
void smz_tc32_destroy(SmzTc32Image* im);
int smz_tc32_pack(SmzTc32Image* im, const SmzNetShape& sh, const float* blob_dev, cudaStream_t s, char* err, size_t err_len);
void smz_tc32_root(SmzTc32Image* im, const SmzArena& a, const SmzNetShape& sh, int n_trees, const float* obs, cudaStream_t s);
void smz_tc32_sim(SmzTc32Image* im, const SmzArena& a, const SmzNetShape& sh, int n_trees, int sim, bool pdl, cudaStream_t s);
void smz_tc32_eval(SmzTc32Image* im, const SmzNetShape& sh, int which, int n_rows, const float* in, const int* idx,
                   float* hidden_out, float* policy_out, float* value_out, float* reward_out, int* code_out,
                   int policy_stride, cudaStream_t s);
// the arena keeps hidden states as [hi 64 | lo 64] fp16 rows in this mode; this widens one slot to fp32 rows
void smz_tc32_read_hidden(const SmzArena& a, int slot, int n_trees, float* out, cudaStream_t s);
