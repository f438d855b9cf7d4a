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

// ---- vision family: the 147 -> H -> .. -> S / A MLP heads of a simulation step on the chain kernel (fp32-grade) ----
struct SmzTc32VisionHeads;
// where one head's Linear layers sit in the vision weight blob (offsets in floats)
struct SmzVisionHeadSrc {
  size_t in_w, in_b, mid_w, mid_b, out_w, out_b;
  int n_out;       // S (categorical support) or A (policy)
  int is_policy;
};
int smz_tc32_vision_create(int A, int S, int H, int L, SmzTc32VisionHeads** out, char* err, size_t err_len);
void smz_tc32_vision_destroy(SmzTc32VisionHeads* im);
// src[5]: afterstate value, afterstate policy, dynamics reward, dynamics value, dynamics policy
int smz_tc32_vision_pack(SmzTc32VisionHeads* im, const float* blob_dev, const SmzVisionHeadSrc* src, cudaStream_t s, char* err,
                         size_t err_len);
// feat: [3 heads: reward, value, policy][2 branches][a.B rows][160] fp32, rows in the compacted order of the simulation
void smz_tc32_vision_heads(SmzTc32VisionHeads* im, const SmzArena& a, int n_trees, int sim, const float* feat, bool pdl, cudaStream_t s);
