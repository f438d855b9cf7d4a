// smz_net_vision.cu — network step of the vision (ResNet-v2) model family, fp32 CUDA cores (parity mode).
//
// Reference: neural_network_vision_model.py — Residual_block v2 with one shared BatchNorm2d and conv_1 used
// twice (:41-79), Down_sample (:81-119), Representation (:122-158), Dynamics / Afterstate_dynamics
// (:161-226, :373-430), Prediction / Afterstate_prediction (:229-296, :433-492), channel-wise
// scale_to_bound_action (:495-503); muzero_model.py facade for RGB models: action as a constant plane
// (a+1)/A (:511-522), softmax on the policy, inverse_transform_with_support on value / reward.
// BatchNorm runs in eval mode and is folded at pack time into a per-channel scale / shift.
//
// Hidden state = [3,7,7] = 147 floats (arena rows of 160).  The 3-channel convolutions are CUDA-core work
// (no tensor-core shape); per simulation one CTA pushes 32 leaves through trunk convs + the 147->H->..->S/A
// MLP heads with activations in shared memory and the head weights streamed through a cp.async ring.
// The representation (98x98x3 -> 3x7x7) runs once per move, one CTA per tree, all feature maps in smem.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/smz.h"
#include "smz_net_vision.h"
#include "smz_net_tc32.h"

namespace {

constexpr int R = 32;               // leaves per CTA
constexpr int NT = 256;
constexpr int HW = 7, PIX = 49, FLAT = 147;
constexpr int VSP = SMZ_VISION_SP;  // 160: hidden row stride == padded K of the first MLP layer
constexpr int LD = VSP + 4;         // activation row stride of the MLP stage
constexpr int KC = 32;
constexpr float BN_EPS = 1e-5f;

struct VRes { const float *bn_s, *bn_t, *c1, *c3; };
struct VMlp { const float *in_wt, *in_b, *mid_wt, *mid_b, *out_wt, *out_b; };
struct VDyn { const float *conv, *bn_s, *bn_t; VRes res; const float *cr_w, *cr_b; VMlp reward; };
struct VPred { VRes res; const float *cv_w, *cv_b; VMlp value; const float *cp_w, *cp_b; VMlp policy; };
struct VRepr { const float* conv_in; VRes res_in; const float* conv_out; VRes res_out, res_last; };
struct VNets { VRepr repr; VDyn dyn, adyn; VPred pred, apred; int A, S, H, L; };

constexpr int RC = 8;               // leaves per CTA of the convolution-only stage (1024 trees = 128 CTAs on 148 SMs; measured on
                                    // cfg5: 32 leaves 8.1, 8 leaves 13.9, 4 leaves 11.8 M sims/s)
constexpr int NTC = 256;            // threads of that stage (512 measured slower: 11.7 M sims/s)
template <int RT>                   // RT leaves per CTA
struct SmemSim {
  float x[RT][LD];                // MLP ping / rows entering an MLP head
  float y[RT == R ? R : 1][LD];   // MLP pong (CUDA-core heads only)
  float w[2][RT == R ? KC : 1][SMZ_HP];   // weight ring (CUDA-core heads only)
  float in4[RT][4 * PIX];         // trunk input: state + action plane
  float f0[RT][FLAT], f1[RT][FLAT], f2[RT][FLAT];   // feature maps
  int tree[RT], slot[RT], act[RT];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// out[r][c] = act( sum_k in[r][k] * Wt[k][c] + b[c] ), 32 rows x 128 columns, K a multiple of 32 (ReLU heads)
__device__ void dense(const float* __restrict__ wt, int K, const float* __restrict__ bias, const float (*in)[LD],
                      float (*out)[LD], bool relu, float (*wbuf)[KC][SMZ_HP]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int nchunks = K / KC;
  auto issue = [&](int c) {
    const float* src = wt + (size_t)c * KC * SMZ_HP;
    float* dst = &wbuf[c & 1][0][0];
#pragma unroll
    for (int i = 0; i < (KC * SMZ_HP / 4) / NT; ++i) {
      const int e = (i * NT + tid) * 4;
      cp_async16(dst + e, src + e);
    }
    cp_async_commit();
  };
  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) { issue(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float(*w)[SMZ_HP] = wbuf[c & 1];
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      const float4 a0 = *reinterpret_cast<const float4*>(&in[ty * 2][c * KC + kk]);
      const float4 a1 = *reinterpret_cast<const float4*>(&in[ty * 2 + 1][c * KC + kk]);
      const float av0[4] = {a0.x, a0.y, a0.z, a0.w}, av1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b0 = *reinterpret_cast<const float4*>(&w[kk + q][tx * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&w[kk + q][tx * 8 + 4]);
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] = fmaf(av0[q], bv[j], acc[0][j]);
          acc[1][j] = fmaf(av1[q], bv[j], acc[1][j]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = tx * 8 + j;
      const float v = acc[i][j] + bias[c];
      out[ty * 2 + i][c] = relu ? fmaxf(v, 0.f) : v;
    }
  __syncthreads();
}

// MLP head (vision:188-199 etc.): Linear(147,H) ReLU, L x [tied Linear(H,H) ReLU], Linear(H,n).  Input rows in
// `in` (columns 147..159 zero).  Returns the buffer holding the n output logits (columns [0,n)).
__device__ float (*mlp_head(const VMlp& m, int L, float (*in)[LD], float (*other)[LD], float (*wbuf)[KC][SMZ_HP]))[LD] {
  float(*cur)[LD] = in;
  float(*nxt)[LD] = other;
  dense(m.in_wt, VSP, m.in_b, cur, nxt, true, wbuf);
  { auto t = cur; cur = nxt; nxt = t; }
  for (int l = 0; l < L; ++l) {
    dense(m.mid_wt, SMZ_HP, m.mid_b, cur, nxt, true, wbuf);
    auto t = cur; cur = nxt; nxt = t;
  }
  dense(m.out_wt, SMZ_HP, m.out_b, cur, nxt, false, wbuf);
  return nxt;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inverse_transform_with_support (muzero_model.py:575-591), S <= 64 logits, one warp per row
__device__ float support_scalar(const float* logits, int S) {
  const int lane = threadIdx.x & 31;
  const float v0 = lane < S ? logits[lane] : -INFINITY, v1 = lane + 32 < S ? logits[lane + 32] : -INFINITY;
  const float m = warp_max(fmaxf(v0, v1));
  const float e0 = lane < S ? expf(v0 - m) : 0.f, e1 = lane + 32 < S ? expf(v1 - m) : 0.f;
  const float z = warp_sum(e0 + e1);
  const int half = S / 2;
  const float y = warp_sum((float)(lane - half) * (e0 / z) + (float)(lane + 32 - half) * (e1 / z));
  const float inner = __fadd_rn(1.f, __fmul_rn(0.004f, __fadd_rn(__fadd_rn(fabsf(y), 1.f), 0.001f)));
  const float t = __fdiv_rn(__fsub_rn(__fsqrt_rn(inner), 1.f), 0.002f);
  const float mag = __fsub_rn(__fmul_rn(t, t), 1.f);
  return y > 0.f ? mag : (y < 0.f ? -mag : 0.f);
}
__device__ void policy_softmax(const float* logits, int n, float* dst) {
  const int lane = threadIdx.x & 31;
  const float v = lane < n ? logits[lane] : -INFINITY;
  const float m = warp_max(v);
  const float e = lane < n ? expf(v - m) : 0.f;
  const float z = warp_sum(e);
  if (lane < n) dst[lane] = e / z;
}

// ---- generic direct convolution over maps held in shared memory ---------------------------------------------
// out[n][co][y][x] = sum_{ci,ky,kx} act(in[n][ci][y*s+ky-1][x*s+kx-1]) * w[co][ci][ky][kx]   (zero padded)
// with act(v) = relu(v * bn_s[ci] + bn_t[ci]) when bn_s != null (pre-activation of the v2 block), else v.
// Optional residual add.  n_img images, strides in floats.
__device__ void conv3x3(const float* in, int in_stride, int Ci, int Hi, int Wi, int stride, float* out, int out_stride,
                        int Co, const float* __restrict__ w, const float* __restrict__ bn_s,
                        const float* __restrict__ bn_t, const float* res, int res_stride, int n_img) {
  const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
  const int per = Co * Ho * Wo;
  for (int o = threadIdx.x; o < n_img * per; o += blockDim.x) {
    const int n = o / per, rem = o - n * per;
    const int co = rem / (Ho * Wo), p = rem - co * (Ho * Wo);
    const int oy = p / Wo, ox = p - oy * Wo;
    const float* src = in + (size_t)n * in_stride;
    float acc = 0.f;
    for (int ci = 0; ci < Ci; ++ci) {
      const float s = bn_s ? __ldg(bn_s + ci) : 1.f, t = bn_s ? __ldg(bn_t + ci) : 0.f;
      const float* wk = w + (co * Ci + ci) * 9;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * stride + ky - 1;
        if (iy < 0 || iy >= Hi) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * stride + kx - 1;
          if (ix < 0 || ix >= Wi) continue;
          float v = src[(ci * Hi + iy) * Wi + ix];
          if (bn_s) v = fmaxf(fmaf(v, s, t), 0.f);
          acc = fmaf(v, __ldg(wk + ky * 3 + kx), acc);
        }
      }
    }
    if (res) acc += res[(size_t)n * res_stride + rem];
    out[(size_t)n * out_stride + rem] = acc;
  }
  __syncthreads();
}

// 3x3 convolution specialised for the per-simulation 7x7 maps: one thread per (leaf, pixel) produces all 3
// output channels from a register-resident copy of the (at most 108) weights — the inner loop is pure FFMA
// on shared-memory taps, no dependent global / constant loads.
template <int CI, bool PRE>
__device__ void conv7(const float* in, int in_stride, float* out, int out_stride, const float* __restrict__ wg,
                      const float* __restrict__ bn_s, const float* __restrict__ bn_t, const float* res, int res_stride,
                      int n_img) {
  float w[3 * CI * 9];
#pragma unroll
  for (int i = 0; i < 3 * CI * 9; ++i) w[i] = __ldg(wg + i);
  float s[CI], t[CI];
#pragma unroll
  for (int c = 0; c < CI; ++c) { s[c] = PRE ? __ldg(bn_s + c) : 1.f; t[c] = PRE ? __ldg(bn_t + c) : 0.f; }
  for (int item = threadIdx.x; item < n_img * PIX; item += blockDim.x) {
    const int n = item / PIX, p = item - n * PIX, y = p / HW, x = p - y * HW;
    const float* src = in + (size_t)n * in_stride;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
    for (int ci = 0; ci < CI; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int iy = y + ky - 1, ix = x + kx - 1;
          float v = 0.f;
          if (iy >= 0 && iy < HW && ix >= 0 && ix < HW) {
            v = src[ci * PIX + iy * HW + ix];
            if (PRE) v = fmaxf(fmaf(v, s[ci], t[ci]), 0.f);
          }
          acc0 = fmaf(v, w[(0 * CI + ci) * 9 + ky * 3 + kx], acc0);
          acc1 = fmaf(v, w[(1 * CI + ci) * 9 + ky * 3 + kx], acc1);
          acc2 = fmaf(v, w[(2 * CI + ci) * 9 + ky * 3 + kx], acc2);
        }
    if (res) {
      const float* rr = res + (size_t)n * res_stride + p;
      acc0 += rr[0]; acc1 += rr[PIX]; acc2 += rr[2 * PIX];
    }
    float* dst = out + (size_t)n * out_stride + p;
    dst[0] = acc0; dst[PIX] = acc1; dst[2 * PIX] = acc2;
  }
  __syncthreads();
}

// the same residual block on batches of 7x7 maps (per-simulation networks)
__device__ void resblock7(const VRes& r, const float* x, float* t1, float* t2, int n_img) {
  conv7<3, true>(x, FLAT, t1, FLAT, r.c1, r.bn_s, r.bn_t, nullptr, 0, n_img);
  conv7<3, true>(t1, FLAT, t2, FLAT, r.c3, r.bn_s, r.bn_t, nullptr, 0, n_img);
  conv7<3, true>(t2, FLAT, t1, FLAT, r.c1, r.bn_s, r.bn_t, x, FLAT, n_img);
}

// Residual_block v2 (vision:41-79): x -> [bn relu conv1] -> [bn relu conv3] -> [bn relu conv1] -> + x.
// x, t1, t2 are three distinct buffers of equal geometry; the result lands in t1.
__device__ void resblock(const VRes& r, const float* x, float* t1, float* t2, int stride, int C, int H, int W, int n_img) {
  conv3x3(x, stride, C, H, W, 1, t1, stride, C, r.c1, r.bn_s, r.bn_t, nullptr, 0, n_img);
  conv3x3(t1, stride, C, H, W, 1, t2, stride, C, r.c3, r.bn_s, r.bn_t, nullptr, 0, n_img);
  conv3x3(t2, stride, C, H, W, 1, t1, stride, C, r.c1, r.bn_s, r.bn_t, x, stride, n_img);
}

// 1x1 convolution with bias, Ci -> 3 channels, on 7x7 maps; output flattened [n][147] into an MLP input row (LD)
__device__ void conv1x1_to_rows(const float* in, int in_stride, int Ci, const float* __restrict__ w,
                                const float* __restrict__ b, float (*rows)[LD], int n_img) {
  for (int o = threadIdx.x; o < n_img * VSP; o += blockDim.x) {
    const int n = o / VSP, e = o - n * VSP;
    float acc = 0.f;
    if (e < FLAT) {
      const int co = e / PIX, p = e - co * PIX;
      acc = __ldg(b + co);
      for (int ci = 0; ci < Ci; ++ci) acc = fmaf(in[(size_t)n * in_stride + ci * PIX + p], __ldg(w + co * Ci + ci), acc);
    }
    rows[n][e] = acc;     // columns 147..159 are the zero padding of K
  }
  __syncthreads();
}

// scale_to_bound_action over the CHANNEL dimension (vision:495-503): per pixel min/max of the 3 channels.
// Optionally applies ReLU first (the trunk ends with an activation, vision:205).  In place.
__device__ void scale_channels(float* f, int stride, bool relu_first, int n_img) {
  for (int o = threadIdx.x; o < n_img * PIX; o += blockDim.x) {
    const int n = o / PIX, p = o - n * PIX;
    float* q = f + (size_t)n * stride + p;
    float a = q[0], b = q[PIX], c = q[2 * PIX];
    if (relu_first) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); c = fmaxf(c, 0.f); }
    const float lo = fminf(a, fminf(b, c)), hi = fmaxf(a, fmaxf(b, c));
    float sc = hi - lo;
    if (sc < 1e-5f) sc += 1e-5f;
    q[0] = (a - lo) / sc; q[PIX] = (b - lo) / sc; q[2 * PIX] = (c - lo) / sc;
  }
  __syncthreads();
}

struct VJob {
  int mode;             // 0 simulation (compacted rows), 1 root prediction (trees in order), 2 stand-alone eval
  int which;            // eval: 1 pred, 2 adyn, 3 apred, 4 dyn
  int n_rows;
  const float* in;      // eval: rows [n][160]
  const int* idx;       // eval: action / code per row
  float* hidden_dst;    // [index][160]
  float* policy_dst; float* value_dst; float* reward_dst;
  int pstride;
  float* feat;          // FEAT stage: [3 heads][2 branches][B rows][160] inputs of the MLP heads (tensor-core stage)
};

// MLP-head input rows of a tile -> the feature scratch the tensor-core stage reads
template <int RT>
__device__ void rows_to_feat(const float (*rows)[LD], float* feat, int head, int branch, int B, int tile, int n_valid) {
  float* dst = feat + ((size_t)(head * 2 + branch) * B + (size_t)tile * RT) * VSP;
  for (int e = threadIdx.x; e < RT * VSP; e += blockDim.x) {
    const int r = e / VSP, c = e - r * VSP;
    if (r < n_valid) dst[(size_t)r * VSP + c] = rows[r][c];
  }
  __syncthreads();
}

// FEAT = false: the whole step on the CUDA cores (three CTAs per tile, one per MLP head).  FEAT = true (simulation
// step only): the convolution stage alone — trunk, new hidden state, residual trunk of the prediction net and the
// three 1x1 convolutions; the rows that enter the MLP heads go to `job.feat`, the heads run on the tensor cores
// (smz_tc32_vision_heads).
template <bool FEAT, int R>
__global__ void __launch_bounds__(FEAT ? NTC : NT) k_vision_step(SmzArena a, VNets nets, VJob job, int sim) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemSim<R>& sm = *reinterpret_cast<SmemSim<R>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  smz_pdl_wait();                 // no-op unless launched with a programmatic dependency (simulation step)
  smz_pdl_launch_dependents();
  // Three CTAs share a tile of 32 leaves, one per MLP head (the heads are ~80 % of the work and independent):
  // head 0 = reward (dynamics pair only), 1 = value (also writes the new hidden state), 2 = policy.  The cheap
  // convolution trunk is recomputed by each.
  const int head = FEAT ? 1 : blockIdx.x % 3;
  int tile = FEAT ? blockIdx.x : blockIdx.x / 3, branch = 0, count = job.n_rows;
  bool do_trunk = true, do_pred = true;
  if (job.mode == 0) {
    const int n0 = a.branch_count[sim * 2 + 0], n1 = a.branch_count[sim * 2 + 1];
    const int t0 = (n0 + R - 1) / R, t1 = (n1 + R - 1) / R;
    if (tile < t0) { branch = 0; count = n0; }
    else if (tile < t0 + t1) { branch = 1; count = n1; tile -= t0; }
    else return;
  } else {
    if (tile * R >= count) return;
    if (job.mode == 1) { do_trunk = false; branch = 1; }              // root: Prediction on slot 0
    else {
      do_trunk = (job.which == 2 || job.which == 4);
      do_pred = (job.which == 1 || job.which == 3);
      branch = (job.which == 4 || job.which == 1) ? 1 : 0;
    }
  }
  if (!FEAT) {
    if (head == 0 && !(do_trunk && branch && (job.mode == 0 || job.which == 4))) return;   // no reward head here
    if (head != 0 && !do_pred && !(head == 1 && do_trunk)) return;                            // nothing but a state write
  }
  const int n_valid = min(R, count - tile * R);
  if (tid < R) {
    const int row = tile * R + tid;
    int tree = -1, slot = 0, act = 0;
    if (row < count) {
      if (job.mode == 0) { const int4 rec = a.rows4[smz_row_index(a, sim, branch, row)]; tree = rec.x; slot = rec.y; act = rec.z; }
      else { tree = row; act = (job.mode == 2 && job.idx) ? job.idx[row] : 0; }
    }
    sm.tree[tid] = tree; sm.slot[tid] = slot; sm.act[tid] = act;
  }
  __syncthreads();
  // ---- load the input state of every row -----------------------------------------------------------------
  for (int e = tid; e < R * 4 * PIX; e += blockDim.x) {
    const int r = e / (4 * PIX), c = e - r * (4 * PIX);
    const int tree = sm.tree[r];
    float v = 0.f;
    if (tree >= 0) {
      if (c < FLAT) {
        const float* src = (job.mode == 2) ? job.in + (size_t)tree * VSP
                                           : a.hidden + ((size_t)sm.slot[r] * a.B + tree) * VSP;
        v = src[c];
      } else {
        v = ((float)sm.act[r] + 1.f) / (float)nets.A;       // muzero_model.py:511-522
      }
    }
    sm.in4[r][c] = v;
    if (c < FLAT) sm.f0[r][c] = v;
  }
  __syncthreads();
  float* state = &sm.f0[0][0];        // where the (new) state lives: [R][147]
  if (do_trunk) {
    const VDyn& d = branch ? nets.dyn : nets.adyn;
    // conv(4->3) + BN + ReLU (folded: BN scale/shift applied to the conv OUTPUT, then ReLU)
    conv7<4, false>(&sm.in4[0][0], 4 * PIX, &sm.f0[0][0], FLAT, d.conv, nullptr, nullptr, nullptr, 0, R);
    for (int e = tid; e < R * FLAT; e += blockDim.x) {
      const int c = (e % FLAT) / PIX;
      float* q = &sm.f0[0][0] + e;
      *q = fmaxf(fmaf(*q, __ldg(d.bn_s + c), __ldg(d.bn_t + c)), 0.f);
    }
    __syncthreads();
    float *x = &sm.f0[0][0], *t1 = &sm.f1[0][0], *t2 = &sm.f2[0][0];
    for (int l = 0; l < nets.L; ++l) {
      resblock7(d.res, x, t1, t2, R);
      float* t = x; x = t1; t1 = t;
    }
    scale_channels(x, FLAT, true, R);
    state = x;
    for (int e = tid; e < R * VSP; e += blockDim.x) {
      const int r = e / VSP, c = e - r * VSP;
      const int tree = sm.tree[r];
      if (head == 1 && tree >= 0 && job.hidden_dst) job.hidden_dst[(size_t)tree * VSP + c] = c < FLAT ? state[r * FLAT + c] : 0.f;
    }
    if (FEAT) {
      if (branch) {       // reward head input (vision:180, :207-215): conv1x1(4->3) on the INPUT x, flattened
        conv1x1_to_rows(&sm.in4[0][0], 4 * PIX, 4, d.cr_w, d.cr_b, sm.x, R);
        rows_to_feat<R>(sm.x, job.feat, 0, branch, a.B, tile, n_valid);
      }
    } else if (head == 0 && branch && (job.mode == 0 || job.which == 4)) {
      if constexpr (!FEAT) {
        // reward head (vision:180, :207-215): conv1x1(4->3) on the INPUT x, flatten, MLP, categorical support
        conv1x1_to_rows(&sm.in4[0][0], 4 * PIX, 4, d.cr_w, d.cr_b, sm.x, R);
        float(*o)[LD] = mlp_head(d.reward, nets.L, sm.x, sm.y, sm.w);
        for (int r = warp; r < R; r += NT / 32) {
          const float rew = support_scalar(o[r], nets.S);
          if (sm.tree[r] >= 0 && lane == 0 && job.reward_dst) job.reward_dst[sm.tree[r]] = rew;
        }
        __syncthreads();
      }
    }
  }
  if (do_pred && head != 0) {
    const VPred& p = branch ? nets.pred : nets.apred;
    float* x = state;
    float* t1 = (x == &sm.f0[0][0]) ? &sm.f1[0][0] : &sm.f0[0][0];
    float* t2 = &sm.f2[0][0];
    if (x == t2) t2 = &sm.f1[0][0];
    for (int l = 0; l < nets.L; ++l) {
      resblock7(p.res, x, t1, t2, R);
      float* t = x; x = t1; t1 = t;
    }
    if (FEAT) {
      conv1x1_to_rows(x, FLAT, 3, p.cv_w, p.cv_b, sm.x, R);
      rows_to_feat<R>(sm.x, job.feat, 1, branch, a.B, tile, n_valid);
      conv1x1_to_rows(x, FLAT, 3, p.cp_w, p.cp_b, sm.x, R);
      rows_to_feat<R>(sm.x, job.feat, 2, branch, a.B, tile, n_valid);
    } else if constexpr (!FEAT) {
      if (head == 1) {
        conv1x1_to_rows(x, FLAT, 3, p.cv_w, p.cv_b, sm.x, R);
        float(*ov)[LD] = mlp_head(p.value, nets.L, sm.x, sm.y, sm.w);
        for (int r = warp; r < R; r += NT / 32) {
          const float val = support_scalar(ov[r], nets.S);
          if (sm.tree[r] >= 0 && lane == 0 && job.value_dst) job.value_dst[sm.tree[r]] = val;
        }
      } else {
        conv1x1_to_rows(x, FLAT, 3, p.cp_w, p.cp_b, sm.x, R);
        float(*op)[LD] = mlp_head(p.policy, nets.L, sm.x, sm.y, sm.w);
        for (int r = warp; r < R; r += NT / 32)
          if (sm.tree[r] >= 0 && job.policy_dst) policy_softmax(op[r], nets.A, job.policy_dst + (size_t)sm.tree[r] * job.pstride);
      }
    }
  }
}

// ---- representation: [3,98,98] -> [3,7,7], one CTA per tree --------------------------------------------------
constexpr int IMG = 98, H1 = 49, H2 = 25, H3 = 13;
struct SmemRepr {
  float a[3 * H2 * H2 > H1 * H1 ? 3 * H2 * H2 : H1 * H1];
  float b[3 * H2 * H2 > H1 * H1 ? 3 * H2 * H2 : H1 * H1];
  float c[3 * H2 * H2 > H1 * H1 ? 3 * H2 * H2 : H1 * H1];
};

__device__ void avgpool3s2(const float* in, int C, int Hi, float* out) {
  const int Ho = (Hi - 1) / 2 + 1;
  for (int o = threadIdx.x; o < C * Ho * Ho; o += blockDim.x) {
    const int c = o / (Ho * Ho), p = o - c * Ho * Ho, oy = p / Ho, ox = p - oy * Ho;
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = oy * 2 + ky - 1, ix = ox * 2 + kx - 1;
        if (iy >= 0 && iy < Hi && ix >= 0 && ix < Hi) acc += in[(c * Hi + iy) * Hi + ix];
      }
    out[o] = acc / 9.f;      // count_include_pad = True
  }
  __syncthreads();
}

// first strided conv reads the observation straight from global memory
__device__ void conv_in_s2(const float* __restrict__ obs, const float* __restrict__ w, float* out) {
  for (int o = threadIdx.x; o < H1 * H1; o += blockDim.x) {
    const int oy = o / H1, ox = o - oy * H1;
    float acc = 0.f;
    for (int ci = 0; ci < 3; ++ci)
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 + ky - 1;
        if (iy < 0 || iy >= IMG) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * 2 + kx - 1;
          if (ix < 0 || ix >= IMG) continue;
          acc = fmaf(obs[(ci * IMG + iy) * IMG + ix], __ldg(w + ci * 9 + ky * 3 + kx), acc);
        }
      }
    out[o] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT) k_vision_repr(VNets nets, int n_trees, const float* __restrict__ obs, float* hidden_dst) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemRepr& sm = *reinterpret_cast<SmemRepr*>(smem_raw);
  const int tree = blockIdx.x;
  if (tree >= n_trees) return;
  const VRepr& rp = nets.repr;
  float *x = sm.a, *t1 = sm.b, *t2 = sm.c, *t;
  conv_in_s2(obs + (size_t)tree * 3 * IMG * IMG, rp.conv_in, x);                     // [1,49,49]
  for (int i = 0; i < 2; ++i) { resblock(rp.res_in, x, t1, t2, 0, 1, H1, H1, 1); t = x; x = t1; t1 = t; }
  conv3x3(x, 0, 1, H1, H1, 2, t1, 0, 3, rp.conv_out, nullptr, nullptr, nullptr, 0, 1);  // [3,25,25]
  t = x; x = t1; t1 = t;
  for (int i = 0; i < 2; ++i) { resblock(rp.res_out, x, t1, t2, 0, 3, H2, H2, 1); t = x; x = t1; t1 = t; }
  avgpool3s2(x, 3, H2, t1);                                                          // [3,13,13]
  t = x; x = t1; t1 = t;
  for (int i = 0; i < 3; ++i) { resblock(rp.res_out, x, t1, t2, 0, 3, H3, H3, 1); t = x; x = t1; t1 = t; }
  avgpool3s2(x, 3, H3, t1);                                                          // [3,7,7]
  t = x; x = t1; t1 = t;
  resblock(rp.res_last, x, t1, t2, 0, 3, HW, HW, 1);
  x = t1;
  scale_channels(x, 0, false, 1);
  for (int c = threadIdx.x; c < VSP; c += blockDim.x) hidden_dst[(size_t)tree * VSP + c] = c < FLAT ? x[c] : 0.f;
}

// ---- weight image -----------------------------------------------------------------------------------------
__global__ void k_copy(float* __restrict__ dst, const float* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
// bn[4][c] = gamma, beta, mean, var  ->  scale[c], shift[c]
__global__ void k_fold_bn(float* __restrict__ s, float* __restrict__ t, const float* __restrict__ bn, int c) {
  const int i = threadIdx.x;
  if (i >= c) return;
  const float sc = bn[i] / sqrtf(bn[3 * c + i] + BN_EPS);
  s[i] = sc;
  t[i] = bn[c + i] - bn[2 * c + i] * sc;
}
// W[out][in] -> Wt[k][128] zero padded
__global__ void k_pack_wt(float* __restrict__ dst, const float* __restrict__ src, int n_out, int n_in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out * n_in) return;
  const int o = i / n_in, k = i % n_in;
  dst[(size_t)k * SMZ_HP + o] = src[(size_t)o * n_in + k];
}

}  // namespace

struct SmzVisionImage {
  VNets nets;
  float* pool;
  size_t pool_floats;
  uint64_t blob_floats;
  // tensor-core stage of the simulation step (null: SMZ_VISION_CC=1 or not available -> CUDA-core heads)
  SmzTc32VisionHeads* tc;
  SmzVisionHeadSrc head_src[5];     // afterstate value / policy, dynamics reward / value / policy
  float* feat;                      // [3][2][max_trees][160]
  int max_trees;
};

namespace {
struct Builder {
  const float* blob; float* pool; size_t boff = 0, poff = 0; cudaStream_t s; bool dry; int A, S, H, L;
  float* take(size_t n) { float* p = pool + poff; poff += (n + 3) / 4 * 4; return p; }
  const float* raw(size_t n) {       // copy n floats verbatim
    float* d = take(n);
    if (!dry) k_copy<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d, blob + boff, (int)n);
    boff += n;
    return d;
  }
  void bn(int c, const float** sc, const float** sh) {
    float *a = take(c), *b = take(c);
    if (!dry) k_fold_bn<<<1, 32, 0, s>>>(a, b, blob + boff, c);
    boff += 4 * (size_t)c;
    *sc = a; *sh = b;
  }
  VRes res(int c) { VRes r; bn(c, &r.bn_s, &r.bn_t); r.c1 = raw((size_t)c * c * 9); r.c3 = raw((size_t)c * c * 9); return r; }
  const float* wt(int n_out, int n_in, int k_pad) {
    float* d = take((size_t)k_pad * SMZ_HP);
    if (!dry) k_pack_wt<<<(n_out * n_in + 255) / 256, 256, 0, s>>>(d, blob + boff, n_out, n_in);
    boff += (size_t)n_out * n_in;
    return d;
  }
  const float* vec128(int n) {
    float* d = take(SMZ_HP);
    if (!dry) k_copy<<<1, 128, 0, s>>>(d, blob + boff, n);
    boff += n;
    return d;
  }
  SmzVisionHeadSrc last_src{};      // blob offsets of the head packed last (for the tensor-core image)
  VMlp mlp(int n_out) {
    VMlp m{};
    SmzVisionHeadSrc q{};
    q.n_out = n_out;
    q.in_w = boff; m.in_wt = wt(H, FLAT, VSP); q.in_b = boff; m.in_b = vec128(H);
    if (L > 0) { q.mid_w = boff; m.mid_wt = wt(H, H, SMZ_HP); q.mid_b = boff; m.mid_b = vec128(H); }
    q.out_w = boff; m.out_wt = wt(n_out, H, SMZ_HP); q.out_b = boff; m.out_b = vec128(n_out);
    last_src = q;
    return m;
  }
  SmzVisionHeadSrc src[5];
  VNets all() {
    VNets n{};
    n.A = A; n.S = S; n.H = H; n.L = L;
    n.repr.conv_in = raw(27); n.repr.res_in = res(1); n.repr.conv_out = raw(27); n.repr.res_out = res(3); n.repr.res_last = res(3);
    for (int i = 0; i < 2; ++i) {
      VDyn& d = i ? n.adyn : n.dyn;
      d.conv = raw(108); bn(3, &d.bn_s, &d.bn_t); d.res = res(3);
      if (i == 0) { d.cr_w = raw(12); d.cr_b = raw(3); d.reward = mlp(S); src[2] = last_src; }
    }
    for (int i = 0; i < 2; ++i) {
      VPred& p = i ? n.apred : n.pred;
      p.res = res(3);
      p.cv_w = raw(9); p.cv_b = raw(3); p.value = mlp(S); src[i ? 0 : 3] = last_src;
      p.cp_w = raw(9); p.cp_b = raw(3); p.policy = mlp(A); src[i ? 1 : 4] = last_src; src[i ? 1 : 4].is_policy = 1;
    }
    return n;
  }
};
}  // namespace

uint64_t smz_vision_blob_floats(int A, int S, int H, int L) {
  Builder b{nullptr, nullptr, 0, 0, nullptr, true, A, S, H, L};
  b.all();
  return b.boff;
}

int smz_vision_create(int A, int S, int H, int L, int max_trees, SmzVisionImage** out, char* err, size_t err_len) {
  if (H > SMZ_HP || S > SMZ_SP || A > 32 || L < 1 || L > 16) {
    snprintf(err, err_len, "vision network: need H<=%d, S<=%d, A<=32, 1<=L<=16 (got %d, %d, %d, %d)", SMZ_HP, SMZ_SP, H, S, A, L);
    return SMZ_E_CAPACITY;
  }
  SmzVisionImage* im = new SmzVisionImage();
  memset(im, 0, sizeof(*im));
  Builder b{nullptr, nullptr, 0, 0, nullptr, true, A, S, H, L};
  b.all();
  im->pool_floats = b.poff;
  im->blob_floats = b.boff;
  memcpy(im->head_src, b.src, sizeof(im->head_src));
  im->max_trees = max_trees;
  if (cudaMalloc(&im->pool, im->pool_floats * sizeof(float)) != cudaSuccess) {
    snprintf(err, err_len, "vision network: cudaMalloc of the weight image failed");
    delete im;
    return SMZ_E_CUDA;
  }
  im->nets.A = A; im->nets.S = S; im->nets.H = H; im->nets.L = L;
  // the MLP heads of the simulation step on the tensor cores (fp32-grade chain) unless SMZ_VISION_CC=1 asks for the
  // all-CUDA-core kernel; stand-alone evaluation and the root prediction always use the latter
  if (!getenv("SMZ_VISION_CC")) {
    char why[256] = "";
    if (smz_tc32_vision_create(A, S, H, L, &im->tc, why, sizeof(why)) != SMZ_OK) im->tc = nullptr;
    if (im->tc && cudaMalloc(&im->feat, (size_t)6 * max_trees * VSP * sizeof(float)) != cudaSuccess) {
      smz_tc32_vision_destroy(im->tc);
      im->tc = nullptr;
    }
  }
  cudaFuncSetAttribute((const void*)k_vision_step<false, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSim<R>));
  cudaFuncSetAttribute((const void*)k_vision_step<true, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSim<RC>));
  cudaFuncSetAttribute((const void*)k_vision_repr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemRepr));
  *out = im;
  return SMZ_OK;
}

void smz_vision_destroy(SmzVisionImage* im) {
  if (!im) return;
  if (im->tc) smz_tc32_vision_destroy(im->tc);
  cudaFree(im->feat);
  cudaFree(im->pool);
  delete im;
}

int smz_vision_pack(SmzVisionImage* im, const float* blob_dev, cudaStream_t s, char* err, size_t err_len) {
  if (cudaMemsetAsync(im->pool, 0, im->pool_floats * sizeof(float), s) != cudaSuccess) {
    snprintf(err, err_len, "vision network: memset failed");
    return SMZ_E_CUDA;
  }
  Builder b{blob_dev, im->pool, 0, 0, s, false, im->nets.A, im->nets.S, im->nets.H, im->nets.L};
  im->nets = b.all();
  if (im->tc) {
    const int rc = smz_tc32_vision_pack(im->tc, blob_dev, im->head_src, s, err, err_len);
    if (rc != SMZ_OK) return rc;
  }
  if (cudaGetLastError() != cudaSuccess) {
    snprintf(err, err_len, "vision network: weight packing launch failed");
    return SMZ_E_CUDA;
  }
  return SMZ_OK;
}

void smz_vision_root(SmzVisionImage* im, const SmzArena& a, int n_trees, const float* obs, cudaStream_t s) {
  k_vision_repr<<<n_trees, NT, sizeof(SmemRepr), s>>>(im->nets, n_trees, obs, a.hidden);
  VJob job{};
  job.mode = 1; job.n_rows = n_trees; job.policy_dst = a.out_policy; job.value_dst = a.out_value; job.pstride = a.W;
  k_vision_step<false, R><<<3 * ((n_trees + R - 1) / R), NT, sizeof(SmemSim<R>), s>>>(a, im->nets, job, 0);
}

bool smz_vision_has_tc(const SmzVisionImage* im) { return im != nullptr && im->tc != nullptr; }

// returns the number of kernels launched
int smz_vision_sim(SmzVisionImage* im, const SmzArena& a, int n_trees, int sim, bool pdl, cudaStream_t s) {
  VJob job{};
  job.mode = 0; job.n_rows = n_trees;
  job.hidden_dst = a.hidden + (size_t)(sim + 1) * a.B * VSP;
  job.policy_dst = a.out_policy; job.value_dst = a.out_value; job.reward_dst = a.out_reward; job.pstride = a.W;
  if (im->tc && n_trees <= im->max_trees) {
    job.feat = im->feat;
    // convolution stage -> head chains -> tree step are chained by programmatic dependent launch like the MLP family's
    // two kernels: each one's prologue (weight tiles, tree mirror) runs while its predecessor drains
    smz_launch(k_vision_step<true, RC>, dim3((n_trees + RC - 1) / RC + 1), dim3(NTC), sizeof(SmemSim<RC>), s, pdl, a, im->nets, job, sim);
    smz_tc32_vision_heads(im->tc, a, n_trees, sim, im->feat, pdl, s);
    return 2;
  }
  k_vision_step<false, R><<<3 * ((n_trees + R - 1) / R + 1), NT, sizeof(SmemSim<R>), s>>>(a, im->nets, job, sim);
  return 1;
}

int smz_vision_eval(SmzVisionImage* im, int which, int n_rows, const float* in, const int* idx, float* hidden_out,
                    float* policy_out, float* value_out, float* reward_out, int policy_stride, cudaStream_t s) {
  if (which == 0) {
    k_vision_repr<<<n_rows, NT, sizeof(SmemRepr), s>>>(im->nets, n_rows, in, hidden_out);
    return SMZ_OK;
  }
  if (which < 1 || which > 4) return SMZ_E_INVALID_ARG;
  VJob job{};
  job.mode = 2; job.which = which; job.n_rows = n_rows; job.in = in; job.idx = idx;
  job.hidden_dst = hidden_out; job.policy_dst = policy_out; job.value_dst = value_out; job.reward_dst = reward_out;
  job.pstride = policy_stride;
  SmzArena dummy{};
  k_vision_step<false, R><<<3 * ((n_rows + R - 1) / R), NT, sizeof(SmemSim<R>), s>>>(dummy, im->nets, job, 0);
  return SMZ_OK;
}
