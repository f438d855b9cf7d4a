// smz_tree.cu — tree kernels of the batched Stochastic-MuZero search (sm_100a).
//
// G threads ("lanes") of a warp cooperate on one tree, one lane per child / per path node; G is a
// template parameter (2..32, 32 = warp per tree).  All arithmetic that decides a visit order follows
// the reference's numpy semantics exactly: explicit round-to-nearest
// intrinsics (no FMA contraction), float64 pUCT prior term, float32 value term, numpy's pairwise
// float32 summation order, RandomState.choice's cdf/searchsorted algorithm.
//
//   k_root_expand    monte_carlo_tree_search.py:203-225  root children + Dirichlet mixing
//   k_select         :235-267   pUCT argmax / chance sampling descent, leaf record, branch compaction
//   k_expand_backup  :289-308   children of the leaf + min-max discounted backup
//   k_read_roots     game.py:179-204 consumer view (child visits, priors, rewards, root value)
#include "smz_common.cuh"
#include "smz_kernels.h"

namespace {

#define FULL 0xffffffffu

// Every kernel below keeps all 32 lanes of a warp converged: loops run for the warp-wide maximum trip
// count and per-group work is predicated, so every shuffle / ballot is a full-mask, constant-mask
// instruction on sub-warp segments of width G (no mask matching, no dependence on independent
// thread scheduling).  Trees beyond n_trees keep their lanes alive with `alive == false`.
template <int G>
struct Group {
  int lane, gl, gbase;
  __device__ Group() {
    lane = threadIdx.x & 31;
    gl = lane & (G - 1);
    gbase = lane & ~(G - 1);
  }
  template <typename T>
  __device__ T bcast(T v, int src) const { return __shfl_sync(FULL, v, src, G); }
  __device__ unsigned ballot(bool p) const {
    unsigned b = __ballot_sync(FULL, p);
    return (G == 32) ? b : ((b >> gbase) & ((1u << G) - 1u));
  }
};

// numpy float32 add.reduce over n values held one per lane (n may differ between the groups of a
// warp): pairwise summation with an 8-way unrolled block (n >= 8) or a plain loop from 0.f (n < 8)
// — numpy/_core/src/umath/loops_utils.h.src.  Trip counts are compile-time (G), adds are predicated.
template <int G>
__device__ float np_sum_f32(const Group<G>& g, float a, int n) {
  float seq = 0.f;
#pragma unroll
  for (int i = 0; i < (G < 7 ? G : 7); ++i) {
    const float v = g.bcast(a, i);
    if (i < n) seq = __fadd_rn(seq, v);
  }
  if (G < 8) return seq;
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = g.bcast(a, j & (G - 1));
  const int nb = n - (n % 8);
#pragma unroll
  for (int i = 8; i + 8 <= G; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = g.bcast(a, (i + j) & (G - 1));
      if (i < nb) r[j] = __fadd_rn(r[j], v);
    }
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const float v = g.bcast(a, (nb + t) & (G - 1));
    if (nb + t < n) res = __fadd_rn(res, v);
  }
  return n < 8 ? seq : res;
}

// cdf of RandomState.choice: float64 cumsum of p (sequential), divided by the last entry.
// Every lane i < n returns cdf[i]; lanes >= n return 2.0 (never <= u).  Lanes >= n must pass p = 0.
template <int G>
__device__ double choice_cdf(const Group<G>& g, double p, int n) {
  double acc = 0.0, mine = 0.0;
#pragma unroll
  for (int i = 0; i < G; ++i) {
    acc = __dadd_rn(acc, g.bcast(p, i));     // + 0.0 beyond n leaves acc unchanged
    if (g.gl == i) mine = acc;
  }
  return (g.gl < n) ? __ddiv_rn(mine, acc) : 2.0;
}

// (policy + 1e-12) / sum in float32 (mcts.py:205-206, :291-292); lanes >= n hold 0.
template <int G>
__device__ float normalise_policy(const Group<G>& g, float pol, int n) {
  const float p = (g.gl < n) ? __fadd_rn(pol, 1e-12f) : 0.f;
  const float s = np_sum_f32(g, p, n);
  return (g.gl < n) ? __fdiv_rn(p, s) : 0.f;
}

// np.random.choice(n, bound, p=p, replace=False): returns the bit set of chosen indices and advances
// the tree's uniform cursor by the number of draws numpy would have consumed.
template <int G>
__device__ unsigned choice_without_replacement(const Group<G>& g, const SmzArena& a, const SmzRng& rng, bool alive, int tree,
                                               float p32, int n, int bound, int& cursor) {
  unsigned found = 0;
  int nf = alive ? 0 : bound;
  const double pd = (g.gl < n) ? (double)p32 : 0.0;
  for (int round = 0; __any_sync(FULL, nf < bound); ++round) {
    if (round > 2 * SMZ_MAX_POLICY) {   // NaN / degenerate policy: numpy would raise; do not hang
      if (nf < bound) {
        *a.error_flag = 2;
        for (int i = 0; i < n && nf < bound; ++i)
          if (!((found >> i) & 1u)) { found |= 1u << i; ++nf; }
      }
      break;
    }
    const int m = bound - nf;                    // 0 for groups that are done
    const double c = choice_cdf(g, ((found >> g.gl) & 1u) ? 0.0 : pd, n);
    for (int j = 0; __any_sync(FULL, j < m); ++j) {
      const bool on = j < m;
      const double u = on ? smz_rng_uniform(rng, tree, cursor + j) : 0.0;
      int idx = __popc(g.ballot(c <= u));
      idx = idx < n ? idx : n - 1;
      if (on && !((found >> idx) & 1u)) { found |= 1u << idx; ++nf; }
    }
    cursor += m;
  }
  return found;
}

// ------------------------------------------------------------------------------------------------
template <int G>
__global__ void k_root_expand(SmzArena a, int n_trees, const float* __restrict__ policy, int pstride,
                              const int* __restrict__ root_to_play, int train,
                              const double* __restrict__ dirichlet) {
  Group<G> g;
  int tree = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool alive = tree < n_trees;
  if (!alive) tree = n_trees - 1;
  const size_t tb = (size_t)tree * a.M;
  const int n = a.A;
  int cursor = 0;
  const float pol = (g.gl < n) ? policy[(size_t)tree * pstride + g.gl] : 0.f;
  const float p = normalise_policy(g, pol, n);
  // all A actions become children (ascending); the call still consumes its draws (T5)
  const SmzRng rng = smz_make_rng(a);
  choice_without_replacement(g, a, rng, alive, tree, p, n, n, cursor);
  if (!alive) return;
  if (g.gl < n) {
    double prior = (double)p;
    if (train) {
      const double nz = dirichlet[(size_t)tree * a.A + g.gl];
      prior = __dadd_rn((double)__fmul_rn(p, a.one_minus_frac_f32), __dmul_rn(nz, a.frac));
    }
    a.root_prior[(size_t)tree * a.A + g.gl] = prior;
    a.stat[tb + 1 + g.gl] = make_int4(0, 0, 0, __float_as_int(p));
    a.link[tb + 1 + g.gl] = make_int2(0, g.gl);
  }
  if (g.gl == 0) {
    a.stat[tb] = make_int4(0, 0, 0, 0);
    a.link[tb] = make_int2(1, -1);
    a.minmax[tree] = make_float2(__int_as_float(0x7f800000), __int_as_float(0xff800000));
    a.ucursor[tree] = cursor;
    a.root_to_play[tree] = root_to_play ? root_to_play[tree] : 0;
    a.path_len[tree] = 0;
    if (a.rec_root_policy)
      for (int i = 0; i < a.W; ++i) a.rec_root_policy[(size_t)tree * a.W + i] = i < n ? policy[(size_t)tree * pstride + i] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// Per-tree state that flows from the backup of one simulation into the descent of the next.  The fused
// kernel keeps it in registers; the stand-alone kernels load / store it.
struct TreeState {
  int cursor;      // uniform draws consumed so far
  float2 mm;       // MinMaxStats
  int2 root;       // root {visit_count, value_sum bits}
};

// The search path is kept as one int4 record per level {node, visit_count, value_sum, reward} captured
// while descending, so the backup needs ONE round trip (the record) instead of two (index -> stat).
template <int G>
__device__ __forceinline__ void select_phase(const Group<G>& g, const SmzArena& a, const SmzRng& rng, int tree, bool alive,
                                             int sim, TreeState ts, int* __restrict__ o_slot, int* __restrict__ o_action,
                                             int* __restrict__ o_branch) {
  const size_t tb = (size_t)tree * a.M;
  int cursor = ts.cursor;
  const float vmin = ts.mm.x, vmax = ts.mm.y;
  int4* path = a.path + (size_t)tree * a.path_stride;

  int depth = 0, cbase = 1, nch = a.A;
  int parent_visit = ts.root.x;
  if (alive && g.gl == 0) path[0] = make_int4(0, ts.root.x, ts.root.y, 0);
  int L = 1, child = 0, child_key = 0;
  bool going = alive;
  while (__any_sync(FULL, going)) {
    const bool act = going && g.gl < nch;
    const bool chance = (depth >> 1) & 1;
    int4 st = make_int4(0, 0, 0, 0);
    int2 lk = make_int2(0, 0);
    double prior0 = 0.0, pbc = 0.0;
    if (act) {
      st = a.stat[tb + cbase + g.gl];
      lk = a.link[tb + cbase + g.gl];
      if (depth == 0) prior0 = a.root_prior[(size_t)tree * a.A + g.gl];
      if (!chance) pbc = __ldg(a.pbc + parent_visit);
    }
    // the draw only depends on the cursor: it is computed while the loads are in flight
    const double u = (going && (chance || act)) ? smz_rng_uniform(rng, tree, chance ? cursor : cursor + g.gl) : 0.0;
    int pick = 0;
    if (__any_sync(FULL, going && chance)) {
      // chance node: sample a child from the smoothed priors (mcts.py:249-255, T9)
      const float p = __int_as_float(st.w);
      const float om = act ? __fadd_rn(__fsub_rn(1.f, p), 1e-12f) : 0.f;
      const float rem = fabsf(__fdiv_rn(np_sum_f32(g, om, nch), (float)nch));
      const float sh = act ? __fadd_rn(p, rem) : 0.f;
      const float tot = np_sum_f32(g, sh, nch);
      const float q = act ? __fdiv_rn(sh, tot) : 0.f;
      const double c = choice_cdf(g, (double)q, nch);
      int pk = __popc(g.ballot(act && c <= u));
      pk = pk < nch ? pk : nch - 1;
      if (chance) { pick = pk; cursor += going ? 1 : 0; }
    }
    if (__any_sync(FULL, going && !chance)) {
      // decision node: argmax of ucb_score, one fresh uniform per child (mcts.py:235-243, T3/T5/T6)
      double score = -__longlong_as_double(0x7ff0000000000000LL);
      int best = -1;
      if (act && !chance) {
        const double prior = (depth == 0) ? prior0 : (double)__int_as_float(st.w);
        // pbc = sqrt(n) * pb_c(n), the left-associated head of mcts.py:237, tabulated by the host
        const double ps = __ddiv_rn(__dmul_rn(pbc, prior), (double)(st.x + 1));
        double vs = 0.0;
        if (st.x > 0) {
          const float val = __fdiv_rn(__int_as_float(st.y), (float)st.x);
          float v = __fadd_rn(__int_as_float(st.z), __fmul_rn(a.discount, val));
          if (vmax > vmin) v = __fdiv_rn(__fsub_rn(v, vmin), __fsub_rn(vmax, vmin));
          vs = (double)v;
        }
        const double noise = __dadd_rn(1e-7, __dmul_rn(2e-7 - 1e-7, u));
        score = __dadd_rn(__dadd_rn(ps, vs), noise);
        best = g.gl;
      }
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) {
        const double os = __shfl_xor_sync(FULL, score, off, G);
        const int ob = __shfl_xor_sync(FULL, best, off, G);
        if (os > score || (os == score && ob > best)) { score = os; best = ob; }
      }
      if (!chance) { pick = best < 0 ? 0 : best; cursor += going ? nch : 0; }
    }
    const int4 cst = make_int4(g.bcast(st.x, pick), g.bcast(st.y, pick), g.bcast(st.z, pick), 0);
    const int child_cb = g.bcast(lk.x, pick);
    const int key = g.bcast(lk.y, pick);
    if (going) {
      child = cbase + pick;
      child_key = key;
      if (g.gl == 0) path[L] = make_int4(child, cst.x, cst.y, cst.z);
      ++L;
      if (child_cb == 0 || L >= a.path_stride) {
        going = false;
      } else {
        // the child (depth+1) was expanded through the dynamics pair iff this node is a chance node (T2)
        nch = chance ? a.Kd : a.Kc;
        parent_visit = cst.x;
        cbase = child_cb;
        ++depth;
      }
    }
  }
  if (alive && g.gl == 0) {
    const int branch = ((depth >> 1) & 1) ? SMZ_BRANCH_DYNAMICS : SMZ_BRANCH_AFTERSTATE;
    const int slot = (cbase == 1) ? 0 : (cbase - 1 - a.A) / a.Kmax + 1;
    a.leaf_node[tree] = child;
    a.leaf_slot[tree] = slot;
    a.leaf_action[tree] = child_key;
    a.leaf_branch[tree] = branch;
    a.path_len[tree] = L;
    a.ucursor[tree] = cursor;
    if (o_slot) o_slot[tree] = slot;
    if (o_action) o_action[tree] = child_key;
    if (o_branch) o_branch[tree] = branch;
    const int r = atomicAdd(&a.branch_count[sim * 2 + branch], 1);
    a.rows[(size_t)branch * a.B + r] = tree;
    a.rows4[(size_t)branch * a.B + r] = make_int4(tree, slot, child_key, 0);
    atomicAdd(a.depth_sum, (unsigned long long)(depth + 1));
  }
}

// ------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ TreeState expand_backup_phase(const Group<G>& g, const SmzArena& a, const SmzRng& rng, int tree,
                                                         bool alive, int sim, const float* __restrict__ policy,
                                                         int pstride, const float* __restrict__ value,
                                                         const float* __restrict__ reward) {
  const size_t tb = (size_t)tree * a.M;
  // everything this phase needs from memory is requested up front (one round trip): the leaf record,
  // the network outputs, the tree state and — speculatively — the first two chunks of path records
  const int4* path = a.path + (size_t)tree * a.path_stride;
  const int4 rec0 = (g.gl < a.path_stride) ? path[g.gl] : make_int4(0, 0, 0, 0);
  const int4 rec1 = (G + g.gl < a.path_stride) ? path[G + g.gl] : make_int4(0, 0, 0, 0);
  const int leaf = a.leaf_node[tree];
  const int branch = a.leaf_branch[tree];
  const int L = alive ? a.path_len[tree] : 0;
  int cursor = a.ucursor[tree];
  float2 mm = a.minmax[tree];
  float v = value[tree];
  const float rew_in = reward[tree];
  const signed char* sign = a.sign + (size_t)a.root_to_play[tree] * (a.N + 2);

  // children of the leaf (mcts.py:289-297): width C after the afterstate pair, A after dynamics
  const int n = branch ? a.A : a.C;
  const int bound = min(a.K, n);
  const float pol = (g.gl < n) ? policy[(size_t)tree * pstride + g.gl] : 0.f;
  const float p = normalise_policy(g, pol, n);
  const unsigned found = choice_without_replacement(g, a, rng, alive, tree, p, n, bound, cursor);
  const int cb = 1 + a.A + sim * a.Kmax;
  if (alive && g.gl < n && ((found >> g.gl) & 1u)) {
    const int r = __popc(found & ((1u << g.gl) - 1u));
    a.stat[tb + cb + r] = make_int4(0, 0, 0, __float_as_int(p));
    a.link[tb + cb + r] = make_int2(0, g.gl);
  }
  const float rew = branch ? rew_in : 0.f;
  if (alive && g.gl == 0) {
    a.link[tb + leaf].x = cb;
    a.ucursor[tree] = cursor;
    if (branch) reinterpret_cast<int*>(a.stat + tb + leaf)[2] = __float_as_int(rew);     // Node.reward of the leaf
    if (a.rec_policy) {
      const size_t ro = ((size_t)tree * a.N + sim);
      for (int i = 0; i < a.W; ++i) a.rec_policy[ro * a.W + i] = i < n ? policy[(size_t)tree * pstride + i] : 0.f;
      a.rec_value[ro] = v;
      a.rec_reward[ro] = rew;
      a.rec_branch[ro] = (signed char)branch;
    }
  }

  // backup leaf -> root (mcts.py:299-308): lanes own path levels, the discounted return is a serial
  // float32 recurrence (mul then add, two roundings) carried through shuffles
  int2 root = make_int2(0, 0);
  int n_chunks = (L + G - 1) / G;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n_chunks = max(n_chunks, __shfl_xor_sync(FULL, n_chunks, off));
  for (int chunk = n_chunks - 1; chunk >= 0; --chunk) {
    const int l = chunk * G + g.gl;
    const bool valid = l < L;
    int4 rec = chunk == 0 ? rec0 : rec1;
    if (chunk > 1 && valid) rec = path[l];
    if (!valid) rec = make_int4(0, 0, 0, 0);
    if (valid && l == L - 1 && branch) rec.w = __float_as_int(rew);
    const float r = __int_as_float(rec.w);
    float myv = 0.f;
#pragma unroll
    for (int t = G - 1; t >= 0; --t) {
      const float ri = g.bcast(r, t);
      if (chunk * G + t < L) {
        if (g.gl == t) myv = v;
        v = __fadd_rn(ri, __fmul_rn(a.discount, v));
      }
    }
    if (valid) {
      const float vs = __fadd_rn(__int_as_float(rec.z), (signed char)__ldg(sign + l) > 0 ? myv : -myv);
      const int2 upd = make_int2(rec.y + 1, __float_as_int(vs));
      *reinterpret_cast<int2*>(a.stat + tb + rec.x) = upd;          // {visit_count, value_sum}
      if (l == 0) root = upd;
      const float nv = __fdiv_rn(vs, (float)upd.x);
      mm.x = fminf(mm.x, nv);
      mm.y = fmaxf(mm.y, nv);
    }
  }
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
    mm.x = fminf(mm.x, __shfl_xor_sync(FULL, mm.x, off, G));
    mm.y = fmaxf(mm.y, __shfl_xor_sync(FULL, mm.y, off, G));
  }
  if (alive && g.gl == 0) a.minmax[tree] = mm;
  TreeState ts;
  ts.cursor = cursor;
  ts.mm = mm;
  ts.root = make_int2(g.bcast(root.x, 0), g.bcast(root.y, 0));   // level 0 lives in lane 0 of chunk 0
  return ts;
}

template <int G>
__global__ void k_select(SmzArena a, int n_trees, int sim, int* __restrict__ o_slot, int* __restrict__ o_action,
                         int* __restrict__ o_branch) {
  Group<G> g;
  int tree = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool alive = tree < n_trees;
  if (!alive) tree = n_trees - 1;
  const SmzRng rng = smz_make_rng(a);
  TreeState ts;
  ts.cursor = a.ucursor[tree];
  ts.mm = a.minmax[tree];
  const int4 rs = a.stat[(size_t)tree * a.M];
  ts.root = make_int2(rs.x, rs.y);
  select_phase(g, a, rng, tree, alive, sim, ts, o_slot, o_action, o_branch);
}

template <int G>
__global__ void k_expand_backup(SmzArena a, int n_trees, int sim, const float* __restrict__ policy, int pstride,
                                const float* __restrict__ value, const float* __restrict__ reward) {
  Group<G> g;
  int tree = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool alive = tree < n_trees;
  if (!alive) tree = n_trees - 1;
  const SmzRng rng = smz_make_rng(a);
  expand_backup_phase(g, a, rng, tree, alive, sim, policy, pstride, value, reward);
}

// expansion + backup of simulation `sim` followed by the descent of simulation `sim + 1` for the same
// tree by the same lanes: one launch per simulation instead of two, the path nodes just updated are
// re-read from L1/L2.
template <int G>
__global__ void k_backup_select(SmzArena a, int n_trees, int sim) {
  smz_pdl_wait();                 // the network step of `sim` must have landed
  smz_pdl_launch_dependents();    // the next network step may set itself up (barriers, TMEM, weights)
  Group<G> g;
  int tree = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool alive = tree < n_trees;
  if (!alive) tree = n_trees - 1;
  const SmzRng rng = smz_make_rng(a);
  const TreeState ts = expand_backup_phase(g, a, rng, tree, alive, sim, a.out_policy, a.W, a.out_value, a.out_reward);
  __syncwarp();
  select_phase(g, a, rng, tree, alive, sim + 1, ts, nullptr, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------------
__global__ void k_read_roots(SmzArena a, int n_trees, int* __restrict__ visits, float* __restrict__ values,
                             double* __restrict__ priors, float* __restrict__ rewards) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_trees * a.A) return;
  const int tree = i / a.A, c = i % a.A;
  const size_t tb = (size_t)tree * a.M;
  const int4 st = a.stat[tb + 1 + c];
  if (visits) visits[i] = st.x;
  if (rewards) rewards[i] = __int_as_float(st.z);
  if (priors) priors[i] = a.root_prior[i];
  if (values && c == 0) {
    const int4 r = a.stat[tb];
    values[tree] = r.x == 0 ? 0.f : __fdiv_rn(__int_as_float(r.y), (float)r.x);
  }
}

// numpy float64 add.reduce order over a small local array (same pairwise scheme as np_sum_f32)
__device__ double np_sum_f64(const double* x, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, x[i]);
    return r;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = x[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], x[i + j]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, x[i]);
  return res;
}

// The step after the search, one thread per tree (game.py:179-216): the policy stored as training target
// (`store_search_statistics`), the temperature-scaled acting policy (`policy_action_reward_from_tree` +
// `softmax_stable`) and the action (`select_action`: sample when T > 0.1 or the policy is flat, else argmax).
__global__ void k_select_actions(SmzArena a, int n_trees, double temperature, const double* __restrict__ u_in,
                                 int* __restrict__ action_out, double* __restrict__ policy_out,
                                 double* __restrict__ stored_out) {
  const int tree = blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= n_trees) return;
  const size_t tb = (size_t)tree * a.M;
  const int n = a.A;
  double v[SMZ_MAX_POLICY], pr[SMZ_MAX_POLICY], p[SMZ_MAX_POLICY];
  for (int i = 0; i < n; ++i) {
    v[i] = (double)a.stat[tb + 1 + i].x;
    pr[i] = a.root_prior[(size_t)tree * n + i];
  }
  const double sv = np_sum_f64(v, n);
  if (stored_out) {
    const double sp = np_sum_f64(pr, n);
    for (int i = 0; i < n; ++i) stored_out[(size_t)tree * n + i] = sv >= 3.0 ? __ddiv_rn(v[i], sv) : __ddiv_rn(pr[i], sp);
  }
  const double e = 1.0 / temperature;
  for (int i = 0; i < n; ++i) {
    double x = sv <= 1.0 ? pr[i] : v[i];
    if (temperature >= 0.3) x = e == 1.0 ? x : (e == 2.0 ? __dmul_rn(x, x) : (e == 0.5 ? sqrt(x) : pow(x, e)));
    p[i] = x;
  }
  const double s = np_sum_f64(p, n);
  bool flat = true;
  for (int i = 0; i < n; ++i) {
    p[i] = __ddiv_rn(p[i], s);
    flat = flat && p[i] == p[0];
    if (policy_out) policy_out[(size_t)tree * n + i] = p[i];
  }
  int pick = 0;
  if (temperature > 0.1 || flat) {
    const double u = u_in ? u_in[tree] : smz_philox_uniform(a.seed_state[0], a.seed_state[1] + (unsigned long long)tree, 0u, 2u);
    double acc = 0.0;
    for (int i = 0; i < n; ++i) { acc = __dadd_rn(acc, p[i]); v[i] = acc; }
    for (int i = 0; i < n; ++i) pick += __ddiv_rn(v[i], acc) <= u;
    pick = pick < n ? pick : n - 1;
  } else {
    for (int i = 1; i < n; ++i) if (p[i] > p[pick]) pick = i;
  }
  if (action_out) action_out[tree] = a.link[tb + 1 + pick].y;
}

// Dirichlet(alpha) per tree on the device: Gamma(alpha) by Marsaglia-Tsang (alpha+1 boost), Philox
// stream 1.  Production mode only; parity runs pass recorded noise (np.random.dirichlet, mcts.py:220).
__global__ void k_dirichlet(SmzArena a, int n_trees) {
  const int tree = blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= n_trees) return;
  unsigned ctr = 0;
  const unsigned long long seed = a.seed_state[0], tid = a.seed_state[1] + (unsigned long long)tree;
  auto U = [&]() { double u = smz_philox_uniform(seed, tid, ctr++, 1u); return u > 0.0 ? u : 1.0 / 9007199254740992.0; };
  double sum = 0.0;
  for (int i = 0; i < a.A; ++i) {
    const double al = a.alpha;
    const double d = (al < 1.0 ? al + 1.0 : al) - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double gsample = d;
    for (int it = 0; it < 256; ++it) {
      const double u1 = U(), u2 = U();
      const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
      const double t = 1.0 + c * x;
      if (t <= 0.0) continue;
      const double v3 = t * t * t, u = U();
      if (log(u) < 0.5 * x * x + d - d * v3 + d * log(v3)) { gsample = d * v3; break; }
    }
    if (al < 1.0) gsample *= pow(U(), 1.0 / al);
    a.dirichlet[(size_t)tree * a.A + i] = gsample;
    sum += gsample;
  }
  for (int i = 0; i < a.A; ++i) a.dirichlet[(size_t)tree * a.A + i] /= (sum > 0.0 ? sum : 1.0);
}

template <int G>
int grid_for(int n_trees, int threads) { return (int)(((long long)n_trees * G + threads - 1) / threads); }

}  // namespace

#define SMZ_DISPATCH_G(G_, CALL)        \
  switch (G_) {                         \
    case 2: { constexpr int G = 2; CALL; } break;   \
    case 4: { constexpr int G = 4; CALL; } break;   \
    case 8: { constexpr int G = 8; CALL; } break;   \
    case 16: { constexpr int G = 16; CALL; } break; \
    default: { constexpr int G = 32; CALL; } break; \
  }

static const int kThreads = 128;

void smz_launch_root_expand(const SmzArena& a, int lanes, int n_trees, const float* policy, int pstride,
                            const int* root_to_play, int train, const double* dirichlet, cudaStream_t s) {
  SMZ_DISPATCH_G(lanes, (k_root_expand<G><<<grid_for<G>(n_trees, kThreads), kThreads, 0, s>>>(
                            a, n_trees, policy, pstride, root_to_play, train, dirichlet)));
}

void smz_launch_select(const SmzArena& a, int lanes, int n_trees, int sim, int* o_slot, int* o_action, int* o_branch,
                       cudaStream_t s) {
  SMZ_DISPATCH_G(lanes, (k_select<G><<<grid_for<G>(n_trees, kThreads), kThreads, 0, s>>>(a, n_trees, sim, o_slot,
                                                                                         o_action, o_branch)));
}

void smz_launch_expand_backup(const SmzArena& a, int lanes, int n_trees, int sim, const float* policy, int pstride,
                              const float* value, const float* reward, cudaStream_t s) {
  SMZ_DISPATCH_G(lanes, (k_expand_backup<G><<<grid_for<G>(n_trees, kThreads), kThreads, 0, s>>>(
                            a, n_trees, sim, policy, pstride, value, reward)));
}

void smz_launch_backup_select(const SmzArena& a, int lanes, int n_trees, int sim, bool pdl, cudaStream_t s) {
  SMZ_DISPATCH_G(lanes, (smz_launch(k_backup_select<G>, dim3(grid_for<G>(n_trees, kThreads)), dim3(kThreads), 0, s, pdl,
                                    a, n_trees, sim)));
}

void smz_launch_read_roots(const SmzArena& a, int n_trees, int* visits, float* values, double* priors, float* rewards,
                           cudaStream_t s) {
  const int n = n_trees * a.A;
  k_read_roots<<<(n + 255) / 256, 256, 0, s>>>(a, n_trees, visits, values, priors, rewards);
}

void smz_launch_select_actions(const SmzArena& a, int n_trees, double temperature, const double* u, int* actions,
                               double* policy, double* stored, cudaStream_t s) {
  k_select_actions<<<(n_trees + 127) / 128, 128, 0, s>>>(a, n_trees, temperature, u, actions, policy, stored);
}

void smz_launch_dirichlet(const SmzArena& a, int n_trees, cudaStream_t s) {
  k_dirichlet<<<(n_trees + 127) / 128, 128, 0, s>>>(a, n_trees);
}
