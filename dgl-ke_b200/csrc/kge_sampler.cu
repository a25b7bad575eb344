// kge_sampler.cu -- device-side edge / negative sampler (replaces DGL's C++ EdgeSampler on the training path:
// dataloader/sampler.py:376-419 create_sampler, :459-512 chunk layout, :823-876 head/tail alternation).
//
// Semantics kept from the reference's use of EdgeSampler (exclude_positive=False, shuffle=True, return_false_neg=False):
//   * positives: the partition's edges in a fresh random order every epoch, batch_size per step, the ragged tail of an
//     epoch dropped (sampler.py:503-504);
//   * negatives: num_chunks * neg_sample_size entity ids drawn uniformly WITH replacement from all entities;
//   * step k corrupts tails for even k, heads for odd k (NewBidirectionalOneShotIterator starts with the tail sampler);
//   * the positive graph's node list = the distinct head/tail ids of the batch (here: in order of first appearance in
//     [heads | tails]) and the edges' endpoints as indices into it.
//
// Everything is counter based (no RNG state): the permutation of an epoch is a 4-round Feistel network over
// 2^(2*hb) >= n_edges with cycle walking, the negatives are splitmix64 hashes of (seed, step, j).  dglke_b200/sampler.py
// restates the same integer arithmetic in numpy: the two produce bit-identical index arrays (tests/test_sampler.py).
#include <cstdio>
#include "kge_common.cuh"

namespace kge {

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ unsigned long long feistel_perm(unsigned long long x, unsigned long long n, int hb,
                                                           unsigned long long key) {
  const unsigned long long mask = (1ull << hb) - 1ull;
  do {
    unsigned long long L = x >> hb, R = x & mask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const unsigned long long f = mix64(R ^ (key + (unsigned long long)r * 0xD1B54A32D192ED03ull)) & mask;
      const unsigned long long nl = R;
      R = L ^ f;
      L = nl;
    }
    x = (L << hb) | R;
  } while (x >= n);                          // cycle walking keeps the map a bijection of [0, n)
  return x;
}

__device__ __forceinline__ int table_insert(const SamplerParams& p, unsigned long long key, int pos) {
  int slot = (int)(mix64(key) & (unsigned long long)p.hmask);
  while (true) {
    const unsigned long long prev = atomicCAS(p.tkey + slot, ~0ull, key);
    if (prev == ~0ull || prev == key) { atomicMin(p.tpos + slot, pos); return slot; }
    slot = (slot + 1) & p.hmask;
  }
}
__device__ __forceinline__ int table_find(const SamplerParams& p, unsigned long long key) {
  int slot = (int)(mix64(key) & (unsigned long long)p.hmask);
  while (p.tkey[slot] != key) slot = (slot + 1) & p.hmask;
  return slot;
}

// positives of step `step` (epoch permutation) + negatives + hash-table insert of the endpoints
__global__ void __launch_bounds__(256) k_sample_draw(SamplerParams p, long long step) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.B) {
    const long long per_epoch = p.n_edges / p.B;
    const long long epoch = step / per_epoch, j = step % per_epoch;
    const unsigned long long key = mix64(p.seed ^ (0xA0761D6478BD642Full * (unsigned long long)(epoch + 1)));
    const long long e = (long long)feistel_perm((unsigned long long)(j * p.B + i), (unsigned long long)p.n_edges, p.half_bits, key);
    const long long h = p.heads[e], t = p.tails[e];
    p.o_head[i] = h; p.o_rel[i] = p.rels[e]; p.o_tail[i] = t;
    table_insert(p, (unsigned long long)h, (int)i);
    table_insert(p, (unsigned long long)t, (int)(p.B + i));
  } else if (i < p.B + p.Nn) {
    const long long j = i - p.B;
    const unsigned long long r = mix64(mix64(p.seed + 0x632BE59BD9B4E019ull * (unsigned long long)(step + 1)) + (unsigned long long)j);
    p.o_neg[j] = (long long)(r % (unsigned long long)p.n_entities);
  }
}

// one CTA: first-occurrence flags -> exclusive scan -> node list in order of first appearance
__global__ void __launch_bounds__(1024) k_sample_unique(SamplerParams p) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int n = (int)(2 * p.B);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int pos = base + tid;
    int flag = 0, slot = 0;
    unsigned long long key = 0;
    if (pos < n) {
      key = (unsigned long long)(pos < p.B ? p.o_head[pos] : p.o_tail[pos - p.B]);
      slot = table_find(p, key);
      flag = (p.tpos[slot] == pos) ? 1 : 0;
    }
    int v = flag;                                   // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    if (wid == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_tot[lane] = w;                           // inclusive totals of the warps
    }
    __syncthreads();
    const int excl = carry + (wid ? warp_tot[wid - 1] : 0) + v - flag;
    if (flag) { p.o_nodes[excl] = (long long)key; p.tloc[slot] = excl; }
    __syncthreads();
    if (tid == 1023) carry += warp_tot[31];
    __syncthreads();
  }
  if (tid == 0) *p.o_n_nodes = carry;
}

// endpoints -> local ids; the table slots are reset for the next step on the way out
__global__ void __launch_bounds__(256) k_sample_local(SamplerParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B) return;
  const int sh = table_find(p, (unsigned long long)p.o_head[i]);
  const int st = table_find(p, (unsigned long long)p.o_tail[i]);
  p.o_hl[i] = p.tloc[sh];
  p.o_tl[i] = p.tloc[st];
}
__global__ void __launch_bounds__(256) k_sample_reset(SamplerParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= p.hmask) { p.tkey[i] = ~0ull; p.tpos[i] = 0x7fffffff; }
}

void launch_sampler(const LaunchCtx& c, const SamplerParams& p, long long step) {
  KGE_LAUNCH(c, k_sample_draw, ceil_div(p.B + p.Nn, 256), 256, 0, p, step);
  KGE_LAUNCH(c, k_sample_unique, 1, 1024, 0, p);
  KGE_LAUNCH(c, k_sample_local, ceil_div(p.B, 256), 256, 0, p);
  KGE_LAUNCH(c, k_sample_reset, ceil_div((long long)p.hmask + 1, 256), 256, 0, p);
}

}  // namespace kge
