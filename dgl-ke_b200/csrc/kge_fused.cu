// kge_fused.cu -- the fused contraction kernel of the step (tcgen05 / TMEM / TMA, sm_100a):
//
//   S = X . Y^T  (tensor cores, accumulator in TMEM)  ->  loss / self-adversarial softmax / backward
//   coefficients V computed from TMEM in registers  ->  V (TF32 hi/lo) written back to TMEM
//   ->  G = V . Y  (tensor cores, A operand read from TMEM)  ->  epilogue.
//
// The negative-score matrix S and the coefficient matrix V never leave the SM.  The kernel runs twice per step:
//
//   mode P  lanes = positives i, columns = negatives j:   S = A.Bn^T, row softmax (thread-local), GA = V.Bn
//           replaces create_neg (score_fun.py:91-108,268-286,345-376,427-449), LossGenerator.get_total_loss
//           (loss.py:69-98) and the dL/da half of loss.backward()
//   mode N  lanes = negatives j, columns = positives i:   S^T = Bn.A^T, V^T from the row statistics mode P left
//           behind, G_neg = V^T.A - colsum*b + reg'(b), mean(G_neg^2)  (the dL/db half of loss.backward() plus
//           phase 1 of ExternalEmbedding.update for the negatives, tensor_models.py:316-328)
//
// Recomputing S in the second orientation costs one extra tensor-core GEMM per chunk and removes every HBM/L2 round
// trip of S and V (and the k_loss / k_colsum / k_state_add launches).  fp32 fidelity: operands are TF32 hi/lo pairs
// and every k-step issues hi*hi + hi*lo + lo*hi (3xTF32, fp32 accumulation in TMEM).
//
// CTA = 384 threads: warps 0-7 epilogue (two warps per TMEM lane quarter, splitting the columns), warp 8 TMA producer,
// warp 9 TMEM allocator + tcgen05.mma issuer (one elected lane), warps 10-11 idle (they only give their registers away).  Persistent over (chunk, 128-row tile) work
// items.  TMEM columns: [0,N1) S -> V_hi | [N1,2N1) distances -> V_lo | [2N1, 2N1+Wc) accumulator of GEMM2, which is
// processed in Wc-wide column chunks of the output (Wc = 96 at Ns = 200).  Shared memory: one 216 KB ring used as
// nS1 stages {X_hi,X_lo,Y_hi,Y_lo} by GEMM1 and as nS2 stages {Y_hi,Y_lo} (MN-major) by GEMM2.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "kge_common.cuh"
#include "kge_tc.cuh"

namespace kge {

using namespace tc;

namespace {

constexpr int kTileM = 128;
constexpr int kThreadsF = 384;                    // warps 0-7 epilogue (two warpgroups), 8 TMA producer, 9 MMA issuer, 10-11 idle
constexpr int kProducerWarp = 8, kMmaWarp = 9;
constexpr int kMaxS1 = 4, kMaxS2 = 8;
constexpr int kMaxPf = 8;                          // row slots per prefetch warp
constexpr uint32_t kRingBytes = 192 * 1024;
constexpr int kStagePitch = 20;                      // floats per row of an epilogue staging tile (16 + 4 pad, 16-byte aligned rows)
constexpr uint32_t kEpiStageBytes = 8 * 32 * kStagePitch * 4;   // one 32 x 16 tile per epilogue warp
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

enum { F_P = 0, F_N = 1 };

struct FusedArgs {
  int model, adversarial;
  float gamma, Tl2e, inv2B, uni;
  float reg_coef;
  int reg_norm;
  int C, Rx, Ry, D;      // rows per chunk on the lane side / on the column side, row length
  int N1;                // Ry rounded up to 16 (UMMA N of GEMM1, TMEM region width)
  int nblkD;             // 32-column slab blocks of D
  int Wc;                // GEMM2 output-column chunk (multiple of 32)
  int nS1, nS2;
  uint32_t stage1Bytes, stage2Bytes;
  const float* x2;       // |x|^2 per lane-side row   (TransE_l2)
  const float* y2;       // |y|^2 per column-side row (TransE_l2)
  // mode P
  const float* pos;      // [B] positive scores
  const float* wt;       // [B] edge weights or null
  const float* wbar;     // [1] mean edge weight (with wt)
  float *gpos, *rowsum, *pl, *nl, *stat_m, *stat_k;
  float* dumpS;          // optional [C*Rx, Ry]: negative scores (kge_debug_read)
  float* dumpV;          // optional [C*Rx, Ry]: backward coefficients (tests)
  // mode N
  const float* cstat_m;  // per positive: softmax shift (log2 domain)
  const float* cstat_k;  // per positive: w_i / (2B den_i)  (or w_i / (2B Ns))
  const float *Xhi, *Xlo;  // slabs of the lane-side rows (negatives): b = hi + lo in the epilogue
  const long long* xids;   // mode N, one GPU: entity ids of the lane-side rows -- b is then read from the table itself
  TableView xtab;          //   (one fp32 load instead of hi + lo; nothing updates the table before k_update)
  const float* xraw;       // mode N, sharded + staged: the lane-side rows as fp32 [C*Rx, D] (the previous step's prefetch)
  float* gsn;            // [C*Rx] mean(G_neg^2)
  float* out;            // P: GA [C*Rx, D]; N: G_neg [C*Rx, D]
  // next step's rows, copied by the two spare warps while this step's tiles are computed (sharded tables: the remote-row
  // latency of step k+1 hides behind the tensor-core work of step k).  Virtual row v of [nodes | negatives]; this launch
  // takes the v with v % 2 == pf_parity (the P and the N kernel split the list).
  int pf_slots;                    // row slots per warp (0 = no prefetch), carved from the ring behind the GEMM stages
  int pf_parity;
  int pf_lag;                      // stores between a slot's own store and its re-load (>= 1)
  uint32_t pf_off, pf_row_bytes;
  const long long* pf_node_ids;    // next batch's unique nodes
  const long long* pf_nU_dev;      // their count on the device, or null
  long long pf_nU, pf_nNeg;
  const long long* pf_neg_ids;
  float* pf_nc;                    // [nU, D] destination of the node rows (the next step's NC)
  float* pf_bn;                    // [nNeg, D] destination of the negative rows
  unsigned long long* dbg;   // optional per-CTA timestamps of the first tile (KGE_B200_FUSED_TIMING=1)
  int exp_halfload;          // experiment (KGE_B200_FUSED_HALFLOAD): skip the TMA loads of the lo tiles (WRONG results;
                             // shows how much of the GEMM time is operand traffic)
};

// regulariser gradient in the epilogue: the default norm (3) inline, anything else out of line (code size: the
// epilogue is instruction-cache sensitive)
static __device__ __noinline__ float4 reg_grad4_any(float4 b, int norm, float coef) { return reg_grad4(b, norm, coef); }
__device__ __forceinline__ float4 reg_grad4_fast(float4 b, int norm, float coef) {
  if (norm == 3) {
    const float c3 = 3.f * coef;
    return make_float4(c3 * fabsf(b.x) * b.x, c3 * fabsf(b.y) * b.y, c3 * fabsf(b.z) * b.z, c3 * fabsf(b.w) * b.w);
  }
  return reg_grad4_any(b, norm, coef);
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(kThreadsF, 1)
k_fused(const __grid_constant__ CUtensorMap mXh, const __grid_constant__ CUtensorMap mXl,
        const __grid_constant__ CUtensorMap mYh1, const __grid_constant__ CUtensorMap mYl1,
        const __grid_constant__ CUtensorMap mYh2, const __grid_constant__ CUtensorMap mYl2, FusedArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full1[kMaxS1], empty1[kMaxS1], full2[kMaxS2], empty2[kMaxS2];
  __shared__ __align__(8) uint64_t s_full, v_ready, acc_full, acc_empty;
  __shared__ __align__(8) uint64_t pf_full[2][kMaxPf];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float colA[256], colB[256], colC[256];
  __shared__ float xch[4][2][kTileM];
  __shared__ float rowscal[kTileM];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* epi_stage = reinterpret_cast<float*>(ring + kRingBytes);

  const int mtiles = (g.Rx + kTileM - 1) / kTileM;
  const int ntiles = g.C * mtiles;
  const int nkb1 = (g.D + 31) >> 5;
  const int nkb2 = (g.Ry + 31) >> 5;
  const int nchunks = (g.D + g.Wc - 1) / g.Wc;
  const uint32_t yBytes1 = (uint32_t)g.N1 * 128u;
  const uint32_t yBytes2 = (uint32_t)(g.Wc >> 5) * 4096u;
  const uint32_t colR2 = (uint32_t)g.N1, colAcc = 2u * (uint32_t)g.N1;
  const bool l2 = g.model == KGE_TRANSE_L2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxS1; ++s) { mbar_init(&full1[s], 1); mbar_init(&empty1[s], 1); }
    for (int s = 0; s < kMaxS2; ++s) { mbar_init(&full2[s], 1); mbar_init(&empty2[s], 1); }
    mbar_init(&s_full, 1); mbar_init(&v_ready, 8); mbar_init(&acc_full, 1); mbar_init(&acc_empty, 8);
    for (int s = 0; s < kMaxPf; ++s) { mbar_init(&pf_full[0][s], 1); mbar_init(&pf_full[1][s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&mXh); tma_prefetch_desc(&mXl); tma_prefetch_desc(&mYh1);
    tma_prefetch_desc(&mYl1); tma_prefetch_desc(&mYh2); tma_prefetch_desc(&mYl2);
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (g.dbg && threadIdx.x == 0) g.dbg[blockIdx.x * 16 + 0] = gtime();
  (void)0;

  // Roles.  The producer and MMA warps run their loops with all 32 lanes (warp-uniform control flow and operands, so the
  // address / descriptor arithmetic stays on the uniform datapath) and one elected lane issues the TMA / tcgen05 ops.
  // Registers: a 12-warp CTA is capped at 168 per thread; the third warpgroup (producer, MMA issuer, two idle warps)
  // hands most of its share to the two epilogue warpgroups, which keep whole accumulator chunks and two chunks' worth of
  // prefetched rows in registers.
  if (warp >= 8) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == kProducerWarp) {
    // ================================ TMA producer ================================
    uint32_t n1 = 0, n2 = 0;      // stage fills issued so far (GEMM1 / GEMM2)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int c = tile / mtiles, m0 = (tile % mtiles) * kTileM;
      // the ring is about to be re-used with the GEMM1 layout: every GEMM2 stage of the previous tile must be consumed
      for (uint32_t k = (n2 > (uint32_t)g.nS2 ? n2 - g.nS2 : 0); k < n2; ++k) mbar_wait(&empty2[k % g.nS2], (k / g.nS2) & 1);
      for (int kb = 0; kb < nkb1; ++kb, ++n1) {
        const uint32_t s = n1 % g.nS1;
        mbar_wait(&empty1[s], ((n1 / g.nS1) & 1) ^ 1);
        uint8_t* st = ring + (size_t)s * g.stage1Bytes;
        // slab layout: TMA row of (chunk c, 32-column block kb, row r) = (c * nblkD + kb) * R + r
        const int yx = (c * g.nblkD + kb) * g.Rx + m0;
        const int yy = (c * g.nblkD + kb) * g.Ry;
        if (elect_one()) {
          if (g.exp_halfload) {
            mbar_expect_tx(&full1[s], 16384u + yBytes1);
            tma_load_2d(st, &mXh, &full1[s], 0, yx);
            tma_load_2d(st + 32768, &mYh1, &full1[s], 0, yy);
          } else {
          mbar_expect_tx(&full1[s], 2u * 16384u + 2u * yBytes1);
          tma_load_2d(st, &mXh, &full1[s], 0, yx);
          tma_load_2d(st + 16384, &mXl, &full1[s], 0, yx);
          tma_load_2d(st + 32768, &mYh1, &full1[s], 0, yy);
          tma_load_2d(st + 32768 + yBytes1, &mYl1, &full1[s], 0, yy);
          }
        }
        __syncwarp();
      }
      // GEMM2 stages overlay the GEMM1 stages: wait until the tensor core has consumed all of them
      for (uint32_t k = (n1 > (uint32_t)g.nS1 ? n1 - g.nS1 : 0); k < n1; ++k) mbar_wait(&empty1[k % g.nS1], (k / g.nS1) & 1);
      for (int ch = 0; ch < nchunks; ++ch) {
        const int d0 = ch * g.Wc;
        int nb = (g.D - d0 + 31) >> 5;
        if (nb > (g.Wc >> 5)) nb = g.Wc >> 5;
        for (int kb = 0; kb < nkb2; ++kb, ++n2) {
          const uint32_t s = n2 % g.nS2;
          mbar_wait(&empty2[s], ((n2 / g.nS2) & 1) ^ 1);
          uint8_t* st = ring + (size_t)s * g.stage2Bytes;
          const int yy0 = (c * g.nblkD + (d0 >> 5)) * g.Ry + kb * 32;
          if (elect_one()) {
            mbar_expect_tx(&full2[s], (g.exp_halfload ? 1u : 2u) * (uint32_t)nb * 4096u);
            for (int b = 0; b < nb; ++b) {
              tma_load_2d(st + b * 4096, &mYh2, &full2[s], 0, yy0 + b * g.Ry);
              if (!g.exp_halfload) tma_load_2d(st + yBytes2 + b * 4096, &mYl2, &full2[s], 0, yy0 + b * g.Ry);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer ================================
    uint32_t n1 = 0, n2 = 0, nacc = 0, it = 0;
    const uint32_t idesc1 = make_idesc(kTileM, g.N1, false, false);
    // descriptor = constant fields | (shared address >> 4): a k-step advances the address field only
    const uint64_t descK = make_desc(0, 16, 1024), descMN = make_desc(0, 4096, 512, 1);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      // ---- GEMM1: S = X . Y^T, K = D ----
      for (int kb = 0; kb < nkb1; ++kb, ++n1) {
        const uint32_t s = n1 % g.nS1;
        mbar_wait(&full1[s], (n1 / g.nS1) & 1);
        if (g.dbg && n1 == 0 && lane == 0) g.dbg[blockIdx.x * 16 + 1] = gtime();
        tc_fence_after();
        const uint32_t st = smem_u32(ring + (size_t)s * g.stage1Bytes);
        const uint64_t dXh = descK | (uint64_t)(st >> 4), dXl = descK | (uint64_t)((st + 16384u) >> 4);
        const uint64_t dYh = descK | (uint64_t)((st + 32768u) >> 4), dYl = descK | (uint64_t)((st + 32768u + yBytes1) >> 4);
        const int kleft = g.D - kb * 32;
        const int ksteps = kleft >= 32 ? 4 : (kleft >> 3);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < ksteps) {
              const uint64_t o = (uint64_t)(ks * 2);     // K-major: +32 bytes per k-step inside the 128-byte swizzle span
              umma_tf32(tmem_base, dXh + o, dYh + o, idesc1, (kb | ks) ? 1u : 0u);
              umma_tf32(tmem_base, dXh + o, dYl + o, idesc1, 1u);
              umma_tf32(tmem_base, dXl + o, dYh + o, idesc1, 1u);
            }
          }
          umma_commit(&empty1[s]);
          if (kb == nkb1 - 1) umma_commit(&s_full);
        }
        __syncwarp();
      }
      // ---- the epilogue warps turn S into V (hi | lo) in TMEM ----
      mbar_wait(&v_ready, it & 1);
      tc_fence_after();
      // ---- GEMM2: G[:, chunk] = V . Y[:, chunk], K = Ry, A operand from TMEM ----
      for (int ch = 0; ch < nchunks; ++ch, ++nacc) {
        const int d0 = ch * g.Wc;
        int nb = (g.D - d0 + 31) >> 5;
        if (nb > (g.Wc >> 5)) nb = g.Wc >> 5;
        const uint32_t idesc2 = make_idesc(kTileM, nb * 32, false, true);
        mbar_wait(&acc_empty, (nacc & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb2; ++kb, ++n2) {
          const uint32_t s = n2 % g.nS2;
          mbar_wait(&full2[s], (n2 / g.nS2) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(ring + (size_t)s * g.stage2Bytes);
          const uint64_t dYh = descMN | (uint64_t)(st >> 4), dYl = descMN | (uint64_t)((st + yBytes2) >> 4);
          const uint32_t aHi = tmem_base + (uint32_t)(kb * 32), aLo = aHi + colR2;
          const int kleft = g.Ry - kb * 32;
          const int ksteps = kleft >= 32 ? 4 : (kleft >> 3);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < ksteps) {
                // MN-major B (128B swizzle, 32B atoms): k-atoms of 4 rows (512 B, SBO), one k-step = 2 atoms = 1024 B,
                // LBO = 4096 between the 32-wide column blocks
                const uint64_t o = (uint64_t)(ks * 64);
                umma_tf32_ts(tmem_base + colAcc, aHi + ks * 8, dYh + o, idesc2, (kb | ks) ? 1u : 0u);
                umma_tf32_ts(tmem_base + colAcc, aHi + ks * 8, dYl + o, idesc2, 1u);
                umma_tf32_ts(tmem_base + colAcc, aLo + ks * 8, dYh + o, idesc2, 1u);
              }
            }
            umma_commit(&empty2[s]);
            if (kb == nkb2 - 1) umma_commit(&acc_full);
          }
          __syncwarp();
        }
      }
    }
  } else if (g.pf_slots > 0) {
    // ================================ prefetch warps 10, 11 ================================
    // Each warp streams table rows (peer memory when the table is sharded) through a few shared-memory slots into the
    // next step's buffers: bulk load -> mbarrier -> bulk store, S slots per warp re-used round robin.  The loop is warp-uniform (one
    // elected lane issues): the row ids arrive 32 at a time, one coalesced load per lane, and are handed out by shuffle
    // -- a per-row dependent id load by a single lane costs more than the copy itself.
    const int w = warp - 10, S = g.pf_slots;
    const long long nU = g.pf_nU_dev ? *g.pf_nU_dev : g.pf_nU;
    const long long total = nU + g.pf_nNeg;
    const long long nhalf = total > g.pf_parity ? (total - g.pf_parity + 1) / 2 : 0;   // v = parity + 2 j < total
    const long long stride = 2ll * gridDim.x, j0 = 2ll * blockIdx.x + w;
    const long long n = j0 < nhalf ? (nhalf - j0 + stride - 1) / stride : 0;
    uint8_t* slots = ring + g.pf_off + (size_t)w * S * g.pf_row_bytes;
    auto vrow = [&](long long k) { return g.pf_parity + 2 * (j0 + k * stride); };
    auto fetch_ids = [&](long long kb) {            // ids of items kb .. kb+31, one per lane
      const long long k = kb + lane;
      if (k >= n) return 0ll;
      const long long v = vrow(k);
      return v < nU ? g.pf_node_ids[v] : g.pf_neg_ids[v - nU];
    };
    long long ids_lo = fetch_ids(0), ids_hi = fetch_ids(32);      // items [base, base+32) and [base+32, base+64)
    long long base = 0;
    auto load = [&](long long k) {                  // k in [base, base + 64)
      const int o = (int)(k - base);
      const long long id = __shfl_sync(0xffffffffu, o < 32 ? ids_lo : ids_hi, o & 31);
      const int s = (int)(k % S);
      if (elect_one()) {
        mbar_expect_tx(&pf_full[w][s], g.pf_row_bytes);
        bulk_g2s(slots + (size_t)s * g.pf_row_bytes, row_ptr(g.xtab, id), g.pf_row_bytes, &pf_full[w][s]);
      }
      __syncwarp();
    };
    for (long long k = 0; k < n && k < S; ++k) load(k);
    // A slot is re-loaded L stores after its own store was issued (S - L loads in flight).  L = 1 keeps the most loads in
    // flight, which is what matters when the rows are remote (8 GPUs: ~4 us per row); a larger L (KGE_B200_PF_LAG) never
    // waits on the copy engine's queue for the newest store -- measured equal at 2 GPUs.
    const int L = g.pf_lag < S - 1 ? g.pf_lag : (S > 2 ? S - 2 : 1);
    for (long long k = 0; k < n; ++k) {
      const int s = (int)(k % S);
      mbar_wait(&pf_full[w][s], (uint32_t)((k / S) & 1));
      const long long v = vrow(k);
      float* dst = v < nU ? g.pf_nc + v * (long long)g.D : g.pf_bn + (v - nU) * (long long)g.D;
      const long long kr = k - L + S;                // the item that takes over the slot of item k - L
      const bool reload = k >= L && kr < n;
      // rotate the id window one batch early: the fresh batch is needed 32 items from now
      if (reload && kr >= base + 32) { base += 32; ids_lo = ids_hi; ids_hi = fetch_ids(base + 32); }
      if (elect_one()) {
        bulk_s2g(dst, slots + (size_t)s * g.pf_row_bytes, g.pf_row_bytes);
        bulk_commit();
        if (reload) { if (L == 3) bulk_wait_read<3>(); else if (L == 2) bulk_wait_read<2>(); else bulk_wait_read<1>(); }
      }
      __syncwarp();
      if (reload) load(kr);
    }
    if (elect_one()) bulk_wait_all();
    __syncwarp();
  }
  } else {
    // ================================ epilogue warps 0..7 ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int ehalf = warp >> 2;                  // two warps share a quarter and split the columns
    const int row = q * 32 + lane;                // accumulator row = lane-side row inside the tile
    const int et = threadIdx.x;                   // 0..255
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int h0 = ((g.N1 >> 1) + 15) & ~15;
    const int cb = ehalf ? h0 : 0, ce = ehalf ? g.N1 : h0;
    uint32_t nacc = 0, it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int c = tile / mtiles, m0 = (tile % mtiles) * kTileM;
      const int m = m0 + row;
      const bool row_ok = m < g.Rx;
      const long long gx = (long long)c * g.Rx + m;
      epi_bar();                                  // everybody is done with the previous tile's shared constants
      for (int y = et; y < g.N1; y += 256) {
        const bool ok = y < g.Ry;
        const long long gy = (long long)c * g.Ry + y;
        colA[y] = (l2 && ok) ? g.y2[gy] : 0.f;
        if (MODE == F_N) { colB[y] = ok ? g.cstat_m[gy] : 0.f; colC[y] = ok ? g.cstat_k[gy] : 0.f; }
      }
      epi_bar();
      const float x2v = (l2 && row_ok) ? g.x2[gx] : 0.f;
      float colsum = 0.f;
      // mode N, one GPU: table rows of the 4 lane-side rows this lane handles in the transposed epilogue mapping
      // (ids fetched now, while GEMM1 runs: the epilogue's row loads then depend on nothing)
      const float* brow[4] = {nullptr, nullptr, nullptr, nullptr};
      const bool bdirect = MODE == F_N && (g.xids || g.xraw);
      if (bdirect) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int mr = m0 + q * 32 + (lane >> 2) + 8 * it;
          const long long xr = (long long)c * g.Rx + (mr < g.Rx ? mr : 0);
          brow[it] = g.xraw ? g.xraw + xr * (long long)g.D : row_ptr(g.xtab, g.xids[xr]);
        }
      }
      mbar_wait(&s_full, it & 1);
      tc_fence_after();
      const bool probe = g.dbg && it == 0 && threadIdx.x == 0;
      if (probe) g.dbg[blockIdx.x * 16 + 2] = gtime();

      float rscale = 1.f;          // mode P: 1 / softmax denominator, applied to the rows of GA in the GEMM2 epilogue
      if (MODE == F_P) {
        const float w_i = (g.wt && row_ok) ? g.wt[gx] : 1.f;
        const float kw = w_i * g.inv2B;
        // ---- pass A: scores (distance epilogue for TransE_l2), running max; 1/dist parked in TMEM region 2 ----
        float mxl = -INFINITY;
        if (l2 || g.adversarial || g.dumpS) {
          for (int col = cb; col < ce; col += 16) {
            float v[16], rr[16];
            tmem_ld16(trow + col, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float s = v[e];
              if (l2) {
                // batched_l2_dist (score_fun.py:26-34): (|b|^2 - 2 a.b) + |a|^2, clamp 1e-30, sqrt
                const float sq = fmaf(-2.f, v[e], colA[col + e]) + x2v;
                const float sqc = fmaxf(sq, 1e-30f);
                const float r = rsqrta(sqc);
                s = g.gamma - sqc * r;
                rr[e] = (sq > 1e-30f) ? r : 0.f;        // clamped distance: zero gradient (clamp_min_), dist ~ 0
                v[e] = s;
              }
              if (col + e < g.Ry) mxl = fmaxf(mxl, s * g.Tl2e);
            }
            if (g.dumpS && row_ok) {
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                if (col + q4 * 4 < g.Ry) st4(g.dumpS + gx * g.Ry + col + q4 * 4, make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]));
            }
            if (l2) tmem_st16(trow + colR2 + col, rr);
          }
          if (l2) tmem_wait_st();
        }
        if (g.adversarial) {
          xch[0][ehalf][row] = mxl;
          epi_bar();
          mxl = fmaxf(mxl, xch[0][ehalf ^ 1][row]);
        } else {
          mxl = 0.f;
        }
        // ---- pass C: softmax numerators, loss terms and (unnormalised) backward coefficients -> TMEM as TF32 hi | lo.
        //      The 1/denominator of the row is a per-row scalar: it is applied to the loss sums here and to the row of
        //      GA in the GEMM2 epilogue, so no separate denominator pass over TMEM is needed.
        float nls = 0.f, rs = 0.f, den = 0.f;
        for (int col = cb; col < ce; col += 16) {
          float v[16], rr[16], hi[16], lo[16];
          tmem_ld16(trow + col, v);
          if (l2) tmem_ld16(trow + colR2 + col, rr);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float s = v[e], rinv = 1.f;
            if (l2) {
              const float sqc = fmaxf(fmaf(-2.f, v[e], colA[col + e]) + x2v, 1e-30f);
              rinv = rr[e];
              s = g.gamma - sqc * rinv;
            }
            const float pe = g.adversarial ? ex2a(fmaf(s, g.Tl2e, -mxl)) : 1.f;
            const float t = ex2a(-fabsf(s) * kLog2e);
            const float u = 1.f + t;
            const float r1 = rcpa(u);
            const float sig = (s >= 0.f) ? r1 : t * r1;                     // sigmoid(s)
            const float sp = fmaf(kLn2, lg2a(u), fmaxf(s, 0.f));            // -logsigmoid(-s)
            const bool ok = row_ok && (col + e < g.Ry);
            float coef = ok ? pe * sig * kw * rinv : 0.f;                    // dL/dneg_ij (/ dist) * denominator
            if (ok) { nls = fmaf(pe, sp, nls); den += pe; }
            rs += coef;
            split_tf32(coef, hi[e], lo[e]);
            v[e] = coef;
          }
          if (g.dumpV && row_ok) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              if (col + q4 * 4 < g.Ry) st4(g.dumpV + gx * g.Ry + col + q4 * 4, make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]));
          }
          tmem_st16(trow + col, hi);
          tmem_st16(trow + colR2 + col, lo);
        }
        xch[1][ehalf][row] = den;
        xch[2][ehalf][row] = nls;
        xch[3][ehalf][row] = rs;
        epi_bar();
        den += xch[1][ehalf ^ 1][row];
        rscale = g.adversarial ? (row_ok ? 1.f / den : 0.f) : g.uni;
        if (g.dumpV && row_ok) {        // test hook: the dump shows the normalised coefficients
          for (int col = cb; col < ce && col < g.Ry; col += 4) {
            float4 x = ld4(g.dumpV + gx * g.Ry + col);
            st4(g.dumpV + gx * g.Ry + col, f4_scale(x, rscale));
          }
        }
        if (ehalf == 0 && row_ok) {
          nls += xch[2][1][row];
          rs += xch[3][1][row];
          const float ps = g.pos[gx];
          const float wb = g.wt ? *g.wbar : 1.f;        // loss.py:75,82: [B] * [B,1] -> mean(pl) * mean(w)
          g.pl[gx] = softplusf(-ps);
          g.nl[gx] = nls * rscale * w_i;
          g.gpos[gx] = -sigmoidf(-ps) * wb * g.inv2B;
          if (l2) g.rowsum[gx] = rs * rscale;
          g.stat_m[gx] = mxl;
          g.stat_k[gx] = kw * rscale;
        }
      } else {
        // ---- mode N: one pass, the softmax statistics of every column (positive) come from mode P ----
        float cs = 0.f;
        for (int col = cb; col < ce; col += 16) {
          float v[16], hi[16], lo[16];
          tmem_ld16(trow + col, v);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float s = v[e], rinv = 1.f;
            if (l2) {
              const float sq = fmaf(-2.f, v[e], x2v) + colA[col + e];
              const float sqc = fmaxf(sq, 1e-30f);
              const float r = rsqrta(sqc);
              s = g.gamma - sqc * r;
              rinv = (sq > 1e-30f) ? r : 0.f;
            }
            const float pe = g.adversarial ? ex2a(fmaf(s, g.Tl2e, -colB[col + e])) : 1.f;
            const float t = ex2a(-fabsf(s) * kLog2e);
            const float r1 = rcpa(1.f + t);
            const float sig = (s >= 0.f) ? r1 : t * r1;
            float coef = pe * colC[col + e] * sig * rinv;
            if (!(row_ok && (col + e < g.Ry))) coef = 0.f;
            cs += coef;
            split_tf32(coef, hi[e], lo[e]);
            v[e] = coef;
          }
          if (g.dumpV && row_ok) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              if (col + q4 * 4 < g.Ry) st4(g.dumpV + gx * g.Ry + col + q4 * 4, make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]));
          }
          tmem_st16(trow + col, hi);
          tmem_st16(trow + colR2 + col, lo);
        }
        xch[2][ehalf][row] = cs;
        epi_bar();
        colsum = cs + xch[2][ehalf ^ 1][row];
      }
      // V is complete in TMEM: hand it to the MMA thread
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_ready);
      if (probe) g.dbg[blockIdx.x * 16 + 3] = gtime();
      const int tr = lane >> 2, tc4 = (lane & 3) * 4;                 // transposed mapping: rows tr + 8*it, columns tc4..tc4+3
      const int mrow0 = m0 + q * 32;                                  // first lane-side row of this warp
      float4 bnext[12];
      // b values of chunk `chn` in the transposed mapping: piece pc = i / 4, rows tr + 8 * (i % 4), columns tc4..tc4+3
      auto load_b = [&](int chn) {
        const int d0n = chn * g.Wc;
        int nbn = (g.D - d0n + 31) >> 5;
        if (nbn > (g.Wc >> 5)) nbn = g.Wc >> 5;
        const int Ncn = nbn * 32;
        const int hcn = ((Ncn >> 1) + 15) & ~15;
        const int cbn = ehalf ? hcn : 0, cen = ehalf ? Ncn : hcn;
        const int npn = (cen - cbn + 15) >> 4;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const int pc = i >> 2, it = i & 3;
          const int k = d0n + cbn + pc * 16 + tc4, mr = mrow0 + tr + 8 * it;
          if (chn < nchunks && pc < npn && k < g.D && mr < g.Rx) {
            if (bdirect) {
              bnext[i] = ld4(brow[it] + k);
            } else {
              const long long so = slab_off(c, g.nblkD, g.Rx, mr, k);
              bnext[i] = f4_add(ld4(g.Xhi + so), ld4(g.Xlo + so));
            }
          } else {
            bnext[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      if (MODE == F_N) load_b(0);                                     // chunk 0's rows: in flight during its MMAs

      // ---- GEMM2 epilogue, one output-column chunk at a time ----
      // TMEM hands every thread one accumulator ROW; writing rows from 32 lanes touches 32 cache lines per
      // instruction (L1 is a few KB next to the 200+ KB of shared memory, so nothing merges).  Each warp therefore
      // transposes its 32 x 16 pieces through a private shared-memory tile: afterwards a lane owns 4 consecutive
      // columns of 4 rows (lane / 4 + 8 * it) and a warp instruction covers 8 rows x 64 contiguous bytes.
      float* stile = epi_stage + warp * (32 * kStagePitch);
      // per-row scalars of this tile in the transposed mapping
      if (ehalf == 0) rowscal[row] = (MODE == F_P) ? rscale : colsum;
      epi_bar();
      float rsc[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) rsc[it] = rowscal[q * 32 + tr + 8 * it];
      float gsq[4] = {0.f, 0.f, 0.f, 0.f};
      for (int ch = 0; ch < nchunks; ++ch, ++nacc) {
        const int d0 = ch * g.Wc;
        int nb = (g.D - d0 + 31) >> 5;
        if (nb > (g.Wc >> 5)) nb = g.Wc >> 5;
        const int Nc = nb * 32;
        const int hc = ((Nc >> 1) + 15) & ~15;
        const int cb2 = ehalf ? hc : 0, ce2 = ehalf ? Nc : hc;
        const int npieces = (ce2 - cb2 + 15) >> 4;
        // mode N: the rows' own values b (for -colsum*b and the regulariser) do not depend on the accumulator.  Their
        // loads are software-pipelined one chunk ahead (issued right after the previous chunk's accumulator was read out
        // of TMEM), so that their latency -- microseconds while the TMA stream saturates the L2 path -- hides behind
        // this chunk's MMAs.
        float4 bq[12];
        if (MODE == F_N && npieces <= 3) {
#pragma unroll
          for (int i = 0; i < 12; ++i) bq[i] = bnext[i];
        }
        if (probe && ch == 1) g.dbg[blockIdx.x * 16 + 8] = gtime();          // b loads issued (and summed)
        mbar_wait(&acc_full, nacc & 1);
        tc_fence_after();
        if (probe && ch == 1) g.dbg[blockIdx.x * 16 + 9] = gtime();          // accumulator of chunk 1 ready
        if (probe && ch == 0) g.dbg[blockIdx.x * 16 + 4] = gtime();
        if (probe && ch == nchunks - 1) g.dbg[blockIdx.x * 16 + 5] = gtime();
        auto release = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty);
        };
        // one 32-row x 16-column piece: registers (row per lane) -> tile -> (4 columns of 4 rows per lane) -> global
        auto process_pre = [&](const uint32_t* r, int col, const int bidx) {      // rows' own values prefetched into bq[bidx..bidx+3]
          __syncwarp();                                              // the previous piece has been read out of the tile
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<float4*>(stile + lane * kStagePitch + q4 * 4) =
                make_float4(__uint_as_float(r[q4 * 4]), __uint_as_float(r[q4 * 4 + 1]), __uint_as_float(r[q4 * 4 + 2]), __uint_as_float(r[q4 * 4 + 3]));
          __syncwarp();
          const int k = d0 + col + tc4;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int mr = mrow0 + tr + 8 * it;
            if (k >= g.D || mr >= g.Rx) continue;
            float4 o = *reinterpret_cast<const float4*>(stile + (tr + 8 * it) * kStagePitch + tc4);
            if (MODE == F_P) o = f4_scale(o, rsc[it]);            // 1 / softmax denominator of the row
            if (MODE == F_N) {
              float4 b;
              b = bq[bidx + it];
              if (l2) o = f4_fma(b, -rsc[it], o);                   // sum_i V_ij a_i - (sum_i V_ij) b_j
              o = f4_add(o, reg_grad4_fast(b, g.reg_norm, g.reg_coef));
              gsq[it] += f4_dot(o, o);
            }
            st4(g.out + ((long long)c * g.Rx + mr) * (long long)g.D + k, o);
          }
        };
        auto process_ld = [&](const uint32_t* r, int col) {                      // wide chunks: the rows' own values are loaded here
          __syncwarp();                                              // the previous piece has been read out of the tile
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<float4*>(stile + lane * kStagePitch + q4 * 4) =
                make_float4(__uint_as_float(r[q4 * 4]), __uint_as_float(r[q4 * 4 + 1]), __uint_as_float(r[q4 * 4 + 2]), __uint_as_float(r[q4 * 4 + 3]));
          __syncwarp();
          const int k = d0 + col + tc4;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int mr = mrow0 + tr + 8 * it;
            if (k >= g.D || mr >= g.Rx) continue;
            float4 o = *reinterpret_cast<const float4*>(stile + (tr + 8 * it) * kStagePitch + tc4);
            if (MODE == F_P) o = f4_scale(o, rsc[it]);            // 1 / softmax denominator of the row
            if (MODE == F_N) {
              float4 b;
              if (bdirect) b = ld4(brow[it] + k);
              else { const long long so = slab_off(c, g.nblkD, g.Rx, mr, k); b = f4_add(ld4(g.Xhi + so), ld4(g.Xlo + so)); }
              if (l2) o = f4_fma(b, -rsc[it], o);                   // sum_i V_ij a_i - (sum_i V_ij) b_j
              o = f4_add(o, reg_grad4_fast(b, g.reg_norm, g.reg_coef));
              gsq[it] += f4_dot(o, o);
            }
            st4(g.out + ((long long)c * g.Rx + mr) * (long long)g.D + k, o);
          }
        };
        if (npieces <= 3) {
          // whole half-chunk in registers: the accumulator is released before any global traffic
          uint32_t r[48];
#pragma unroll
          for (int pc = 0; pc < 3; ++pc)
            if (pc < npieces) tmem_ld16_nowait(trow + colAcc + cb2 + pc * 16, r + pc * 16);
          tmem_wait_ld();
          release();
          if (MODE == F_N) load_b(ch + 1);                            // next chunk's rows, one chunk ahead
          if (probe && ch == 1) g.dbg[blockIdx.x * 16 + 10] = gtime();       // TMEM read, accumulator released
#pragma unroll
          for (int pc = 0; pc < 3; ++pc)
            if (pc < npieces) process_pre(r + pc * 16, cb2 + pc * 16, pc * 4);
          if (probe && ch == 1) g.dbg[blockIdx.x * 16 + 11] = gtime();       // chunk 1 stored
          if (probe && ch == 0) g.dbg[blockIdx.x * 16 + 7] = gtime();        // chunk 0 stored
        } else {
          for (int pc = 0; pc < npieces; ++pc) {
            uint32_t r[16];
            tmem_ld16_nowait(trow + colAcc + cb2 + pc * 16, r);
            tmem_wait_ld();
            if (pc == npieces - 1) release();
            process_ld(r, cb2 + pc * 16);
          }
          if (npieces == 0) release();
        }
      }
      if (probe) g.dbg[blockIdx.x * 16 + 6] = gtime();
      if (MODE == F_N) {
        // mean(G_neg^2) per row: the 4 lanes that share a row, then the two warps that share the quarter
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          float v = gsq[it];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if ((lane & 3) == 0) xch[3][ehalf][q * 32 + tr + 8 * it] = v;
        }
        epi_bar();
        if (ehalf == 0 && row_ok) g.gsn[gx] = (xch[3][0][row] + xch[3][1][row]) / (float)g.D;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

int pad16(int x) { return (x + 15) & ~15; }

}  // namespace

// The fused kernel keeps a whole row of the chunk's score matrix in TMEM twice (hi | lo) next to the GEMM2
// accumulator: 2 * pad16(columns) + 32 <= 512 TMEM columns.
bool fused_supported(const StepParams& p) {
  const bool model_ok = p.model == KGE_TRANSE_L2 || p.model == KGE_DISTMULT || p.model == KGE_COMPLEX || p.model == KGE_RESCAL;
  if (!model_ok) return false;
  if (p.hinge || p.pairwise || p.neg_deg) return false;   // the fused epilogue is the plain (non-pairwise, unmasked) Logsigmoid criterion
  if ((p.D % 8) || (p.Cs % 8) || (p.Ns % 8) || p.D < 32 || p.Cs < 8 || p.Ns < 8) return false;
  return pad16(p.Cs) <= 240 && pad16(p.Ns) <= 240;
}

// mode 0 (P): S = A.Bn^T -> loss, coefficients -> GA;  mode 1 (N): S^T -> coefficients -> G_neg (+ mean square)
namespace {
// GEMM stage geometry of one mode + what the ring leaves for prefetch row slots
struct Geometry { int Rx, Ry, N1, Wc, nS1, nS2, pf_slots; uint32_t stage1Bytes, stage2Bytes, pf_off; bool ok; };
Geometry geometry(const StepParams& p, int mode, bool want_prefetch) {
  Geometry q{};
  const bool P = mode == 0;
  q.Rx = P ? p.Cs : p.Ns; q.Ry = P ? p.Ns : p.Cs;
  q.N1 = pad16(q.Ry);
  int wc = (512 - 2 * q.N1) & ~31;
  if (wc > 256) wc = 256;
  const int dpad = (p.D + 31) & ~31;
  if (wc > dpad) wc = dpad;
  q.Wc = wc;
  q.stage1Bytes = 2u * 16384u + 2u * (uint32_t)q.N1 * 128u;
  q.stage2Bytes = 2u * (uint32_t)(wc >> 5) * 4096u;
  q.nS1 = (int)(kRingBytes / q.stage1Bytes); if (q.nS1 > kMaxS1) q.nS1 = kMaxS1;
  q.nS2 = (int)(kRingBytes / q.stage2Bytes); if (q.nS2 > kMaxS2) q.nS2 = kMaxS2;
  q.ok = q.nS1 >= 2 && q.nS2 >= 2 && wc >= 32;
  if (q.ok && want_prefetch) {
    // GEMM1 keeps its stages; GEMM2 gives up stages (never below 4) until both prefetch warps have kMaxPf row slots
    const uint32_t row = (uint32_t)p.D * 4u;
    int nS1 = q.nS1, nS2 = q.nS2;
    uint32_t want = 2u * kMaxPf * row;
    while (nS1 > 2 && kRingBytes - (uint32_t)nS1 * q.stage1Bytes < want) --nS1;   // a deep GEMM1 ring gives up stages first
    uint32_t used = (uint32_t)nS1 * q.stage1Bytes;
    if (want > kRingBytes - used) want = kRingBytes - used;           // never more than GEMM1 leaves
    while (nS2 > 4 && kRingBytes - (uint32_t)nS2 * q.stage2Bytes < want) --nS2;
    if ((uint32_t)nS2 * q.stage2Bytes > used) used = (uint32_t)nS2 * q.stage2Bytes;
    int slots = (int)((kRingBytes - used) / (2u * row));
    if (slots > kMaxPf) slots = kMaxPf;
    if (slots >= 2) { q.pf_slots = slots; q.nS1 = nS1; q.nS2 = nS2; q.pf_off = used; }
  }
  return q;
}
}  // namespace

int fused_prefetch_slots(const StepParams& p, int mode) { return geometry(p, mode, true).pf_slots; }

int fused_launch(const LaunchCtx& c, const StepParams& p, const StepWs& w, int mode, const float* wt, float* dumpS,
                 float* dumpV, const TableView* ent, const long long* neg_ids, const FusedPrefetch* pf, char* err, size_t errlen) {
  FusedArgs g{};
  g.model = p.model; g.adversarial = p.adversarial;
  g.gamma = p.gamma; g.Tl2e = p.adv_temperature * kLog2e; g.inv2B = 0.5f / (float)p.B; g.uni = 1.f / (float)p.Ns;
  g.reg_coef = p.reg_coef; g.reg_norm = p.reg_norm;
  g.C = p.C; g.D = p.D; g.nblkD = slab_blocks(p.D);
  const bool P = mode == 0;
  const Geometry q = geometry(p, mode, pf != nullptr && ent != nullptr);
  if (!q.ok) { snprintf(err, errlen, "fused kernel: shape does not fit (N1=%d)", q.N1); return KGE_ERR_UNSUPPORTED; }
  g.Rx = q.Rx; g.Ry = q.Ry; g.N1 = q.N1; g.Wc = q.Wc; g.nS1 = q.nS1; g.nS2 = q.nS2;
  g.stage1Bytes = q.stage1Bytes; g.stage2Bytes = q.stage2Bytes;
  if (pf && ent && q.pf_slots >= 2) {
    static const int lag_env = getenv("KGE_B200_PF_LAG") ? atoi(getenv("KGE_B200_PF_LAG")) : 1;
    g.pf_lag = lag_env < 1 ? 1 : (lag_env > 3 ? 3 : lag_env);
    g.pf_slots = q.pf_slots; g.pf_off = q.pf_off; g.pf_row_bytes = (uint32_t)p.D * 4u; g.pf_parity = mode;
    g.pf_node_ids = pf->node_ids; g.pf_nU_dev = pf->nU_dev; g.pf_nU = pf->nU; g.pf_nNeg = pf->nNeg;
    g.pf_neg_ids = pf->neg_ids; g.pf_nc = pf->nc; g.pf_bn = pf->bn;
    g.xtab = *ent;
  }
  const float *Xh = P ? w.Ahi : w.Bhi, *Xl = P ? w.Alo : w.Blo, *Yh = P ? w.Bhi : w.Ahi, *Yl = P ? w.Blo : w.Alo;
  g.x2 = P ? w.a2 : w.b2; g.y2 = P ? w.b2 : w.a2;
  g.pos = w.pos; g.wt = wt; g.wbar = w.wbar;
  g.gpos = w.gpos; g.rowsum = w.rowsum; g.pl = w.pl; g.nl = w.nl; g.stat_m = w.stat_m; g.stat_k = w.stat_k;
  g.dumpS = P ? dumpS : nullptr; g.dumpV = dumpV;
  g.cstat_m = w.stat_m; g.cstat_k = w.stat_k;
  g.Xhi = Xh; g.Xlo = Xl;
  // one GPU: the negatives' own rows come straight from the table (exact fp32, one load); sharded tables would make
  // that a remote read per row, so they use the local hi + lo slabs
  if (!P && ent && ent->n_shards == 1 && neg_ids) { g.xids = neg_ids; g.xtab = *ent; }
  if (!P && w.BnRaw) g.xraw = w.BnRaw;
  g.gsn = w.gsn;
  g.out = P ? w.GA : w.Bn;
  const long long rowsX = (long long)p.C * g.Rx * g.nblkD, rowsY = (long long)p.C * g.Ry * g.nblkD;
  CUtensorMap mXh, mXl, mYh1, mYl1, mYh2, mYl2;
  if (!tc_make_map(&mXh, Xh, rowsX, 32, kTileM, err, errlen) || !tc_make_map(&mXl, Xl, rowsX, 32, kTileM, err, errlen) ||
      !tc_make_map(&mYh1, Yh, rowsY, 32, g.N1, err, errlen) || !tc_make_map(&mYl1, Yl, rowsY, 32, g.N1, err, errlen) ||
      !tc_make_map(&mYh2, Yh, rowsY, 32, 32, err, errlen, true) || !tc_make_map(&mYl2, Yl, rowsY, 32, 32, err, errlen, true))
    return KGE_ERR_CUDA;
  const size_t smem = kRingBytes + kEpiStageBytes + 1024;
  static bool attr_set[2][64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[mode][dev]) {
    cudaError_t e = P ? cudaFuncSetAttribute(k_fused<F_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(k_fused<F_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute(k_fused): %s", cudaGetErrorString(e)); return KGE_ERR_CUDA; }
    attr_set[mode][dev] = true;
  }
  const int mtiles = (g.Rx + kTileM - 1) / kTileM;
  int grid = p.C * mtiles;
  if (grid > c.num_sms) grid = c.num_sms;
  static const bool timing = getenv("KGE_B200_FUSED_TIMING") != nullptr;
  static const bool halfload = getenv("KGE_B200_FUSED_HALFLOAD") != nullptr;
  g.exp_halfload = halfload ? 1 : 0;
  unsigned long long* dbg = nullptr;
  if (timing) { cudaMalloc(&dbg, (size_t)grid * 16 * sizeof(unsigned long long)); cudaMemset(dbg, 0, (size_t)grid * 128); g.dbg = dbg; }
  if (P) KGE_LAUNCH_NAMED(c, "k_fused<P: S=A.Bn^T, loss, GA=V.Bn>", k_fused<F_P>, grid, kThreadsF, smem, mXh, mXl, mYh1, mYl1, mYh2, mYl2, g);
  else KGE_LAUNCH_NAMED(c, "k_fused<N: S^T, G_neg=V^T.A, mean sq>", k_fused<F_N>, grid, kThreadsF, smem, mXh, mXl, mYh1, mYl1, mYh2, mYl2, g);
  if (timing) {
    cudaStreamSynchronize(c.stream);
    unsigned long long* hb = (unsigned long long*)malloc((size_t)grid * 128);
    cudaMemcpy(hb, dbg, (size_t)grid * 128, cudaMemcpyDeviceToHost);
    double t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long mn = ~0ull, mx = 0;
    for (int i = 0; i < grid; ++i) {
      const unsigned long long* r = hb + (size_t)i * 16;
      for (int k = 1; k < 12; ++k) t[k] += (double)(r[k] - r[0]);
      if (r[0] < mn) mn = r[0];
      if (r[6] > mx) mx = r[6];
    }
    fprintf(stderr, "[fused timing] mode %c ctas=%d tiles=%d (first tile, us since CTA start) first_stage=%.2f gemm1_done=%.2f softmax_done=%.2f "
            "gemm2_chunk0=%.2f gemm2_last=%.2f tile_end=%.2f span=%.2f  (nS1=%d nS2=%d Wc=%d N1=%d)\n"
            "               epilogue warp 2: chunk0 stored=%.2f | chunk1: b loaded=%.2f acc ready=%.2f released=%.2f stored=%.2f\n",
            P ? 'P' : 'N', grid, p.C * mtiles,
            t[1] / grid / 1e3, t[2] / grid / 1e3, t[3] / grid / 1e3, t[4] / grid / 1e3, t[5] / grid / 1e3, t[6] / grid / 1e3,
            (double)(mx - mn) / 1e3, g.nS1, g.nS2, g.Wc, g.N1,
            t[7] / grid / 1e3, t[8] / grid / 1e3, t[9] / grid / 1e3, t[10] / grid / 1e3, t[11] / grid / 1e3);
    free(hb); cudaFree(dbg);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(err, errlen, "k_fused launch: %s", cudaGetErrorString(e)); return KGE_ERR_CUDA; }
  return KGE_OK;
}

}  // namespace kge
