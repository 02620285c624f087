// Tensor engine, N-split variant ("tc3"): same fused computation as pnr_field_tc.cu, different mapping.
//
// pnr_field_tc.cu gives each CTA of a pair 64 rows x all 512 features (tcgen05 cta_group::2, M=128 per pair).
// ncu showed that shape to be operand-delivery bound: with only 64 rows per SM every MMA moves a full
// 256 x 16 weight tile between the two SMs and runs at about half rate (tensor math active 53 % while the
// issuing thread is blocked 96 % of the time; profiles/r1_k_field_tc_v22.txt).
//
// Here the CTA pair still owns a tile of 128 points, but the split is along FEATURES:
//   * each CTA holds ALL 128 rows and HALF of the hidden features: X_c, H_c = 128 x 256 fp32 -> 256 + 256 TMEM
//     columns (lane = row), so every MMA is a full-rate cta_group::1 M=128 N=256 K=16;
//   * each CTA streams only its half of every weight matrix (256 output features x 512 k);
//   * a layer's A operand (relu(activation), fp16 hi/lo, K = 512) is produced half by each CTA: the epilogue
//     writes every 16-byte unit twice -- into its own ring slot and, through DSMEM (st.shared::cluster), into
//     the peer's ring slot.  The A ring has NA slots of 32 KB (128 rows x 64 k, hi + lo); "positions"
//     alternate my chunk / peer's chunk, and slot (pos % NA) may be rewritten once BOTH CTAs' MMAs have
//     consumed position pos - NA (tcgen05.commit multicast to both CTAs' F barriers);
//   * lin_out (512 -> 4) is one more tensor-core layer with N = 16, so no cross-CTA reduction is needed.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "pnr_common.cuh"
#include "pnr_geom.cuh"
#include "pnr_tc_ptx.cuh"

namespace pnr {
namespace tc3 {

using namespace tcptx;

constexpr int D = 512;
constexpr int HALF = 256;                // hidden features per CTA
constexpr int ROWS = 128;                // rows (points) per tile; both CTAs hold all of them
constexpr int NWORKER_WARPS = 16;
constexpr int WARP_MMA = NWORKER_WARPS;
constexpr int NTHREADS = (NWORKER_WARPS + 2) * 32;
constexpr int A_SLOT = 32768;            // 128 rows x 64 k fp16: hi 16 KB then lo 16 KB
constexpr int NA = 3;
constexpr int B_SLOT = 32768;            // 256 weight rows x 64 k fp16 (one of hi / lo)
constexpr int NB = 4;
constexpr uint32_t X_COL = 0, H_COL = 256;

constexpr int SM_A = 0;
constexpr int SM_B = SM_A + NA * A_SLOT;
constexpr int SM_BAR = SM_B + NB * B_SLOT;          // 229376
constexpr int BAR_WL = 0;                            // [NA] chunk written by my own workers (16 warp arrivals)
constexpr int BAR_WR = BAR_WL + NA;                  // [NA] chunk written by the peer (st.async complete_tx, 32 KB)
constexpr int BAR_F = BAR_WR + NA;                   // [NA] slot free in BOTH CTAs (2 commits)
constexpr int BAR_BF = BAR_F + NA;                   // [NB] weight slot landed (tx)
constexpr int BAR_BE = BAR_BF + NB;                  // [NB] weight slot free (commit)
constexpr int BAR_ACC = BAR_BE + NB;
constexpr int BAR_COUNT = BAR_ACC + 1;
constexpr int SM_TMEM_PTR = SM_BAR + BAR_COUNT * 8;
constexpr int SMEM_BYTES = SM_BAR + 256;

// packed weights, per rank: [lin_in hi, lo][10 fc layers x 8 positions x (hi, lo)][lin_out]
constexpr int SLOTS_LIN_IN = 2;
constexpr int SLOTS_FC = 16;
constexpr int SLOTS_PER_RANK = SLOTS_LIN_IN + 10 * SLOTS_FC + 1;   // 163
constexpr int HEADER_BYTES = 256;

// k-chunk (64-feature block of the layer input) consumed at position t (0..7) of a layer by CTA `rank`
__host__ __device__ __forceinline__ int kidx_of(int t, int rank) { return ((t & 1) ? 4 * (1 - rank) : 4 * rank) + (t >> 1); }

struct Params {
  PnrScene sc;
  PointSource src;
  PnrMlp mlp;
  const uint8_t* packed;   // tc3 section of the packed weights (starts with the 256 B header)
  const float* proj;       // [3][V][Hl][Wl][512]
  float* scratch;          // [gridDim.x][256][128]
  float* out;              // [total_points][4]
  int64_t total_points;
  int64_t n_tiles;
  int* status;
  int debug;   // experiment switches (PNR_TC3_DEBUG): 2 = no gather loads, 4 = no MMA issue
};

enum { MODE_GATHER = 0, MODE_BIAS_WB = 1, MODE_HIDDEN = 2, MODE_COMBINE = 3, MODE_FINAL = 4 };

struct WorkerCtx {
  uint32_t smem_u;      // shared::cta address of smem base
  uint32_t peer_u;      // the same offset in the peer CTA (shared::cluster address)
  uint32_t tmem;        // base tmem address incl. this warp's lane quarter
  uint32_t bar_base;
  uint32_t rank;
  int lane, s, row;
  float w_scale, w_inv;
};

__device__ __forceinline__ void gather_issue(float4* g, const float* __restrict__ proj_i, const uint32_t* off, int n0) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4* src = reinterpret_cast<const float4*>(proj_i + off[k] + n0);
    g[2 * k] = __ldg(src);
    g[2 * k + 1] = __ldg(src + 1);
  }
}

// TMEM load of 8 columns WITHOUT waiting: the registers are valid only after tmem_ld_wait8 (which carries them as
// in/out operands so that no use can be scheduled ahead of the wait).
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// relu(y0), relu(y1) -> packed fp16 hi pair and lo pair (error-compensated split); F2FP packs two values per
// instruction on the fast pipe and saturates instead of producing inf.
__device__ __forceinline__ void split_relu2(float y0, float y1, uint32_t& hi, uint32_t& lo) {
  const float a0 = fmaxf(y0, 0.f), a1 = fmaxf(y1, 0.f);
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(a1), "f"(a0));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(a1 - hf.y), "f"(a0 - hf.x));
}

// One epilogue pass over this thread's 64 local features (4 waves x 2 steps of 8) of one layer.
//   pos_base: ring position of the first chunk of the layer that will consume what this pass produces.
// TMEM loads, gathers and bias loads of step i+1 are issued before step i is processed.
template <int MODE>
__device__ __forceinline__ void epilogue(const WorkerCtx& c, const Params& p, uint32_t acc_col,
                                         const float* __restrict__ bias, const float* __restrict__ proj_i,
                                         const uint32_t* off, const float* wt, int view, float* __restrict__ scratch,
                                         uint32_t pos_base, uint32_t acc_phase, int tag, long long& t_acc, long long& t_free) {
  float4 g[8];      // gather taps of the CURRENT step; refilled for the next step right after use
  float4 bb[2][2];
  uint32_t raw[8];
  const int n_base = (int)c.rank * HALF;   // global index of this CTA's first feature
  const bool do_gather = !(p.debug & 2);
  if (MODE == MODE_GATHER) {
    if (do_gather) {
      gather_issue(g, proj_i, off, n_base + c.s * 16);  // in flight during the MMA tail
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    bb[0][0] = __ldg(reinterpret_cast<const float4*>(bias + n_base + c.s * 16));
    bb[0][1] = __ldg(reinterpret_cast<const float4*>(bias + n_base + c.s * 16) + 1);
  }
  mbar_wait_timed(c.bar_base + BAR_ACC * 8, acc_phase, p.status, tag, t_acc);
  tc_fence_after();
  tmem_ld8_issue(c.tmem + acc_col + c.s * 16, raw);
  const int NS = p.sc.NS;
  const bool produce = !(MODE == MODE_COMBINE && view != NS - 1);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int wave = i >> 1, h = i & 1;
    const int f = wave * 64 + c.s * 16 + h * 8;   // local feature = TMEM column inside the accumulator
    const int fn = ((i + 1) >> 1) * 64 + c.s * 16 + ((i + 1) & 1) * 8;   // next step's
    float y[8];
    tmem_ld_wait8(raw);
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(raw[e]) * c.w_inv;
    if (i + 1 < 8) {
      tmem_ld8_issue(c.tmem + acc_col + fn, raw);
      if (MODE != MODE_GATHER) {
        bb[(i + 1) & 1][0] = __ldg(reinterpret_cast<const float4*>(bias + n_base + fn));
        bb[(i + 1) & 1][1] = __ldg(reinterpret_cast<const float4*>(bias + n_base + fn) + 1);
      }
    }
    if (MODE == MODE_GATHER) {
      const float4* t = g;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        y[4 * hh + 0] += ((t[0 + hh].x * wt[0] + t[2 + hh].x * wt[1]) + t[4 + hh].x * wt[2]) + t[6 + hh].x * wt[3];
        y[4 * hh + 1] += ((t[0 + hh].y * wt[0] + t[2 + hh].y * wt[1]) + t[4 + hh].y * wt[2]) + t[6 + hh].y * wt[3];
        y[4 * hh + 2] += ((t[0 + hh].z * wt[0] + t[2 + hh].z * wt[1]) + t[4 + hh].z * wt[2]) + t[6 + hh].z * wt[3];
        y[4 * hh + 3] += ((t[0 + hh].w * wt[0] + t[2 + hh].w * wt[1]) + t[4 + hh].w * wt[2]) + t[6 + hh].w * wt[3];
      }
      if (i + 1 < 8 && do_gather) gather_issue(g, proj_i, off, n_base + fn);  // taps of the next step, in flight during the stores below
    } else {
      const float4 b0 = bb[i & 1][0], b1 = bb[i & 1][1];
      y[0] += b0.x; y[1] += b0.y; y[2] += b0.z; y[3] += b0.w;
      y[4] += b1.x; y[5] += b1.y; y[6] += b1.z; y[7] += b1.w;
    }
    if (MODE == MODE_COMBINE && NS > 1) {
      // multi-view mean (util.combine_interleaved): sum in view order, then divide
      float* sp = scratch + (size_t)f * ROWS + c.row;  // [local feature][row]: lanes are contiguous
      if (view == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) sp[e * ROWS] = y[e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = sp[e * ROWS] + y[e];
        if (view < NS - 1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) sp[e * ROWS] = y[e];
        } else {
          const float ns = (float)NS;
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = y[e] / ns;
        }
      }
    }
    if (MODE == MODE_GATHER || MODE == MODE_BIAS_WB || (MODE == MODE_COMBINE && produce)) {
      float z[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] = y[e] * c.w_scale;
      tmem_st8(c.tmem + X_COL + f, z);  // residual stream write-back
    }
    if (produce) {
      // my chunk `wave` sits at ring position pos_base + 2*wave here and at pos_base + 2*wave + 1 in the peer
      const uint32_t pos_l = pos_base + 2 * wave, pos_r = pos_l + 1;
      const uint32_t slot_l = pos_l % NA, slot_r = pos_r % NA;
      if (h == 0) {
        // both CTAs must have consumed the previous occupants of the two slots
        mbar_wait_timed(c.bar_base + (BAR_F + slot_l) * 8, ((pos_l / NA) + 1) & 1, p.status, 500 + slot_l, t_free);
        mbar_wait_timed(c.bar_base + (BAR_F + slot_r) * 8, ((pos_r / NA) + 1) & 1, p.status, 510 + slot_r, t_free);
      }
      uint4 vhi, vlo;
      split_relu2(y[0], y[1], vhi.x, vlo.x);
      split_relu2(y[2], y[3], vhi.y, vlo.y);
      split_relu2(y[4], y[5], vhi.z, vlo.z);
      split_relu2(y[6], y[7], vhi.w, vlo.w);
      const uint32_t in_slot = (uint32_t)c.row * 128 + (uint32_t)(((2 * c.s + h) ^ (c.row & 7)) * 16);
      const uint32_t lp = c.smem_u + SM_A + slot_l * A_SLOT + in_slot;
      st_shared_v4(lp, vhi);
      st_shared_v4(lp + 16384, vlo);
      // peer's copy: asynchronous DSMEM stores that complete_tx on the peer's WR barrier (no fence, no arrive)
      const uint32_t rp = c.peer_u + SM_A + slot_r * A_SLOT + in_slot;
      const uint32_t rbar = c.peer_u + SM_BAR + (BAR_WR + slot_r) * 8;
      st_async_v4(rp, vhi, rbar);
      st_async_v4(rp + 16384, vlo, rbar);
      if (h == 1) {
        // this warp's 32 x 16 slice of my own copy of chunk `wave` is in place: publish to my MMA warp
        fence_proxy_async();
        if (MODE != MODE_HIDDEN && MODE != MODE_FINAL) tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (c.lane == 0) mbar_arrive(c.bar_base + (BAR_WL + slot_l) * 8);
      }
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) k_field_tc3(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const uint32_t smem_u = smem_u32(smem);
  const uint32_t bar_base = smem_u + SM_BAR;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM_TMEM_PTR);
  const int NS = p.sc.NS;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) {
      mbar_init(bar_base + (BAR_WL + i) * 8, NWORKER_WARPS);
      mbar_init(bar_base + (BAR_WR + i) * 8, 1);
      mbar_init(bar_base + (BAR_F + i) * 8, 2);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(bar_base + (BAR_BF + i) * 8, 1);
      mbar_init(bar_base + (BAR_BE + i) * 8, 1);
    }
    mbar_init(bar_base + BAR_ACC * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const float w_scale = reinterpret_cast<const float*>(p.packed)[0];
  const float w_inv = reinterpret_cast<const float*>(p.packed)[1];
  const uint8_t* slots = p.packed + HEADER_BYTES + (size_t)rank * SLOTS_PER_RANK * B_SLOT;
  const size_t map_stride = (size_t)p.sc.SB * NS * p.sc.Hl * p.sc.Wl * D;
  // ring positions per tile: per view 1 (lin_in) + 6 fc layers x 8, then 4 fc layers x 8 + lin_out x 8
  const uint32_t pos_per_view = 1 + 48;

  if (warp < NWORKER_WARPS) {
    // =============================== worker warps ===============================
    WorkerCtx c;
    c.smem_u = smem_u;
    c.peer_u = mapa_cluster(smem_u, rank ^ 1);
    c.lane = lane;
    const int q = warp & 3;
    c.s = warp >> 2;                  // which 16 columns of every 64-feature chunk
    c.row = 32 * q + lane;            // row of the 128-row tile == TMEM lane
    c.tmem = tmem_base + ((uint32_t)(32 * q) << 16);
    c.bar_base = bar_base;
    c.rank = rank;
    c.w_scale = w_scale;
    c.w_inv = w_inv;
    long long t_acc = 0, t_free = 0;
    const long long t_wstart = clock64();
    float* scratch = p.scratch + (size_t)blockIdx.x * HALF * ROWS;
    uint32_t acc_phase = 0;
    uint32_t pos = 0;                 // ring position of the next layer's first chunk

    for (int64_t tile = pair; tile < p.n_tiles; tile += n_pairs) {
      int64_t pt = tile * ROWS + c.row;
      const bool valid = pt < p.total_points;
      if (!valid) pt = p.total_points - 1;
      const int sb = (int)(pt / p.src.P);
      float x[3], d[3];
      load_point(p.src, pt, x, d);
      for (int v = 0; v < NS; ++v) {
        // ---- geometry of my row for view v; the 42 input channels -> ring position `pos` (lin_in operand),
        //      written locally by BOTH CTAs (each needs all rows, nothing to exchange) ----
        uint32_t off[4];
        float wt[4];
        {
          PointGeom pg = point_geometry(p.sc, sb, v, x, d);
          const uint32_t vbase = (uint32_t)(sb * NS + v) * p.sc.Hl * p.sc.Wl;
          off[0] = (vbase + pg.y0 * p.sc.Wl + pg.x0) * D;
          off[1] = (vbase + pg.y0 * p.sc.Wl + pg.x1) * D;
          off[2] = (vbase + pg.y1 * p.sc.Wl + pg.x0) * D;
          off[3] = (vbase + pg.y1 * p.sc.Wl + pg.x1) * D;
          wt[0] = pg.w_nw; wt[1] = pg.w_ne; wt[2] = pg.w_sw; wt[3] = pg.w_se;
          const uint32_t slot = pos % NA;
          mbar_wait_timed(bar_base + (BAR_F + slot) * 8, ((pos / NA) + 1) & 1, p.status, 520 + slot, t_free);
          uint8_t* row_hi = smem + SM_A + slot * A_SLOT + c.row * 128;
          uint8_t* row_lo = row_hi + 16384;
#pragma unroll 1
          for (int e = 0; e < 12; e += 2) {
            const int ch = c.s * 12 + e;
            float f0 = feat_channel(pg, ch), f1 = feat_channel(pg, ch + 1);
            f0 = fmaxf(fminf(f0, 65504.f), -65504.f);
            f1 = fmaxf(fminf(f1, 65504.f), -65504.f);
            __half h0 = __float2half_rn(f0), h1 = __float2half_rn(f1);
            __half l0 = __float2half_rn(f0 - __half2float(h0)), l1 = __float2half_rn(f1 - __half2float(h1));
            const uint32_t hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            const uint32_t lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            const int byte = ((ch >> 3) ^ (c.row & 7)) * 16 + (ch & 7) * 2;
            *reinterpret_cast<uint32_t*>(row_hi + byte) = hi;
            *reinterpret_cast<uint32_t*>(row_lo + byte) = lo;
          }
          fence_proxy_async();
          tc_fence_before();   // this warp's earlier TMEM reads precede the next lin_in MMA
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_base + (BAR_WL + slot) * 8);
          pos += 1;
        }
        // ---- lin_in, then blocks 0..2 ----
        for (int blk = 0; blk < 3; ++blk) {
          epilogue<MODE_GATHER>(c, p, X_COL, nullptr, p.proj + (size_t)blk * map_stride, off, wt, v, nullptr, pos,
                                acc_phase, 100 + blk, t_acc, t_free);   // X ready -> A of fc_0
          acc_phase ^= 1;
          pos += 8;
          epilogue<MODE_HIDDEN>(c, p, H_COL, p.mlp.fc0_b[blk], nullptr, off, wt, v, nullptr, pos, acc_phase,
                                110 + blk, t_acc, t_free);              // H ready -> A of fc_1
          acc_phase ^= 1;
          pos += 8;
        }
        epilogue<MODE_COMBINE>(c, p, X_COL, p.mlp.fc1_b[2], nullptr, off, wt, v, scratch, pos, acc_phase, 120, t_acc, t_free);
        acc_phase ^= 1;
      }
      pos += 8;  // fc_0 of block 3 consumed what the last view's COMBINE produced
      // ---- blocks 3..4 on the view-averaged rows ----
      epilogue<MODE_HIDDEN>(c, p, H_COL, p.mlp.fc0_b[3], nullptr, nullptr, nullptr, 0, nullptr, pos, acc_phase, 130, t_acc, t_free);
      acc_phase ^= 1;
      pos += 8;
      epilogue<MODE_BIAS_WB>(c, p, X_COL, p.mlp.fc1_b[3], nullptr, nullptr, nullptr, 0, nullptr, pos, acc_phase, 131, t_acc, t_free);
      acc_phase ^= 1;
      pos += 8;
      epilogue<MODE_HIDDEN>(c, p, H_COL, p.mlp.fc0_b[4], nullptr, nullptr, nullptr, 0, nullptr, pos, acc_phase, 132, t_acc, t_free);
      acc_phase ^= 1;
      pos += 8;
      epilogue<MODE_FINAL>(c, p, X_COL, p.mlp.fc1_b[4], nullptr, nullptr, nullptr, 0, nullptr, pos, acc_phase, 133, t_acc, t_free);
      acc_phase ^= 1;
      pos += 8;
      // ---- lin_out accumulator (16 columns at H_COL): sigmoid rgb / relu sigma (models.py:260-264) ----
      mbar_wait_timed(bar_base + BAR_ACC * 8, acc_phase, p.status, 134, t_acc);
      acc_phase ^= 1;
      tc_fence_after();
      if (rank == 0 && c.s == 0) {
        float o[4];
        tmem_ld4(c.tmem + H_COL, o);
        if (valid) {
          const float* bo = p.mlp.lin_out_b;
          float4 r;
          r.x = 1.0f / (1.0f + expf(-(o[0] * w_inv + bo[0])));
          r.y = 1.0f / (1.0f + expf(-(o[1] * w_inv + bo[1])));
          r.z = 1.0f / (1.0f + expf(-(o[2] * w_inv + bo[2])));
          r.w = fmaxf(o[3] * w_inv + bo[3], 0.f);
          reinterpret_cast<float4*>(p.out)[pt] = r;
        }
      }
      tc_fence_before();
    }
    if (threadIdx.x == 0) {
      unsigned long long* cnt = reinterpret_cast<unsigned long long*>(p.status + 2);
      atomicAdd(cnt + 4, (unsigned long long)(clock64() - t_wstart));
      atomicAdd(cnt + 5, (unsigned long long)t_acc);
      atomicAdd(cnt + 6, (unsigned long long)t_free);
    }
  } else if (warp == WARP_MMA) {
    // =============================== MMA issuer (every CTA issues for itself) ===============================
    uint32_t pos = 0, seq = 0;
    long long t_w = 0, t_b = 0, t_dummy = 0;
    const long long t_start = clock64();
    const uint32_t a_base = smem_u + SM_A, b_base = smem_u + SM_B;
    const uint64_t desc0 = make_desc(0);
    const bool issuer = elect_one();
    const bool do_mma = !(p.debug & 4);
    // wait until ring slot `slot` holds the next chunk: my workers arrive on WL; the peer's st.async stores
    // complete_tx 32 KB on WR (armed here).  The DSMEM data arrives through the generic proxy, so a proxy fence
    // precedes the tensor core's (async proxy) reads.
    uint32_t wl_par = 0, wr_par = 0;   // phase parity per slot
    auto wait_chunk = [&](uint32_t slot, bool remote_written) {
      if (remote_written) {
        if (issuer) mbar_expect_tx(bar_base + (BAR_WR + slot) * 8, A_SLOT);
        mbar_wait_timed(bar_base + (BAR_WR + slot) * 8, (wr_par >> slot) & 1, p.status, 230 + slot, t_w);
        wr_par ^= 1u << slot;
        fence_proxy_async();
      } else {
        mbar_wait_timed(bar_base + (BAR_WL + slot) * 8, (wl_par >> slot) & 1, p.status, 220 + slot, t_w);
        wl_par ^= 1u << slot;
      }
      tc_fence_after();
    };
    // one ring position: D[dcol] (+)= A(pos) * W^T for this CTA's 256 output features
    auto fc_position = [&](uint32_t dcol, bool overwrite, int ksteps, bool remote_written) {
      const uint32_t slot = pos % NA;
      wait_chunk(slot, remote_written);
      const uint64_t a_hi = desc0 + ((a_base + slot * A_SLOT) >> 4);
      const uint64_t a_lo = a_hi + (16384 >> 4);
      const uint32_t d = tmem_base + dcol;
      {  // W_hi slot: D += Ahi*Whi + Alo*Whi
        const uint32_t sl = seq % NB, ph = (seq / NB) & 1;
        mbar_wait_timed(bar_base + (BAR_BF + sl) * 8, ph, p.status, 200 + sl, t_b);
        tc_fence_after();
        const uint64_t bd = desc0 + ((b_base + sl * B_SLOT) >> 4);
        if (issuer) {
          if (do_mma) {
            umma_f16_1sm(d, a_hi, bd, IDESC_M128_N256, overwrite ? 0u : 1u);
            umma_f16_1sm(d, a_hi + 2, bd + 2, IDESC_M128_N256, 1u);
            umma_f16_1sm(d, a_hi + 4, bd + 4, IDESC_M128_N256, 1u);
            if (ksteps == 4) umma_f16_1sm(d, a_hi + 6, bd + 6, IDESC_M128_N256, 1u);
            umma_f16_1sm(d, a_lo, bd, IDESC_M128_N256, 1u);
            umma_f16_1sm(d, a_lo + 2, bd + 2, IDESC_M128_N256, 1u);
            umma_f16_1sm(d, a_lo + 4, bd + 4, IDESC_M128_N256, 1u);
            if (ksteps == 4) umma_f16_1sm(d, a_lo + 6, bd + 6, IDESC_M128_N256, 1u);
          }
          umma_commit_local(bar_base + (BAR_BE + sl) * 8);
        }
        __syncwarp();
        ++seq;
      }
      {  // W_lo slot: D += Ahi*Wlo
        const uint32_t sl = seq % NB, ph = (seq / NB) & 1;
        mbar_wait_timed(bar_base + (BAR_BF + sl) * 8, ph, p.status, 200 + sl, t_b);
        tc_fence_after();
        const uint64_t bd = desc0 + ((b_base + sl * B_SLOT) >> 4);
        if (issuer) {
          if (do_mma) {
            umma_f16_1sm(d, a_hi, bd, IDESC_M128_N256, 1u);
            umma_f16_1sm(d, a_hi + 2, bd + 2, IDESC_M128_N256, 1u);
            umma_f16_1sm(d, a_hi + 4, bd + 4, IDESC_M128_N256, 1u);
            if (ksteps == 4) umma_f16_1sm(d, a_hi + 6, bd + 6, IDESC_M128_N256, 1u);
          }
          umma_commit_local(bar_base + (BAR_BE + sl) * 8);
          umma_commit_both(bar_base + (BAR_F + slot) * 8);   // ring slot consumed (seen by both CTAs)
        }
        __syncwarp();
        ++seq;
      }
      ++pos;
    };
    auto fc_layer = [&](uint32_t dcol, bool overwrite) {
      for (int t = 0; t < 8; ++t) fc_position(dcol, overwrite && t == 0, 4, (t & 1) != 0);
      if (issuer) umma_commit_local(bar_base + BAR_ACC * 8);
      __syncwarp();
    };
    for (int64_t tile = pair; tile < p.n_tiles; tile += n_pairs) {
      for (int v = 0; v < NS; ++v) {
        fc_position(X_COL, true, 3, false);                       // lin_in (K = 42 -> 48), written locally
        if (issuer) umma_commit_local(bar_base + BAR_ACC * 8);
        __syncwarp();
        for (int blk = 0; blk < 3; ++blk) {
          fc_layer(H_COL, true);                                   // fc_0
          fc_layer(X_COL, false);                                  // fc_1 accumulates onto the residual
        }
      }
      for (int blk = 3; blk < 5; ++blk) {
        fc_layer(H_COL, true);
        fc_layer(X_COL, false);
      }
      // ---- lin_out as a tensor-core layer: N = 16 (4 used), all 8 k-chunks of hi and lo in ONE weight slot ----
      {
        const uint32_t sl = seq % NB, ph = (seq / NB) & 1;
        mbar_wait_timed(bar_base + (BAR_BF + sl) * 8, ph, p.status, 240 + sl, t_b);
        const uint32_t d = tmem_base + H_COL;
        for (int t = 0; t < 8; ++t) {
          const uint32_t slot = pos % NA;
          wait_chunk(slot, (t & 1) != 0);
          const uint64_t a_hi = desc0 + ((a_base + slot * A_SLOT) >> 4);
          const uint64_t a_lo = a_hi + (16384 >> 4);
          const uint64_t b_hi = desc0 + ((b_base + sl * B_SLOT + kidx_of(t, rank) * 2048) >> 4);
          const uint64_t b_lo = b_hi + (16384 >> 4);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_1sm(d, a_hi + 2 * k, b_hi + 2 * k, IDESC_M128_N16, (t == 0 && k == 0) ? 0u : 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_1sm(d, a_lo + 2 * k, b_hi + 2 * k, IDESC_M128_N16, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_1sm(d, a_hi + 2 * k, b_lo + 2 * k, IDESC_M128_N16, 1u);
            umma_commit_both(bar_base + (BAR_F + slot) * 8);
          }
          __syncwarp();
          ++pos;
        }
        if (issuer) {
          umma_commit_local(bar_base + (BAR_BE + sl) * 8);
          umma_commit_local(bar_base + BAR_ACC * 8);
        }
        __syncwarp();
        ++seq;
      }
    }
    (void)t_dummy;
    if (lane == 0) {
      unsigned long long* cnt = reinterpret_cast<unsigned long long*>(p.status + 2);
      atomicAdd(cnt + 0, (unsigned long long)(clock64() - t_start));
      atomicAdd(cnt + 1, (unsigned long long)t_w);
      atomicAdd(cnt + 2, (unsigned long long)t_b);
    }
  } else {
    // =============================== weight streamer ===============================
    if (lane == 0) {
      uint32_t seq = 0;
      long long t_empty = 0;
      const uint32_t b_base = smem_u + SM_B;
      auto stream = [&](int first, int count) {
        for (int i = 0; i < count; ++i) {
          const uint32_t sl = seq % NB, ph = (seq / NB) & 1;
          mbar_wait_timed(bar_base + (BAR_BE + sl) * 8, ph ^ 1, p.status, 400 + sl, t_empty);
          const uint32_t full = bar_base + (BAR_BF + sl) * 8;
          mbar_expect_tx(full, B_SLOT);
          bulk_g2s(b_base + sl * B_SLOT, slots + (size_t)(first + i) * B_SLOT, B_SLOT, full);
          ++seq;
        }
      };
      for (int64_t tile = pair; tile < p.n_tiles; tile += n_pairs) {
        for (int v = 0; v < NS; ++v) stream(0, SLOTS_LIN_IN + 6 * SLOTS_FC);          // lin_in, blocks 0..2
        stream(SLOTS_LIN_IN + 6 * SLOTS_FC, 4 * SLOTS_FC + 1);                         // blocks 3..4, lin_out
      }
      if (rank == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.status + 2) + 7, (unsigned long long)t_empty);
    }
  }

  // ---- teardown: the peer may still be writing into my ring / arriving on my barriers until it is done too ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  (void)pos_per_view;
}

// ---------------------------------------------------------------------------------------
// weight packing for tc3: per rank, 32 KB slots in consumption order
// ---------------------------------------------------------------------------------------
// fc layers and lin_in: grid (slots_in_layer, 2 ranks); slot ls = t*2 + part (lin_in: t = 0)
__global__ void k_pack_layer3(const float* __restrict__ W, int K, uint8_t* __restrict__ dst_rank0,
                              uint8_t* __restrict__ dst_rank1, const float* __restrict__ header) {
  const int ls = blockIdx.x, rank = blockIdx.y;
  const int part = ls & 1, t = ls >> 1;
  const int kc = kidx_of(t, rank);
  uint8_t* dst = (rank ? dst_rank1 : dst_rank0) + (size_t)ls * B_SLOT;
  const float s = header[0];
  for (int idx = threadIdx.x; idx < 256 * 64; idx += blockDim.x) {
    const int i = idx >> 6, kk = idx & 63;
    const int n = rank * HALF + i, k = (K == D ? kc * 64 : 0) + kk;
    float w = (k < K) ? W[(size_t)n * K + k] * s : 0.f;
    __half hi = __float2half_rn(w);
    __half val = part ? __float2half_rn(w - __half2float(hi)) : hi;
    const int byte = i * 128 + (((kk >> 3) ^ (i & 7)) * 16) + (kk & 7) * 2;
    *reinterpret_cast<__half*>(dst + byte) = val;
  }
}
// lin_out [4][512] -> one slot: part p at p*16 KB, k-chunk kc at kc*2 KB: [16 rows][64 k] (rows >= 4 zero)
__global__ void k_pack_lin_out3(const float* __restrict__ W, uint8_t* __restrict__ dst, const float* __restrict__ header) {
  const float s = header[0];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * 8 * 16 * 64; idx += gridDim.x * blockDim.x) {
    const int kk = idx & 63, i = (idx >> 6) & 15, kc = (idx >> 10) & 7, part = idx >> 13;
    float w = (i < 4) ? W[(size_t)i * D + kc * 64 + kk] * s : 0.f;
    __half hi = __float2half_rn(w);
    __half val = part ? __float2half_rn(w - __half2float(hi)) : hi;
    const int byte = part * 16384 + kc * 2048 + i * 128 + (((kk >> 3) ^ (i & 7)) * 16) + (kk & 7) * 2;
    *reinterpret_cast<__half*>(dst + byte) = val;
  }
}

}  // namespace tc3

// ---- host side -------------------------------------------------------------------------------------
size_t tc3_packed_bytes() { return (size_t)tc3::HEADER_BYTES + (size_t)2 * tc3::SLOTS_PER_RANK * tc3::B_SLOT; }

// `base` = tc3 section; its header (scale, 1/scale) is copied from the tc section's header by the caller
int tc3_pack(const PnrMlp* mlp, uint8_t* base, const float* header_src, cudaStream_t s) {
  PNR_CUDA(cudaMemcpyAsync(base, header_src, tc3::HEADER_BYTES, cudaMemcpyDeviceToDevice, s));
  const float* header = reinterpret_cast<const float*>(base);
  uint8_t* r0 = base + tc3::HEADER_BYTES;
  uint8_t* r1 = r0 + (size_t)tc3::SLOTS_PER_RANK * tc3::B_SLOT;
  tc3::k_pack_layer3<<<dim3(tc3::SLOTS_LIN_IN, 2), 256, 0, s>>>(mlp->lin_in_w, mlp->d_in, r0, r1, header);
  PNR_LAUNCH_CHECK();
  for (int i = 0; i < 5; ++i) {
    size_t o0 = (size_t)(tc3::SLOTS_LIN_IN + (2 * i) * tc3::SLOTS_FC) * tc3::B_SLOT;
    size_t o1 = (size_t)(tc3::SLOTS_LIN_IN + (2 * i + 1) * tc3::SLOTS_FC) * tc3::B_SLOT;
    tc3::k_pack_layer3<<<dim3(tc3::SLOTS_FC, 2), 256, 0, s>>>(mlp->fc0_w[i], tc3::D, r0 + o0, r1 + o0, header);
    PNR_LAUNCH_CHECK();
    tc3::k_pack_layer3<<<dim3(tc3::SLOTS_FC, 2), 256, 0, s>>>(mlp->fc1_w[i], tc3::D, r0 + o1, r1 + o1, header);
    PNR_LAUNCH_CHECK();
  }
  size_t oo = (size_t)(tc3::SLOTS_LIN_IN + 10 * tc3::SLOTS_FC) * tc3::B_SLOT;
  PNR_CUDA(cudaMemsetAsync(r0 + oo, 0, tc3::B_SLOT, s));
  PNR_CUDA(cudaMemsetAsync(r1 + oo, 0, tc3::B_SLOT, s));
  tc3::k_pack_lin_out3<<<16, 256, 0, s>>>(mlp->lin_out_w, r0 + oo, header);
  PNR_LAUNCH_CHECK();
  tc3::k_pack_lin_out3<<<16, 256, 0, s>>>(mlp->lin_out_w, r1 + oo, header);
  PNR_LAUNCH_CHECK();
  return PNR_OK;
}

size_t tc3_workspace_bytes(int pairs) { return (size_t)pairs * 2 * tc3::HALF * tc3::ROWS * sizeof(float) + 1024; }

int tc3_field_eval(const PnrScene& sc, const PnrMlp& mlp, const uint8_t* packed3, const float* proj,
                   const PointSource& src, int64_t total_points, float* out, void* ws, int pairs, int* status,
                   cudaStream_t s) {
  tc3::Params p;
  p.sc = sc;
  p.src = src;
  p.mlp = mlp;
  p.packed = packed3;
  p.proj = proj;
  p.scratch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  p.out = out;
  p.total_points = total_points;
  p.n_tiles = (total_points + tc3::ROWS - 1) / tc3::ROWS;
  p.status = status;
  {
    const char* e = getenv("PNR_TC3_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    PNR_CUDA(cudaFuncSetAttribute(tc3::k_field_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::SMEM_BYTES));
    attr_set[dev] = true;
  }
  prof_before(s);
  tc3::k_field_tc3<<<dim3(pairs * 2), dim3(tc3::NTHREADS), tc3::SMEM_BYTES, s>>>(p);
  prof_after(s);
  PNR_LAUNCH_CHECK();
  return PNR_OK;
}

}  // namespace pnr
