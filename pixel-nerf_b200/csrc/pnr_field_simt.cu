// SIMT (fp32 FFMA) engine for the conditioned MLP: the bring-up / any-shape engine.
// Evaluates ResnetFC (src/model/resnetfc.py:132-184) layer by layer over a chunk of point
// rows with a register-tiled SGEMM whose prologue/epilogue fuse ReLU, bias and the residual
// add.  Row order inside a chunk is point-major: row = local_point * NS + view, so the
// multi-view mean (src/util/util.py:461-471) reduces NS adjacent rows.
#include "pnr_common.cuh"

namespace pnr {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int TM = 8, TN = 8;
constexpr int kGemmThreads = (BM / TM) * (BN / TN);  // 256

// C[M][N] (+)= act(A[M][lda]) * W[N][K]^T + bias.   K % 16 == 0, lda >= K, all 16B aligned.
template <bool RELU_A, bool ACCUM>
__global__ void __launch_bounds__(kGemmThreads)
k_sgemm_nt(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ bias,
           float* __restrict__ C, int ldc, int M, int N, int K) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  // global->smem: each thread moves two float4 of A and two of W per k-tile
  const int lrow = tid / 4;        // 0..63 (+64)
  const int lk = (tid % 4) * 4;    // 0,4,8,12
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int m = m0 + lrow + h * 64;
      ra[h] = (m < M) ? *reinterpret_cast<const float4*>(A + (size_t)m * lda + k0 + lk)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      if (RELU_A) {
        ra[h].x = fmaxf(ra[h].x, 0.f); ra[h].y = fmaxf(ra[h].y, 0.f);
        ra[h].z = fmaxf(ra[h].z, 0.f); ra[h].w = fmaxf(ra[h].w, 0.f);
      }
      int n = n0 + lrow + h * 64;
      rw[h] = (n < N) ? *reinterpret_cast<const float4*>(W + (size_t)n * K + k0 + lk)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y;
      Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], w[TN];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + 4]);
      *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * TN]);
      *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * TN + 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      float* c = C + (size_t)m * ldc + n;
      *c = ACCUM ? (*c + v) : v;
    }
  }
}

int sgemm(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int M, int N,
                 int K, bool relu_a, bool accum, cudaStream_t s) {
  if (M == 0) return PNR_OK;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  prof_before(s);
  if (relu_a && accum) k_sgemm_nt<true, true><<<grid, kGemmThreads, 0, s>>>(A, lda, W, bias, C, ldc, M, N, K);
  else if (relu_a) k_sgemm_nt<true, false><<<grid, kGemmThreads, 0, s>>>(A, lda, W, bias, C, ldc, M, N, K);
  else if (accum) k_sgemm_nt<false, true><<<grid, kGemmThreads, 0, s>>>(A, lda, W, bias, C, ldc, M, N, K);
  else k_sgemm_nt<false, false><<<grid, kGemmThreads, 0, s>>>(A, lda, W, bias, C, ldc, M, N, K);
  prof_after(s);
  PNR_LAUNCH_CHECK();
  return PNR_OK;
}

// lin_in weights [d][d_in] -> zero-padded [d][48] so that K is a multiple of BK
__global__ void k_pad_rows(const float* __restrict__ src, float* __restrict__ dst, int rows, int k_src, int k_dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * k_dst) return;
  int r = i / k_dst, k = i % k_dst;
  dst[i] = (k < k_src) ? src[r * k_src + k] : 0.f;
}

// util.combine_interleaved(average): mean over NS adjacent rows (sum, then divide: torch.mean)
__global__ void k_view_mean(const float* __restrict__ X, float* __restrict__ Y, int64_t n_pts, int NS, int d) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts * d) return;
  int64_t p = i / d;
  int c = (int)(i - p * d);
  float s = X[(p * NS) * d + c];
  for (int v = 1; v < NS; ++v) s += X[(p * NS + v) * d + c];
  Y[i] = s / (float)NS;
}

// lin_out(relu(x)) + sigmoid / relu (resnetfc.py:183, models.py:260-264).  One warp per point.
__global__ void k_lin_out(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b,
                          float* __restrict__ out, int64_t n_pts, int d) {
  int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (p >= n_pts) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k = lane; k < d; k += 32) {
    float x = fmaxf(X[p * d + k], 0.f);
    a0 = fmaf(x, W[0 * d + k], a0);
    a1 = fmaf(x, W[1 * d + k], a1);
    a2 = fmaf(x, W[2 * d + k], a2);
    a3 = fmaf(x, W[3 * d + k], a3);
  }
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  if (lane == 0) {
    float4 r;
    r.x = 1.0f / (1.0f + expf(-(a0 + b[0])));
    r.y = 1.0f / (1.0f + expf(-(a1 + b[1])));
    r.z = 1.0f / (1.0f + expf(-(a2 + b[2])));
    r.w = fmaxf(a3 + b[3], 0.f);
    reinterpret_cast<float4*>(out)[p] = r;
  }
}

static int64_t simt_chunk_points(const PnrScene& sc, int64_t total_points) {
  int64_t c = 32768 / sc.NS;  // rows per chunk <= 32768
  if (c > total_points) c = total_points;
  return c < 1 ? 1 : c;
}

size_t simt_workspace_bytes(const PnrScene& sc, const PnrMlp& mlp, int64_t total_points) {
  int64_t cp = simt_chunk_points(sc, total_points);
  int64_t rows = cp * sc.NS;
  size_t b = 0;
  b += align_up((size_t)rows * 48 * 4, 256);                 // feat
  b += align_up((size_t)rows * mlp.d_latent * 4, 256);       // gathered latent
  b += align_up((size_t)rows * mlp.d_hidden * 4, 256) * 2;   // X, H
  b += align_up((size_t)cp * mlp.d_hidden * 4, 256);         // Xm (after the view mean)
  b += align_up((size_t)mlp.d_hidden * 48 * 4, 256);         // padded lin_in
  return b + 1024;
}

int simt_field_eval(const PnrScene& sc, const PnrMlp& mlp, const PointSource& src, int64_t total_points,
                    float* out, void* ws, size_t ws_bytes, cudaStream_t s) {
  PNR_CHECK_ARG(mlp.d_in == 42, "SIMT engine expects d_in == 42 (use_xyz, 6-frequency code, viewdirs)");
  PNR_CHECK_ARG(mlp.d_out == 4, "d_out must be 4");
  PNR_CHECK_ARG(mlp.d_hidden % 16 == 0 && mlp.d_latent % 16 == 0, "d_hidden and d_latent must be multiples of 16");
  PNR_CHECK_ARG(mlp.d_latent == sc.C, "latent channel mismatch");
  PNR_CHECK_ARG(mlp.n_blocks <= PNR_MAX_BLOCKS, "too many blocks");
  if (ws_bytes < simt_workspace_bytes(sc, mlp, total_points)) {
    set_error("workspace too small: %zu < %zu", ws_bytes, simt_workspace_bytes(sc, mlp, total_points));
    return PNR_ERR_WORKSPACE;
  }
  const int d = mlp.d_hidden, L = mlp.d_latent, NS = sc.NS;
  const int64_t cp = simt_chunk_points(sc, total_points);
  Arena ar(ws, ws_bytes);
  float* feat = ar.take<float>((size_t)cp * NS * 48);
  float* lat = ar.take<float>((size_t)cp * NS * L);
  float* X = ar.take<float>((size_t)cp * NS * d);
  float* H = ar.take<float>((size_t)cp * NS * d);
  float* Xm = ar.take<float>((size_t)cp * d);
  float* w_in = ar.take<float>((size_t)d * 48);
  k_pad_rows<<<(d * 48 + 255) / 256, 256, 0, s>>>(mlp.lin_in_w, w_in, d, mlp.d_in, 48);
  PNR_LAUNCH_CHECK();

  const int comb = mlp.combine_layer < mlp.n_blocks ? mlp.combine_layer : mlp.n_blocks;
  for (int64_t g0 = 0; g0 < total_points; g0 += cp) {
    const int64_t n = (total_points - g0 < cp) ? (total_points - g0) : cp;
    const int rows = (int)(n * NS);
    int rc = launch_build_rows(sc, src, g0, n, feat, lat, s);
    if (rc) return rc;
    // x = lin_in(z_feature)                                        resnetfc.py:147
    if ((rc = sgemm(feat, 48, w_in, mlp.lin_in_b, X, d, rows, d, 48, false, false, s))) return rc;
    float* cur = X;
    int cur_rows = rows;
    for (int blk = 0; blk < mlp.n_blocks; ++blk) {
      if (blk == comb && mlp.combine_layer < mlp.n_blocks) {         // resnetfc.py:152-172
        if (NS > 1) {
          int64_t tot = n * d;
          k_view_mean<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(X, Xm, n, NS, d);
          PNR_LAUNCH_CHECK();
          cur = Xm;
        }
        cur_rows = (int)n;
      }
      if (blk < comb) {                                             // x = x + lin_z[blk](z)   :175,180
        if ((rc = sgemm(lat, L, mlp.lin_z_w[blk], mlp.lin_z_b[blk], cur, d, cur_rows, d, L, false, true, s)))
          return rc;
      }
      // net = fc_0(relu(x)); x = x + fc_1(relu(net))                 resnetfc.py:53-62
      if ((rc = sgemm(cur, d, mlp.fc0_w[blk], mlp.fc0_b[blk], H, d, cur_rows, d, d, true, false, s))) return rc;
      if ((rc = sgemm(H, d, mlp.fc1_w[blk], mlp.fc1_b[blk], cur, d, cur_rows, d, d, true, true, s))) return rc;
    }
    if (mlp.combine_layer >= mlp.n_blocks && NS > 1) {
      set_error("combine_layer >= n_blocks with NS > 1 is not supported");
      return PNR_ERR_UNSUPPORTED;
    }
    k_lin_out<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(cur, mlp.lin_out_w, mlp.lin_out_b, out + g0 * 4, n, d);
    PNR_LAUNCH_CHECK();
  }
  return PNR_OK;
}

}  // namespace pnr
