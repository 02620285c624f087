// Single-process multi-GPU driver of the ray-sharded render (C ABI: pnr_mgpu_*, include/pnr.h).
//
// Replaces `torch.nn.DataParallel(_RenderWrapper, gpus, dim=1)` (reference src/render/nerf.py:354-371), which on EVERY
// forward call re-broadcasts the whole module (~113 MB + latents) from gpus[0], scatters the rays, runs one Python
// thread per GPU and gathers the outputs.  Here:
//   * the read-only scene state (packed weights, projected maps, cameras) is sent ONCE per change with
//     pnr_mgpu_broadcast: peer copies over NVLink / NVSwitch on per-device streams;
//   * a render call enqueues, from one host thread, per shard: a peer copy of the shard's rays (32 B/ray), ONE fused
//     render launch on that device, and the return of its pixels.  For a single object (SB = 1, every eval script) the
//     final rgb / depth are not copied at all: the kernel's compositing epilogue on GPU i stores them straight into
//     the caller's output tensor on gpus[0] through peer memory (the exchange is 16 B/ray against ~2 GFLOP/ray, so
//     the "collective" is fused away as the epilogue's store target); other outputs and SB > 1 use strided peer copies.
// Shards are torch.chunk pieces of the ray axis, so the gathered ray order equals DataParallel's.  Everything is
// asynchronous; the caller's stream on gpus[0] waits on per-shard events.  No NCCL: inside one process peer access
// is the whole transport (bench.py's one-process-per-GPU launcher uses NCCL through torch.distributed instead).
#include <stdlib.h>

#include <vector>

#include "pnr_common.cuh"

struct PnrMgpu {
  std::vector<int> dev;
  std::vector<cudaStream_t> stream;   // per device (index 0 unused: device 0 work runs on the caller's stream)
  std::vector<cudaEvent_t> done;      // per device
  std::vector<char> peer_to_0;        // device i can address device 0's memory
  cudaEvent_t start;                  // recorded on the caller's stream of device 0
};

namespace pnr {

struct DevGuard {
  int prev;
  DevGuard() { cudaGetDevice(&prev); }
  ~DevGuard() { cudaSetDevice(prev); }
};

// torch.chunk piece i of n over B rays: [a, b)
static void chunk_bounds(int64_t B, int n, int i, int64_t* a, int64_t* b) {
  const int64_t per = (B + n - 1) / n;
  *a = per * i < B ? per * i : B;
  *b = *a + per < B ? *a + per : B;
}

}  // namespace pnr

using namespace pnr;

extern "C" {

int pnr_mgpu_create(const int32_t* device_ids, int32_t n, PnrMgpu** out) {
  PNR_CHECK_ARG(device_ids && out && n >= 1 && n <= 64, "bad device list");
  DevGuard guard;
  PnrMgpu* h = new PnrMgpu();
  h->dev.assign(device_ids, device_ids + n);
  h->stream.assign(n, nullptr);
  h->done.assign(n, nullptr);
  h->peer_to_0.assign(n, 0);
  for (int i = 0; i < n; ++i) {
    if (cudaSetDevice(h->dev[i]) != cudaSuccess) {
      set_error("pnr_mgpu_create: cannot select device %d", h->dev[i]);
      delete h;
      return PNR_ERR_CUDA;
    }
    if (i > 0) cudaStreamCreateWithFlags(&h->stream[i], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->done[i], cudaEventDisableTiming);
    if (i > 0) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, h->dev[i], h->dev[0]);
      if (can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(h->dev[0], 0);
        if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) h->peer_to_0[i] = 1;
        cudaGetLastError();
      }
    }
  }
  cudaSetDevice(h->dev[0]);
  cudaEventCreateWithFlags(&h->start, cudaEventDisableTiming);
  *out = h;
  return PNR_OK;
}

int pnr_mgpu_destroy(PnrMgpu* h) {
  if (!h) return PNR_OK;
  DevGuard guard;
  for (size_t i = 0; i < h->dev.size(); ++i) {
    cudaSetDevice(h->dev[i]);
    if (h->stream[i]) {
      cudaStreamSynchronize(h->stream[i]);
      cudaStreamDestroy(h->stream[i]);
    }
    if (h->done[i]) cudaEventDestroy(h->done[i]);
  }
  cudaSetDevice(h->dev[0]);
  cudaEventDestroy(h->start);
  delete h;
  return PNR_OK;
}

int32_t pnr_mgpu_size(const PnrMgpu* h) { return h ? (int32_t)h->dev.size() : 0; }

int32_t pnr_mgpu_peer_store(const PnrMgpu* h, int32_t i) {
  return (h && i >= 0 && i < (int32_t)h->dev.size()) ? (i == 0 ? 1 : h->peer_to_0[i]) : 0;
}

int pnr_mgpu_broadcast(PnrMgpu* h, const void* src, void* const* dst, size_t bytes, void* const* streams) {
  PNR_CHECK_ARG(h && src && dst, "NULL argument");
  if (bytes == 0) return PNR_OK;
  DevGuard guard;
  const int n = (int)h->dev.size();
  PNR_CUDA(cudaSetDevice(h->dev[0]));
  PNR_CUDA(cudaEventRecord(h->start, streams ? (cudaStream_t)streams[0] : (cudaStream_t)0));      // the source is ready
  for (int i = 1; i < n; ++i) {
    if (!dst[i]) continue;                                  // this device does not take part
    cudaStream_t s = (streams && streams[i]) ? (cudaStream_t)streams[i] : h->stream[i];
    PNR_CUDA(cudaSetDevice(h->dev[i]));
    PNR_CUDA(cudaStreamWaitEvent(s, h->start, 0));
    PNR_CUDA(cudaMemcpyPeerAsync(dst[i], h->dev[i], src, h->dev[0], bytes, s));
    // (renders on device i are enqueued on the same stream, so they are ordered after the copy)
  }
  return PNR_OK;
}

// strided copy of a per-ray tensor: rows = objects, row payload = rays of the shard x `width` floats
static int copy_rows(float* dst, int64_t dst_pitch_f, const float* src, int64_t src_pitch_f, int64_t width_f, int64_t rows,
                     cudaStream_t s) {
  if (!dst || !src || width_f == 0 || rows == 0) return PNR_OK;
  PNR_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_f * 4, src, (size_t)src_pitch_f * 4, (size_t)width_f * 4, (size_t)rows,
                             cudaMemcpyDefault, s));
  return PNR_OK;
}

int pnr_mgpu_render(PnrMgpu* h, const PnrShard* shards, const PnrRenderCfg* cfg, const float* rays0,
                    const PnrRenderOut* out0, int64_t B, void* stream0) {
  PNR_CHECK_ARG(h && shards && cfg && rays0 && out0, "NULL argument");
  PNR_CHECK_ARG(B >= 0, "B must be >= 0");
  DevGuard guard;
  const int n = (int)h->dev.size();
  const int SB = shards[0].scene ? shards[0].scene->SB : 0;
  PNR_CHECK_ARG(SB >= 1, "shard 0 has no scene");
  const int Kc = cfg->n_coarse, K = cfg->n_coarse + cfg->n_fine;
  const bool fine = cfg->n_fine > 0;
  PNR_CUDA(cudaSetDevice(h->dev[0]));
  PNR_CUDA(cudaEventRecord(h->start, (cudaStream_t)stream0));
  int rc = PNR_OK;
  int used = 0;
  for (int i = 0; i < n && rc == PNR_OK; ++i) {
    int64_t a, b;
    chunk_bounds(B, n, i, &a, &b);
    const int64_t Bi = b - a;
    if (Bi <= 0) continue;
    const PnrShard& sh = shards[i];
    PNR_CHECK_ARG(sh.scene && sh.mlp_coarse && sh.noise && sh.workspace, "incomplete shard");
    PNR_CHECK_ARG(sh.scene->SB == SB, "all shards must hold the same objects");
    cudaStream_t s = i == 0 ? (cudaStream_t)stream0 : (sh.stream ? (cudaStream_t)sh.stream : h->stream[i]);
    PNR_CUDA(cudaSetDevice(h->dev[i]));
    if (i > 0) PNR_CUDA(cudaStreamWaitEvent(s, h->start, 0));
    // rays of the shard: in place on device 0 for one object, else a (strided) peer copy into the shard's stage
    const float* rays_i = rays0 + a * 8;
    if (i > 0 || SB > 1) {
      PNR_CHECK_ARG(sh.rays_stage, "shard needs a ray staging buffer");
      if ((rc = copy_rows(sh.rays_stage, Bi * 8, rays0 + a * 8, B * 8, Bi * 8, SB, s))) break;
      rays_i = sh.rays_stage;
    }
    // outputs: final rgb / depth of a single object go straight into the caller's tensors (peer stores from the
    // kernel's epilogue); everything else is rendered locally and copied back with the object stride
    const bool direct = SB == 1 && (i == 0 || h->peer_to_0[i]);
    PnrRenderOut o = sh.stage;
    float* best_rgb0 = fine ? out0->rgb_fine : out0->rgb_coarse;
    float* best_dep0 = fine ? out0->depth_fine : out0->depth_coarse;
    if (direct) {
      if (fine) {
        if (best_rgb0) o.rgb_fine = best_rgb0 + a * 3;
        if (best_dep0) o.depth_fine = best_dep0 + a;
      } else {
        if (best_rgb0) o.rgb_coarse = best_rgb0 + a * 3;
        if (best_dep0) o.depth_coarse = best_dep0 + a;
      }
    }
    if (!out0->weights_coarse) o.weights_coarse = nullptr;
    if (!out0->z_coarse) o.z_coarse = nullptr;
    if (!out0->weights_fine) o.weights_fine = nullptr;
    if (!out0->z_fine) o.z_fine = nullptr;
    rc = pnr_render(sh.scene, sh.mlp_coarse, sh.mlp_fine, cfg, rays_i, sh.noise, &o, Bi, sh.workspace, sh.workspace_bytes, s);
    if (rc) break;
    struct Item { float* dst; const float* src; int64_t w; };
    const Item items[] = {
        {out0->rgb_coarse, o.rgb_coarse, 3}, {out0->depth_coarse, o.depth_coarse, 1},
        {out0->weights_coarse, o.weights_coarse, Kc}, {out0->z_coarse, o.z_coarse, Kc},
        {fine ? out0->rgb_fine : nullptr, o.rgb_fine, 3}, {fine ? out0->depth_fine : nullptr, o.depth_fine, 1},
        {fine ? out0->weights_fine : nullptr, o.weights_fine, K}, {fine ? out0->z_fine : nullptr, o.z_fine, K}};
    for (const Item& it : items) {
      if (!it.dst || !it.src) continue;
      if (it.src == it.dst + a * it.w) continue;   // written in place by the kernel
      if ((rc = copy_rows(it.dst + a * it.w, B * it.w, it.src, Bi * it.w, Bi * it.w, SB, s))) break;
    }
    if (rc) break;
    if (i > 0) PNR_CUDA(cudaEventRecord(h->done[i], s));
    used = i + 1;
  }
  PNR_CUDA(cudaSetDevice(h->dev[0]));
  for (int i = 1; i < used; ++i) PNR_CUDA(cudaStreamWaitEvent((cudaStream_t)stream0, h->done[i], 0));
  return rc;
}

}  // extern "C"
