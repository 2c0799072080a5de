// Event -> voxel-grid kernels (HBM-bound scatter).
//
// Replaces events_to_voxel_torch / events_to_image_torch
// (reference: events_contrast_maximization/utils/event_utils.py:466-509, :330-376) for a whole
// sequence of windows per launch, writing straight into the zero-padded grid that Croper.pad
// (utils_func/inference_utils.py:104-111) would produce.
//
// Arithmetic contract (bit-exact bin indices):
//   dt = ts[last] - ts[first];  tn = ((t - ts[first]) / dt) * (B-1)      -- fp32, that op order
//   for every bin b: w = p * max(0, 1 - |tn - b|);  grid[b, int(y), int(x)] += w
// Only bins floor(tn) and floor(tn)+1 can receive a non-zero weight, so only those are touched.
// A NaN tn (dt == 0) poisons every bin of the touched pixel exactly like the reference does.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace bde {

__device__ __forceinline__ float tnorm(float t, float t0, float dt, float bm1) {
  // __f*_rn intrinsics are never contracted into FMAs and ignore fast-math flags
  return __fmul_rn(__fdiv_rn(__fsub_rn(t, t0), dt), bm1);
}

struct EventContrib {
  int b0;       // left bin (floor(tn)); -1 -> NaN event (all bins)
  float w0, w1; // weights for b0 and b0+1
};

__device__ __forceinline__ EventContrib contrib(float t, float p, float t0, float dt, float bm1, int bins) {
  EventContrib c;
  float tn = tnorm(t, t0, dt, bm1);
  if (!(tn == tn)) {  // NaN
    c.b0 = -1;
    c.w0 = tn;
    c.w1 = tn;
    return c;
  }
  float fl = floorf(tn);
  int b0 = (int)fl;
  b0 = max(0, min(b0, bins - 1));
  float fb0 = (float)b0;
  // w = p * max(0, 1 - |tn - b|), evaluated exactly as written for both bins
  c.w0 = __fmul_rn(p, fmaxf(0.0f, __fsub_rn(1.0f, fabsf(__fsub_rn(tn, fb0)))));
  c.w1 = __fmul_rn(p, fmaxf(0.0f, __fsub_rn(1.0f, fabsf(__fsub_rn(tn, fb0 + 1.0f)))));
  c.b0 = b0;
  return c;
}

// ------------------------------------------------------------------------------------------------
// Algorithm 1: one CTA per (row band, window).  The CTA scans the window's events once (coalesced,
// float4-vectorised; re-reads by other bands hit L2), accumulates the events that fall into its
// band in shared memory, then writes its band of the padded output exactly once, fully coalesced.
// No memset pass and no global atomics: HBM traffic == algorithmic bytes.
// Lanes of a warp that hit the same cell are combined before the shared-memory atomic.
// ------------------------------------------------------------------------------------------------
template <bool kAggregate>
__device__ __forceinline__ void smem_accumulate(float* tile, int idx0, int idx1, float w0, float w1, bool live) {
  if (kAggregate) {
    // warp-aggregated: lanes targeting the same cell elect a leader which adds the group's sum
    unsigned active = __activemask();
    int key = live ? idx0 : -1 - (int)(threadIdx.x & 31);
    unsigned peers = __match_any_sync(active, key);
    int leader = __ffs(peers) - 1;
    int lane = threadIdx.x & 31;
    if (__popc(peers) > 1) {
      float s0 = 0.f, s1 = 0.f;
      unsigned rem = peers;
      // every peer walks the same peer list so the shuffles are convergent within the group
      while (rem) {
        int src = __ffs(rem) - 1;
        rem &= rem - 1;
        s0 += __shfl_sync(peers, w0, src);
        s1 += __shfl_sync(peers, w1, src);
      }
      w0 = s0;
      w1 = s1;
    }
    if (live && lane == leader) {
      if (w0 != 0.0f) atomicAdd(tile + idx0, w0);
      if (idx1 >= 0 && w1 != 0.0f) atomicAdd(tile + idx1, w1);
    }
  } else {
    if (live) {
      if (w0 != 0.0f) atomicAdd(tile + idx0, w0);
      if (idx1 >= 0 && w1 != 0.0f) atomicAdd(tile + idx1, w1);
    }
  }
}

template <bool kAggregate>
__global__ void __launch_bounds__(512) voxel_band_kernel(
    const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ ts,
    const float* __restrict__ ps, const int64_t* __restrict__ offsets, int bins, int H, int W,
    int pad_top, int pad_left, int Hp, int Wp, int band_rows, int vec_ok, float* __restrict__ out,
    size_t win_stride, int* oob_count, int min_events) {
  extern __shared__ float tile[];  // [bins][band_rows][W]
  const int win = blockIdx.y;
  const int r0 = blockIdx.x * band_rows;           // first padded row of this band
  const int r1 = min(Hp, r0 + band_rows);
  const int y0 = r0 - pad_top;                     // sensor row of tile row 0 (may be negative)
  const int plane = band_rows * W;
  const int tile_elems = bins * plane;
  for (int i = threadIdx.x; i < tile_elems; i += blockDim.x) tile[i] = 0.0f;
  __syncthreads();

  const int64_t ea = offsets[win], eb = offsets[win + 1];
  if (eb - ea >= min_events) {
    const float t0 = ts[ea];
    const float dt = __fsub_rn(ts[eb - 1], t0);
    const float bm1 = (float)(bins - 1);
    int oob = 0;
    auto one = [&](float x, float y, float t, float p) {
      int xi = (int)x, yi = (int)y;  // truncation == .long() (event_utils.py:371-374)
      bool inside = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H);
      if (!inside && blockIdx.x == 0) oob++;
      int ty = yi - y0;
      bool live = inside & (ty >= 0) & (yi < r1 - pad_top);
      EventContrib c = contrib(t, p, t0, dt, bm1, bins);
      if (c.b0 < 0) {  // NaN event: every bin of the pixel becomes NaN
        if (live)
          for (int b = 0; b < bins; ++b) atomicAdd(tile + b * plane + ty * W + xi, c.w0);
        live = false;
      }
      int idx0 = live ? (c.b0 * plane + ty * W + xi) : 0;
      int idx1 = (live && c.b0 + 1 < bins) ? idx0 + plane : -1;
      smem_accumulate<kAggregate>(tile, idx0, idx1, c.w0, c.w1, live);
    };
    // scalar head up to 16-byte alignment, float4 body, scalar tail
    int64_t body_a = vec_ok ? min(eb, (ea + 3) & ~(int64_t)3) : eb;
    int64_t body_b = vec_ok ? max(body_a, eb & ~(int64_t)3) : eb;
    // head + tail: at most 6 events, handled by the first lanes; everyone joins the warp-collective
    int64_t n_edge = (body_a - ea) + (eb - body_b);
    for (int64_t base = 0; base < n_edge; base += blockDim.x) {
      int64_t i = base + threadIdx.x;
      bool valid = i < n_edge;
      int64_t e = valid ? (i < body_a - ea ? ea + i : body_b + (i - (body_a - ea))) : ea;
      if (valid) one(xs[e], ys[e], ts[e], ps[e]);
    }
    const int64_t nvec = (body_b - body_a) >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(xs + body_a);
    const float4* y4 = reinterpret_cast<const float4*>(ys + body_a);
    const float4* t4 = reinterpret_cast<const float4*>(ts + body_a);
    const float4* p4 = reinterpret_cast<const float4*>(ps + body_a);
    for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
      float4 x = __ldg(x4 + v), y = __ldg(y4 + v), t = __ldg(t4 + v), p = __ldg(p4 + v);
      one(x.x, y.x, t.x, p.x);
      one(x.y, y.y, t.y, p.y);
      one(x.z, y.z, t.z, p.z);
      one(x.w, y.w, t.w, p.w);
    }
    if (oob_count != nullptr && oob > 0) atomicAdd(oob_count, oob);
  }
  __syncthreads();

  // coalesced write of the band, padding included
  float* dst = out + (size_t)win * win_stride;
  const int rows = r1 - r0;
  const int total = bins * rows * Wp;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    int col = i % Wp;
    int rr = (i / Wp) % rows;
    int b = i / (Wp * rows);
    int y = r0 + rr - pad_top, x = col - pad_left;
    float v = 0.0f;
    if (y >= 0 && y < H && x >= 0 && x < W) v = tile[b * plane + (y - y0) * W + x];
    dst[((size_t)b * Hp + (r0 + rr)) * Wp + col] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Algorithm 2 (default): global float reductions (RED.ADD.F32, executed in place by L2) into a zeroed grid.
// Event sources: the loader format (four float32 streams, 16 B / event) or the reference's ON-DISK format
// (events_contrast_maximization/tools/event_packagers.py:44-47: xs, ys int16, ts float64 seconds, ps bool; 13 B / event)
// with the loader's conversions done in registers (data_loader/h5_dataset.py:222-225, :414):
//   x, y -> float32 (exact), t -> float32(ts - ts[window start]) (float64 subtraction, then the cast), p -> 2 p - 1.
// Each thread takes FOUR consecutive events per iteration with 128-bit (64- / 32-bit for the narrow types) loads;
// kAgg adds a warp-level pre-reduction: lanes whose event hits the same (pixel, left bin) cell are found with
// __match_any_sync and only the group leader issues the reductions -- a win for spatially clustered streams, a loss
// for uniform ones (tools/voxel_probe.py measures both), hence selectable.
// ------------------------------------------------------------------------------------------------
struct EvSrcF32 {
  const float *x, *y, *t, *p;
  __device__ __forceinline__ bool vec_ok() const {
    return ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)t) | ((uintptr_t)p)) & 15) == 0;
  }
  __device__ __forceinline__ void window(int64_t ea, int64_t eb, float& dt) const { t0 = t[ea]; dt = __fsub_rn(t[eb - 1], t0); }
  __device__ __forceinline__ void load1(int64_t e, float& xx, float& yy, float& tt, float& pp) const {
    xx = __ldg(x + e); yy = __ldg(y + e); tt = __fsub_rn(__ldg(t + e), t0); pp = __ldg(p + e);
  }
  __device__ __forceinline__ void load4(int64_t e, float (&xx)[4], float (&yy)[4], float (&tt)[4], float (&pp)[4]) const {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + e)), b = __ldg(reinterpret_cast<const float4*>(y + e));
    const float4 c = __ldg(reinterpret_cast<const float4*>(t + e)), d = __ldg(reinterpret_cast<const float4*>(p + e));
    xx[0] = a.x; xx[1] = a.y; xx[2] = a.z; xx[3] = a.w;
    yy[0] = b.x; yy[1] = b.y; yy[2] = b.z; yy[3] = b.w;
    tt[0] = __fsub_rn(c.x, t0); tt[1] = __fsub_rn(c.y, t0); tt[2] = __fsub_rn(c.z, t0); tt[3] = __fsub_rn(c.w, t0);
    pp[0] = d.x; pp[1] = d.y; pp[2] = d.z; pp[3] = d.w;
  }
  mutable float t0;
};

struct EvSrcRaw {
  const int16_t *x, *y;
  const double* t;
  const uint8_t* p;
  __device__ __forceinline__ bool vec_ok() const {
    return ((((uintptr_t)x) | ((uintptr_t)y)) & 7) == 0 && (((uintptr_t)t) & 15) == 0 && (((uintptr_t)p) & 3) == 0;
  }
  // h5_dataset.py:222-225: ts_0 = ts[0] (float64); ts = (ts - ts_0).astype(float32); then event_utils.py:489 dt = ts[-1] - ts[0]
  __device__ __forceinline__ void window(int64_t ea, int64_t eb, float& dt) const {
    t0 = t[ea];
    dt = __fsub_rn(__double2float_rn(__dsub_rn(t[eb - 1], t0)), 0.0f);
  }
  __device__ __forceinline__ float rel(double tv) const { return __fsub_rn(__double2float_rn(__dsub_rn(tv, t0)), 0.0f); }
  __device__ __forceinline__ void load1(int64_t e, float& xx, float& yy, float& tt, float& pp) const {
    xx = (float)x[e]; yy = (float)y[e]; tt = rel(t[e]); pp = p[e] ? 1.0f : -1.0f;     // h5_dataset.py:414 ps * 2.0 - 1.0
  }
  __device__ __forceinline__ void load4(int64_t e, float (&xx)[4], float (&yy)[4], float (&tt)[4], float (&pp)[4]) const {
    const short4 a = __ldg(reinterpret_cast<const short4*>(x + e)), b = __ldg(reinterpret_cast<const short4*>(y + e));
    const double2 c0 = __ldg(reinterpret_cast<const double2*>(t + e)), c1 = __ldg(reinterpret_cast<const double2*>(t + e + 2));
    const uchar4 d = __ldg(reinterpret_cast<const uchar4*>(p + e));
    xx[0] = (float)a.x; xx[1] = (float)a.y; xx[2] = (float)a.z; xx[3] = (float)a.w;
    yy[0] = (float)b.x; yy[1] = (float)b.y; yy[2] = (float)b.z; yy[3] = (float)b.w;
    tt[0] = rel(c0.x); tt[1] = rel(c0.y); tt[2] = rel(c1.x); tt[3] = rel(c1.y);
    pp[0] = d.x ? 1.0f : -1.0f; pp[1] = d.y ? 1.0f : -1.0f; pp[2] = d.z ? 1.0f : -1.0f; pp[3] = d.w ? 1.0f : -1.0f;
  }
  mutable double t0;
};

template <typename Src, bool kAgg>
__global__ void __launch_bounds__(256, 4) voxel_atomic_kernel(
    const Src src, const int64_t* __restrict__ offsets, int win0, int bins, int H, int W, int pad_top, int pad_left, int Hp, int Wp,
    float* __restrict__ out, size_t win_stride, int* oob_count, int min_events, const float* __restrict__ hot_mask,
    int zero_win0, int zero_n, unsigned long long grid_elems4) {
  // Chunk pipeline (launch_atomic): this launch reduces into the windows [win0, win0 + gridDim.y), which the PREVIOUS
  // launch (or a memset, for the first chunk) zeroed, and zeroes the windows [zero_win0, zero_win0 + zero_n) of the NEXT
  // chunk.  Launched with programmatic stream serialisation, so everything up to pdl_wait() -- the window bounds, t0 / dt
  // and the first four events of every thread, all inputs that no kernel of the chain writes -- overlaps the tail of the
  // previous launch; the reductions and the zero stores come after pdl_wait() (previous grid complete, its zeroes visible).
  pdl_trigger();
  const int win = win0 + blockIdx.y;
  const int64_t ea = offsets[win], eb = offsets[win + 1];
  // loader contract (h5_dataset.py:219-221): windows with fewer than `min_events` events give an all-zero grid
  const bool live_win = !(eb - ea < (int64_t)min_events || eb <= ea);
  float dt = 1.0f;
  if (live_win) src.window(ea, eb, dt);
  const float bm1 = (float)(bins - 1);
  float* dst = out + (size_t)win * win_stride;
  const size_t plane = (size_t)Hp * Wp;
  int oob = 0;
  // `live` = false lanes only take part in the warp votes of the aggregated form
  auto one = [&](float x, float y, float t, float p, bool live) {
    const int xi = (int)x, yi = (int)y;  // truncation == .long() (event_utils.py:371-374)
    const bool inside = live && xi >= 0 && xi < W && yi >= 0 && yi < H;
    if (live && !inside) oob++;
    // hot-pixel mask (h5_dataset.py:163-172,364: voxel * mask with mask in {0, 1}): masked pixels receive nothing
    const bool keep = inside && (hot_mask == nullptr || __ldg(hot_mask + (size_t)yi * W + xi) != 0.0f);
    // tn = ((t - t0) / dt) * (B - 1); the sources hand over t - t0
    EventContrib c;
    {
      const float tn = __fmul_rn(__fdiv_rn(t, dt), bm1);
      if (!(tn == tn)) {
        c.b0 = -1; c.w0 = tn; c.w1 = tn;
      } else {
        int b0 = (int)floorf(tn);
        b0 = max(0, min(b0, bins - 1));
        const float fb0 = (float)b0;
        c.w0 = __fmul_rn(p, fmaxf(0.0f, __fsub_rn(1.0f, fabsf(__fsub_rn(tn, fb0)))));
        c.w1 = __fmul_rn(p, fmaxf(0.0f, __fsub_rn(1.0f, fabsf(__fsub_rn(tn, fb0 + 1.0f)))));
        c.b0 = b0;
      }
    }
    const size_t pix = keep ? (size_t)(yi + pad_top) * Wp + (xi + pad_left) : 0;
    if (keep && c.b0 < 0) {  // NaN event (dt == 0): every bin of the pixel becomes NaN, as in the reference
      for (int b = 0; b < bins; ++b) atomicAdd(dst + b * plane + pix, c.w0);
    }
    bool add = keep && c.b0 >= 0;
    float w0 = c.w0, w1 = c.w1;
    if (kAgg) {
      const int lane = threadIdx.x & 31;
      const int key = add ? (int)(c.b0 * plane + pix) : -1 - lane;
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      if (__popc(peers) > 1) {
        float s0 = 0.f, s1 = 0.f;
        unsigned rem = peers;
        while (rem) {
          const int srcl = __ffs(rem) - 1;
          rem &= rem - 1;
          s0 += __shfl_sync(peers, w0, srcl);
          s1 += __shfl_sync(peers, w1, srcl);
        }
        w0 = s0; w1 = s1;
        add = add && lane == __ffs(peers) - 1;
      }
    }
    if (add) {
      if (w0 != 0.0f) atomicAdd(dst + c.b0 * plane + pix, w0);
      if (c.b0 + 1 < bins && w1 != 0.0f) atomicAdd(dst + (c.b0 + 1) * plane + pix, w1);
    }
  };
  const bool vec = src.vec_ok();
  // scalar head up to a multiple of 4, vector body, scalar tail
  const int64_t body_a = vec ? min(eb, (ea + 3) & ~(int64_t)3) : eb;
  const int64_t body_b = vec ? max(body_a, eb & ~(int64_t)3) : eb;
  const int64_t n_edge = live_win ? (body_a - ea) + (eb - body_b) : 0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  const int64_t nvec = live_win ? (body_b - body_a) >> 2 : 0;
  float x[4], y[4], t[4], p[4];
  if (nvec > 0 && (kAgg || tid < nvec)) src.load4(body_a + 4 * (tid < nvec ? tid : 0), x, y, t, p);
  pdl_wait();
  // (warp-uniform trip counts so that the aggregated form's votes stay convergent)
  for (int64_t base = 0; base < n_edge; base += nthr) {
    const int64_t i = base + tid;
    const bool valid = i < n_edge;
    const int64_t e = valid ? (i < body_a - ea ? ea + i : body_b + (i - (body_a - ea))) : ea;
    float x1, y1, t1, p1;
    src.load1(e, x1, y1, t1, p1);
    if (kAgg || valid) one(x1, y1, t1, p1, valid);
  }
  for (int64_t base = 0; base < nvec; base += nthr) {
    const int64_t v = base + tid;
    const bool valid = v < nvec;
    if (!kAgg && !valid) break;
    if (base != 0) src.load4(body_a + 4 * (valid ? v : 0), x, y, t, p);   // the first group was loaded ahead of pdl_wait()
#pragma unroll
    for (int k = 0; k < 4; ++k) one(x[k], y[k], t[k], p[k], valid);
  }
  // zero the next chunk's grids (16-byte stores, the whole launch strides over them)
  if (zero_n > 0) {
    const unsigned long long gtid = ((unsigned long long)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned long long gn = (unsigned long long)gridDim.y * gridDim.x * blockDim.x;
    const unsigned long long total = (unsigned long long)zero_n * grid_elems4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned long long i = gtid; i < total; i += gn) {
      const unsigned long long w = i / grid_elems4, r = i - w * grid_elems4;
      reinterpret_cast<float4*>(out + (size_t)(zero_win0 + w) * win_stride)[r] = z;
    }
  }
  if (oob_count != nullptr && oob > 0) atomicAdd(oob_count, oob);
}


// ------------------------------------------------------------------------------------------------
// Algorithm 3: one thread-block CLUSTER per window; the whole voxel grid lives in the distributed
// shared memory of the cluster (CTA r owns the sensor rows [r * band, (r + 1) * band) of every bin).
// Every event is read from HBM exactly once -- CTA r scans the r-th share of the window -- and is
// added into the owning CTA's tile with a shared::cluster reduction (red.add.f32 over DSMEM); after a
// cluster barrier each CTA streams its band of the zero-padded output with 16-byte stores.
// HBM traffic == algorithmic bytes (16 N + 4 B Hp Wp per window) and, unlike the row-band algorithm,
// no CTA re-scans events that belong to other bands.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vx_mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void vx_red_add(uint32_t cluster_addr, float v) {
  asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void vx_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(512, 1) voxel_cluster_kernel(
    const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ ts,
    const float* __restrict__ ps, const int64_t* __restrict__ offsets, int bins, int H, int W,
    int pad_top, int pad_left, int Hp, int Wp, int band_rows, int nc, int vec_ok, float* __restrict__ out,
    size_t win_stride, int* oob_count, int scan_all, int min_events) {
  extern __shared__ __align__(16) float tile[];  // [bins][band_rows][W] of this CTA's band
  const int win = blockIdx.x / nc;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // scan_all: every CTA reads ALL events of the window (L2 hits after the first CTA) and accumulates only those of its
  // own rows with local shared-memory atomics -- no remote reductions, at the price of nc x the event reads from L2
  const int plane = band_rows * W;
  const int tile_elems = bins * plane;
  {
    float4* t4 = reinterpret_cast<float4*>(tile);
    const int n4 = tile_elems >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = (n4 << 2) + threadIdx.x; i < tile_elems; i += blockDim.x) tile[i] = 0.f;
  }
  vx_cluster_sync();  // every tile of the cluster is zeroed before anyone adds into it

  const int64_t ea = offsets[win], eb = offsets[win + 1];
  if (eb - ea >= min_events) {
    const float t0 = ts[ea];
    const float dt = __fsub_rn(ts[eb - 1], t0);
    const float bm1 = (float)(bins - 1);
    const uint32_t tile_u32 = (uint32_t)__cvta_generic_to_shared(tile);
    int oob = 0;
    auto one = [&](float x, float y, float t, float p) {
      const int xi = (int)x, yi = (int)y;  // truncation == .long() (event_utils.py:371-374)
      if (xi < 0 || xi >= W || yi < 0 || yi >= H) {
        if (!scan_all || rank == 0) oob++;
        return;
      }
      const int owner = yi / band_rows;
      if (scan_all) {
        if (owner != (int)rank) return;
        const EventContrib c = contrib(t, p, t0, dt, bm1, bins);
        float* cellp = tile + (yi - owner * band_rows) * W + xi;
        if (c.b0 < 0) {
          for (int b = 0; b < bins; ++b) atomicAdd(cellp + b * plane, c.w0);
          return;
        }
        if (c.w0 != 0.0f) atomicAdd(cellp + c.b0 * plane, c.w0);
        if (c.b0 + 1 < bins && c.w1 != 0.0f) atomicAdd(cellp + (c.b0 + 1) * plane, c.w1);
        return;
      }
      const EventContrib c = contrib(t, p, t0, dt, bm1, bins);
      const uint32_t cell = vx_mapa(tile_u32 + (uint32_t)(((yi - owner * band_rows) * W + xi) * 4), (uint32_t)owner);
      if (c.b0 < 0) {  // NaN event: every bin of the pixel becomes NaN
        for (int b = 0; b < bins; ++b) vx_red_add(cell + (uint32_t)(b * plane * 4), c.w0);
        return;
      }
      if (c.w0 != 0.0f) vx_red_add(cell + (uint32_t)(c.b0 * plane * 4), c.w0);
      if (c.b0 + 1 < bins && c.w1 != 0.0f) vx_red_add(cell + (uint32_t)((c.b0 + 1) * plane * 4), c.w1);
    };
    // this CTA's share of the window: scalar head / tail around a 16-byte aligned float4 body
    const int64_t n = eb - ea;
    const int64_t sa = scan_all ? ea : ea + (n * rank) / nc, sb_ = scan_all ? eb : ea + (n * (rank + 1)) / nc;
    int64_t body_a = vec_ok ? min(sb_, (sa + 3) & ~(int64_t)3) : sb_;
    int64_t body_b = vec_ok ? max(body_a, sb_ & ~(int64_t)3) : sb_;
    const int64_t n_edge = (body_a - sa) + (sb_ - body_b);
    for (int64_t i = threadIdx.x; i < n_edge; i += blockDim.x) {
      const int64_t e = i < body_a - sa ? sa + i : body_b + (i - (body_a - sa));
      one(xs[e], ys[e], ts[e], ps[e]);
    }
    const int64_t nvec = (body_b - body_a) >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(xs + body_a);
    const float4* y4 = reinterpret_cast<const float4*>(ys + body_a);
    const float4* t4 = reinterpret_cast<const float4*>(ts + body_a);
    const float4* p4 = reinterpret_cast<const float4*>(ps + body_a);
    for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
      const float4 x = __ldg(x4 + v), y = __ldg(y4 + v), t = __ldg(t4 + v), p = __ldg(p4 + v);
      one(x.x, y.x, t.x, p.x);
      one(x.y, y.y, t.y, p.y);
      one(x.z, y.z, t.z, p.z);
      one(x.w, y.w, t.w, p.w);
    }
    if (oob_count != nullptr && oob > 0) atomicAdd(oob_count, oob);
  }
  vx_cluster_sync();  // all reductions into this CTA's tile have landed

  // ---- coalesced write of this CTA's rows of the padded grid (first / last CTA add the padding rows) ----
  float* dst = out + (size_t)win * win_stride;
  const int y_lo = min(H, (int)rank * band_rows), y_hi = min(H, y_lo + band_rows);   // sensor rows owned
  const int r_lo = rank == 0 ? 0 : y_lo + pad_top;
  const int r_hi = (int)rank == nc - 1 ? Hp : y_hi + pad_top;
  const int nrows = max(0, r_hi - r_lo);
  const int wq = Wp >> 2;   // Wp % 4 == 0 (checked by the host)
  for (int b = 0; b < bins; ++b) {
    const float* tb = tile + b * plane;
    for (int i = threadIdx.x; i < nrows * wq; i += blockDim.x) {
      const int rr = i / wq, c4 = (i - rr * wq) << 2;
      const int r = r_lo + rr, y = r - pad_top;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= y_lo && y < y_hi) {
        const float* trow = tb + (y - y_lo) * W;
        const int x = c4 - pad_left;
        v.x = (x >= 0 && x < W) ? trow[x] : 0.f;
        v.y = (x + 1 >= 0 && x + 1 < W) ? trow[x + 1] : 0.f;
        v.z = (x + 2 >= 0 && x + 2 < W) ? trow[x + 2] : 0.f;
        v.w = (x + 3 >= 0 && x + 3 < W) ? trow[x + 3] : 0.f;
      }
      *reinterpret_cast<float4*>(dst + ((size_t)b * Hp + r) * Wp + c4) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Algorithm 4: one 8-CTA cluster per window, no shared-memory tile.  Each CTA first ZEROES its slice of the window's
// padded grid in global memory (16-byte stores; the slice stays in L2), the cluster barrier (release / acquire at
// cluster scope) orders those stores before any reduction, then each CTA adds its share of the window's events with
// global RED.ADD.F32, which L2 executes in place.  One launch per sequence, every event and every grid cell touched
// once; measured fastest on B200 (L2 float atomics outrun both the shared-memory tile with warp aggregation and the
// distributed-shared-memory reductions of algorithm 3).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) voxel_cluster_red_kernel(
    const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ ts,
    const float* __restrict__ ps, const int64_t* __restrict__ offsets, int bins, int H, int W,
    int pad_top, int pad_left, int Hp, int Wp, int nc, int vec_ok, float* __restrict__ out, size_t win_stride,
    int* oob_count, int min_events) {
  const int win = blockIdx.x / nc;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  float* dst = out + (size_t)win * win_stride;
  const size_t plane = (size_t)Hp * Wp;
  {
    // zero this CTA's 1 / nc of the grid (grid size is a multiple of 4 floats: Wp % 4 == 0)
    const size_t n4 = (size_t)bins * plane / 4;
    const size_t a = n4 * rank / nc, b = n4 * (rank + 1) / nc;
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (size_t i = a + threadIdx.x; i < b; i += blockDim.x) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __threadfence();
  vx_cluster_sync();
  const int64_t ea = offsets[win], eb = offsets[win + 1];
  if (eb - ea < min_events) return;
  const float t0 = ts[ea];
  const float dt = __fsub_rn(ts[eb - 1], t0);
  const float bm1 = (float)(bins - 1);
  int oob = 0;
  auto one = [&](float x, float y, float t, float p) {
    const int xi = (int)x, yi = (int)y;  // truncation == .long() (event_utils.py:371-374)
    if (xi < 0 || xi >= W || yi < 0 || yi >= H) {
      oob++;
      return;
    }
    const EventContrib c = contrib(t, p, t0, dt, bm1, bins);
    float* cell = dst + (size_t)(yi + pad_top) * Wp + (xi + pad_left);
    if (c.b0 < 0) {
      for (int b = 0; b < bins; ++b) atomicAdd(cell + b * plane, c.w0);
      return;
    }
    if (c.w0 != 0.0f) atomicAdd(cell + c.b0 * plane, c.w0);
    if (c.b0 + 1 < bins && c.w1 != 0.0f) atomicAdd(cell + (c.b0 + 1) * plane, c.w1);
  };
  const int64_t n = eb - ea;
  const int64_t sa = ea + (n * rank) / nc, sb_ = ea + (n * (rank + 1)) / nc;
  int64_t body_a = vec_ok ? min(sb_, (sa + 3) & ~(int64_t)3) : sb_;
  int64_t body_b = vec_ok ? max(body_a, sb_ & ~(int64_t)3) : sb_;
  const int64_t n_edge = (body_a - sa) + (sb_ - body_b);
  for (int64_t i = threadIdx.x; i < n_edge; i += blockDim.x) {
    const int64_t e = i < body_a - sa ? sa + i : body_b + (i - (body_a - sa));
    one(xs[e], ys[e], ts[e], ps[e]);
  }
  const int64_t nvec = (body_b - body_a) >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(xs + body_a);
  const float4* y4 = reinterpret_cast<const float4*>(ys + body_a);
  const float4* t4 = reinterpret_cast<const float4*>(ts + body_a);
  const float4* p4 = reinterpret_cast<const float4*>(ps + body_a);
  for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
    const float4 x = __ldg(x4 + v), y = __ldg(y4 + v), t = __ldg(t4 + v), p = __ldg(p4 + v);
    one(x.x, y.x, t.x, p.x);
    one(x.y, y.y, t.y, p.y);
    one(x.z, y.z, t.z, p.z);
    one(x.w, y.w, t.w, p.w);
  }
  if (oob_count != nullptr && oob > 0) atomicAdd(oob_count, oob);
}

// planar fp32 [T, bins, Hp, Wp] -> NHWC [T, Hp, Wp, c_pad] (zero-padded channels)
template <typename T>
__global__ void pack_voxel_kernel(const float* __restrict__ vox, int bins, size_t plane, int c_pad,
                                  T* __restrict__ out, size_t total_pix) {
  size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over T*Hp*Wp
  if (pix >= total_pix) return;
  size_t t = pix / plane, p = pix % plane;
  const float* src = vox + t * bins * plane + p;
  T* dst = out + pix * c_pad;
  for (int c = 0; c < c_pad; ++c) dst[c] = from_f32<T>(c < bins ? __ldg(src + (size_t)c * plane) : 0.0f);
}

}  // namespace bde

using namespace bde;

// Reduction kernel over chunks of windows whose grids fit L2 together: the zeroed lines are still resident when the
// reductions hit them, so the grid goes to HBM once (write-back) instead of zero-write + reduction read + write-back.
// Chunk 0 is zeroed by a memset; every launch zeroes the NEXT chunk's grids after its own reductions have been issued,
// and the launches are chained with programmatic stream serialisation (see the kernel), so that the zero stores (HBM /
// L2 write bandwidth) run under the reductions (L2 atomic units) and the event loads of chunk c + 1 under the tail of
// chunk c.  Round 2 measured memset + kernel back to back: 36 + 64 us per 100 windows of the bench shape.
template <typename Src, bool kAgg>
static cudaError_t launch_voxel_atomic(dim3 grid, cudaStream_t s, const Src& src, const int64_t* offsets, int w0, int num_bins, int H,
                                       int W, int pad_top, int pad_left, int Hp, int Wp, float* out, size_t stride, int* oob_count,
                                       int min_events, const float* hot_mask, int zero_win0, int zero_n, unsigned long long grid_elems4) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(256);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  unsigned n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, voxel_atomic_kernel<Src, kAgg>, src, offsets, w0, num_bins, H, W, pad_top, pad_left, Hp, Wp, out,
                            stride, oob_count, min_events, hot_mask, zero_win0, zero_n, grid_elems4);
}

template <typename Src>
static int launch_atomic(const Src& src, bool agg, const int64_t* offsets, int T, int num_bins, int H, int W, int pad_top, int pad_left,
                         int Hp, int Wp, float* out, size_t out_window_stride, int* oob_count, int min_events, const float* hot_mask,
                         cudaStream_t s) {
  const size_t grid_elems = (size_t)num_bins * Hp * Wp;
  // the chunk being reduced and the chunk being zeroed share L2.  Measured (tools/voxel_probe.py --pipeline2): 32 MB is
  // best at 346x260 (17 windows per chunk), 40 MB at 1280x720 (two 18.4 MB windows per chunk instead of one)
  size_t chunk_mb = grid_elems * sizeof(float) > (8u << 20) ? 40 : 32;
  if (const char* e = getenv("BDE2VID_VOXEL_CHUNK_MB")) {
    const long v = atol(e);
    if (v > 0) chunk_mb = (size_t)v;
  }
  int per_chunk = (int)((chunk_mb << 20) / (grid_elems * sizeof(float)));   // L2 footprint of a window = its grid, not the stride
  per_chunk = per_chunk < 1 ? 1 : (per_chunk > T ? T : per_chunk);
  const int n_sm = device_sm_count();
  // in-kernel zeroing uses 16-byte stores
  bool zero_in_kernel = grid_elems % 4 == 0 && out_window_stride % 4 == 0 && (((uintptr_t)out) & 15) == 0;
  if (const char* e = getenv("BDE2VID_VOXEL_ZERO_IN_KERNEL")) zero_in_kernel = zero_in_kernel && e[0] != '0';
  // exactly one wave (4 CTAs of 256 threads x 64 registers per SM): the next launch of the chain can only start once every
  // CTA of this one has started, so a partial second wave delays it (5 or 6 CTAs per SM: 0.089 ms vs 0.076 ms per 100 windows)
  size_t ctas_per_sm = 4;
  if (const char* e = getenv("BDE2VID_VOXEL_CTAS_PER_SM")) {
    const long v = atol(e);
    if (v > 0) ctas_per_sm = (size_t)v;
  }
  for (int w0 = 0; w0 < T; w0 += per_chunk) {
    const int n = (T - w0 < per_chunk) ? T - w0 : per_chunk;
    if (w0 == 0 || !zero_in_kernel) {
      float* base = out + (size_t)w0 * out_window_stride;
      cudaError_t e = cudaMemset2DAsync(base, out_window_stride * sizeof(float), 0, grid_elems * sizeof(float), (size_t)n, s);
      BDE_REQUIRE(e == cudaSuccess, "bde_voxelize_seq: memset: %s", cudaGetErrorString(e));
    }
    const int zw0 = w0 + n;
    const int zn = zero_in_kernel ? ((T - zw0 < per_chunk) ? T - zw0 : per_chunk) : 0;
    // grid.x sized so that the chunk is a few waves over the SMs regardless of its window count
    int bx = (int)ceil_div((size_t)n_sm * ctas_per_sm, (size_t)n);
    bx = bx < 1 ? 1 : (bx > 1024 ? 1024 : bx);
    dim3 grid(bx, n);
    const cudaError_t e =
        agg ? launch_voxel_atomic<Src, true>(grid, s, src, offsets, w0, num_bins, H, W, pad_top, pad_left, Hp, Wp, out, out_window_stride,
                                             oob_count, min_events, hot_mask, zw0, zn, (unsigned long long)(grid_elems / 4))
            : launch_voxel_atomic<Src, false>(grid, s, src, offsets, w0, num_bins, H, W, pad_top, pad_left, Hp, Wp, out, out_window_stride,
                                              oob_count, min_events, hot_mask, zw0, zn, (unsigned long long)(grid_elems / 4));
    BDE_REQUIRE(e == cudaSuccess, "voxel_atomic_kernel: launch: %s", cudaGetErrorString(e));
    const int rc = check_launch("voxel_atomic_kernel");
    if (rc != 0) return rc;
  }
  return 0;
}

extern "C" int bde_voxelize_seq_strided(const float* xs, const float* ys, const float* ts, const float* ps,
                                        const int64_t* offsets, int T, int num_bins, int H, int W, int pad_top,
                                        int pad_left, int Hp, int Wp, float* out, size_t out_window_stride,
                                        int* oob_count, int algo, int min_events, const float* hot_mask, void* stream);

extern "C" int bde_voxelize_seq(const float* xs, const float* ys, const float* ts, const float* ps,
                                const int64_t* offsets, int T, int num_bins, int H, int W, int pad_top,
                                int pad_left, int Hp, int Wp, float* out, int* oob_count, int algo, int min_events,
                                void* stream) {
  return bde_voxelize_seq_strided(xs, ys, ts, ps, offsets, T, num_bins, H, W, pad_top, pad_left, Hp, Wp, out,
                                  (size_t)num_bins * Hp * Wp, oob_count, algo, min_events, nullptr, stream);
}

extern "C" int bde_voxelize_raw_strided(const int16_t* xs, const int16_t* ys, const double* ts, const uint8_t* ps,
                                        const int64_t* offsets, int T, int num_bins, int H, int W, int pad_top,
                                        int pad_left, int Hp, int Wp, float* out, size_t out_window_stride,
                                        int* oob_count, int algo, int min_events, const float* hot_mask, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const size_t grid_elems = (size_t)num_bins * Hp * Wp;
  BDE_REQUIRE(out_window_stride >= grid_elems, "bde_voxelize_raw: window stride smaller than one grid");
  BDE_REQUIRE(T >= 0 && num_bins >= 1 && H > 0 && W > 0, "bde_voxelize_raw: bad sizes");
  BDE_REQUIRE(pad_top >= 0 && pad_left >= 0 && Hp >= H + pad_top && Wp >= W + pad_left,
              "bde_voxelize_raw: padded grid %dx%d cannot hold %dx%d at (%d,%d)", Hp, Wp, H, W, pad_top, pad_left);
  BDE_REQUIRE(algo == 0 || algo == 2 || algo == 5, "bde_voxelize_raw: only the reduction kernels (algo 2 / 5) read the raw format");
  if (T == 0) return 0;
  if (algo == 0) {
    const char* e = getenv("BDE2VID_VOXEL_ALGO");
    algo = (e != nullptr && e[0] == '5') ? 5 : 2;
  }
  EvSrcRaw src;
  src.x = xs; src.y = ys; src.t = ts; src.p = ps; src.t0 = 0.0;
  return launch_atomic(src, algo == 5, offsets, T, num_bins, H, W, pad_top, pad_left, Hp, Wp, out, out_window_stride, oob_count,
                       min_events, hot_mask, s);
}

extern "C" int bde_voxelize_seq_strided(const float* xs, const float* ys, const float* ts, const float* ps,
                                        const int64_t* offsets, int T, int num_bins, int H, int W, int pad_top,
                                        int pad_left, int Hp, int Wp, float* out, size_t out_window_stride,
                                        int* oob_count, int algo, int min_events, const float* hot_mask, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const size_t grid_elems = (size_t)num_bins * Hp * Wp;
  BDE_REQUIRE(out_window_stride >= grid_elems, "bde_voxelize_seq: window stride smaller than one grid");
  BDE_REQUIRE(T >= 0 && num_bins >= 1 && H > 0 && W > 0, "bde_voxelize_seq: bad sizes");
  BDE_REQUIRE(pad_top >= 0 && pad_left >= 0 && Hp >= H + pad_top && Wp >= W + pad_left,
              "bde_voxelize_seq: padded grid %dx%d cannot hold %dx%d at (%d,%d)", Hp, Wp, H, W, pad_top, pad_left);
  if (T == 0) return 0;
  const size_t smem_budget = 200 * 1024;
  int band_rows = (int)(smem_budget / ((size_t)num_bins * W * sizeof(float)));
  band_rows = band_rows > Hp ? Hp : band_rows;
  int bands = band_rows > 0 ? (int)ceil_div(Hp, band_rows) : 1 << 30;
  // algorithm 3 (cluster / distributed shared memory): the grid of one window must fit the shared memory of at most 8 CTAs
  const size_t tile_budget = 227 * 1024 - 1024;
  int nc = 0, crow = 0;
  for (int c = 1; c <= 8; c *= 2) {
    const int rows = (int)ceil_div((size_t)H, (size_t)c);
    if ((size_t)num_bins * rows * W * sizeof(float) <= tile_budget) {
      nc = c;
      crow = rows;
      break;
    }
  }
  const bool cluster_ok = nc > 0 && Wp % 4 == 0 && (((uintptr_t)out) & 15) == 0 && out_window_stride % 4 == 0;
  if (algo == 0) {
    const char* e = getenv("BDE2VID_VOXEL_ALGO");
    if (e != nullptr && e[0] >= '1' && e[0] <= '5') algo = e[0] - '0';
  }
  // algorithm 4 (cluster: zero + global reductions in one launch) needs 16-byte stores into the grid
  const bool red_ok = Wp % 4 == 0 && (((uintptr_t)out) & 15) == 0 && out_window_stride % 4 == 0;
  // default: algorithm 2 (memset + global RED.ADD.F32).  Measured on B200, 346x260, 100 windows x 31,500 events:
  //   1 row-band tiles + warp aggregation 0.55 ms | 2 global atomics 0.14 ms | 3 cluster + DSMEM reductions 0.25 ms
  //   (scan-all + local atomics 0.33 ms) | 4 cluster zero + global reductions 0.20 ms
  // -- L2 executes float reductions in place faster than any shared-memory staging, and 8-CTA cluster launches are slow.
  if (algo == 0) algo = 2;
  if (hot_mask != nullptr && algo != 5) algo = 2;   // only the reduction kernels apply the hot-pixel mask
  if (algo == 4 && !red_ok) algo = 2;
  if (algo == 4) {
    const int vec_ok = ((((uintptr_t)xs) | ((uintptr_t)ys) | ((uintptr_t)ts) | ((uintptr_t)ps)) & 15) == 0;
    const int ncl = 8;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(T * ncl));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, voxel_cluster_red_kernel, xs, ys, ts, ps, offsets, num_bins, H, W, pad_top, pad_left, Hp, Wp,
                                       ncl, vec_ok, out, out_window_stride, oob_count, min_events);
    BDE_REQUIRE(e == cudaSuccess, "bde_voxelize_seq: cluster launch: %s", cudaGetErrorString(e));
    return check_launch("voxel_cluster_red_kernel");
  }
  // the cluster form needs a grid that fits 8 CTAs' shared memory, Wp % 4 == 0 and a 16-byte aligned output; otherwise
  // the request degrades to the row-band / atomic kernels
  if (algo == 3 && !cluster_ok) algo = (bands <= 16) ? 1 : 2;
  if (algo == 3) {
    const size_t smem = (size_t)num_bins * crow * W * sizeof(float);
    {  // set per call: the attribute is per device and the call is cheap
      cudaError_t e = cudaFuncSetAttribute(voxel_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      BDE_REQUIRE(e == cudaSuccess, "bde_voxelize_seq: smem attr: %s", cudaGetErrorString(e));
    }
    const int vec_ok = ((((uintptr_t)xs) | ((uintptr_t)ys) | ((uintptr_t)ts) | ((uintptr_t)ps)) & 15) == 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nc;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(T * nc));
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const char* sm = getenv("BDE2VID_VOXEL_SCANALL");
    const int scan_all = (sm != nullptr && sm[0] == '1') ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, voxel_cluster_kernel, xs, ys, ts, ps, offsets, num_bins, H, W, pad_top, pad_left, Hp, Wp,
                                       crow, nc, vec_ok, out, out_window_stride, oob_count, scan_all, min_events);
    BDE_REQUIRE(e == cudaSuccess, "bde_voxelize_seq: cluster launch: %s", cudaGetErrorString(e));
    return check_launch("voxel_cluster_kernel");
  }
  if (algo == 1) {
    BDE_REQUIRE(band_rows >= 1, "bde_voxelize_seq: sensor row too wide for the shared-memory algorithm");
    // balance the bands
    band_rows = (int)ceil_div(Hp, bands);
    size_t smem = (size_t)num_bins * band_rows * W * sizeof(float);
    auto kern = voxel_band_kernel<true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    BDE_REQUIRE(e == cudaSuccess, "bde_voxelize_seq: smem attr: %s", cudaGetErrorString(e));
    dim3 grid(bands, T);
    int vec_ok = ((((uintptr_t)xs) | ((uintptr_t)ys) | ((uintptr_t)ts) | ((uintptr_t)ps)) & 15) == 0;
    kern<<<grid, 512, smem, s>>>(xs, ys, ts, ps, offsets, num_bins, H, W, pad_top, pad_left, Hp, Wp,
                                 band_rows, vec_ok, out, out_window_stride, oob_count, min_events);
    return check_launch("voxel_band_kernel");
  }
  BDE_REQUIRE(algo == 2 || algo == 5, "bde_voxelize_seq: unknown algo %d", algo);
  EvSrcF32 src;
  src.x = xs; src.y = ys; src.t = ts; src.p = ps; src.t0 = 0.f;
  return launch_atomic(src, algo == 5, offsets, T, num_bins, H, W, pad_top, pad_left, Hp, Wp, out, out_window_stride, oob_count,
                       min_events, hot_mask, s);
}

extern "C" int bde_pack_voxel_nhwc(const float* vox, int T, int bins, int Hp, int Wp, int c_pad, void* out,
                                   int dtype, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  BDE_REQUIRE(c_pad >= bins, "bde_pack_voxel_nhwc: c_pad < bins");
  size_t plane = (size_t)Hp * Wp, total = plane * T;
  if (total == 0) return 0;
  unsigned blocks = (unsigned)ceil_div(total, 256);
  if (dtype == BDE_F32)
    pack_voxel_kernel<float><<<blocks, 256, 0, s>>>(vox, bins, plane, c_pad, (float*)out, total);
  else if (dtype == BDE_BF16)
    pack_voxel_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(vox, bins, plane, c_pad, (__nv_bfloat16*)out, total);
  else
    BDE_REQUIRE(false, "bde_pack_voxel_nhwc: bad dtype");
  return check_launch("pack_voxel_kernel");
}
