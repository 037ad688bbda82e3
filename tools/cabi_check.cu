// cabi_check.cu — hardware check of the C ABI without Python: a few seconds on one GPU.
//
// Exercises, through include/gcdlss_b200.h only, the entry points that were written without a GPU at hand and compares
// each with the established path or a host computation:
//   gcd_runtable_build + gcd_kmap_subm_runs   == gcd_hash_build + gcd_kmap_subm            (bit exact)
//   gcd_pairs_from_table                      == lists derived on the host                  (GCD_PAIRS_FUSED=1 selects the 2-pass build)
//   gcd_kmap_tile_sort                        == stable sort of the presence-mask keys      (bit exact; offsets per tile before / after)
//   gcd_conv_forward(out_rows = ...)          == scan-order gcd_conv_forward, and a host reference on sampled rows; timings of both
//   gcd_rows_gather                           == host gather                                (GCD_GATHER_FLAT=1 selects the flat kernel)
//   gcd_consistency_rows                      == host double-precision softmax / mse / max
// Prints one line per check and "ALL OK" / "FAILED n".  Run it twice: plain, and with
//   GCD_PAIRS_FUSED=1 GCD_GATHER_FLAT=1
// Build (tools/build_cabi_check.sh):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Iinclude -o tools/cabi_check tools/cabi_check.cu \
//        -L<pkg>/gcdlss_b200 -lgcdlss_sm100a -Xlinker -rpath -Xlinker '$ORIGIN/../<pkg>/gcdlss_b200'
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <unordered_set>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "gcdlss_b200.h"

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(2); } } while (0)
#define GCD(x) do { int32_t rc__ = (x); if (rc__ != 0) { printf("library error %d at %s:%d: %s\n", rc__, __FILE__, __LINE__, gcd_last_error_string()); exit(3); } } while (0)

static int g_failed = 0;
static void report(const char* what, bool ok, const char* detail = "") {
  printf("%-58s %s %s\n", what, ok ? "OK  " : "FAIL", detail);
  if (!ok) ++g_failed;
}
template <typename T> static T* dalloc(size_t n) { T* p; CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); return p; }
template <typename T> static std::vector<T> download(const T* d, size_t n) { std::vector<T> h(n); CK(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost)); return h; }
template <typename T> static T* upload(const std::vector<T>& h) { T* d = dalloc<T>(h.size()); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }

template <typename F> static float time_ms(F&& fn, int reps = 5) {
  fn();
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) fn();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

static int sort_bit(int k) {            // csrc/tilesort.cuh: offsets ranked by (number of non-zero components, k)
  auto cls = [](int j) { return ((j % 3) != 1) + (((j / 3) % 3) != 1) + ((j / 9) != 1); };
  int bit = 0;
  for (int j = 0; j < 27; ++j) bit += (cls(j) < cls(k) || (cls(j) == cls(k) && j < k)) ? 1 : 0;
  return bit;
}
static double offsets_per_tile(const std::vector<int32_t>& nbr, int kv, int64_t n) {
  int64_t tiles = (n + 127) / 128, active = 0;
  for (int64_t t = 0; t < tiles; ++t)
    for (int k = 0; k < kv; ++k) {
      bool hit = false;
      for (int64_t o = t * 128; o < std::min<int64_t>(n, t * 128 + 128) && !hit; ++o) hit = nbr[(size_t)k * n + o] >= 0;
      active += hit;
    }
  return (double)active / tiles;
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, sm_%d%d; GCD_PAIRS_FUSED=%s GCD_GATHER_FLAT=%s\n", prop.name, prop.major, prop.minor,
         getenv("GCD_PAIRS_FUSED") ? getenv("GCD_PAIRS_FUSED") : "-", getenv("GCD_GATHER_FLAT") ? getenv("GCD_GATHER_FLAT") : "-");
  std::mt19937 rng(1234);
  std::uniform_real_distribution<float> uni(0.f, 1.f);

  // ---- a surface-like voxel set: ground plane, two walls, clutter (negative coordinates included)
  std::vector<int32_t> coords;
  {
    std::unordered_set<uint64_t> seen;
    auto add = [&](int x, int y, int z) {
      const uint64_t key = ((uint64_t)(uint32_t)(x + 4096) << 40) | ((uint64_t)(uint32_t)(y + 4096) << 20) | (uint32_t)(z + 4096);
      if (seen.insert(key).second) { coords.push_back(0); coords.push_back(x); coords.push_back(y); coords.push_back(z); }
    };
    for (int x = -150; x < 150; ++x) for (int y = -150; y < 150; ++y) if (uni(rng) < 0.7f) add(x, y, (x * x + y * y) / 9000);
    for (int y = -150; y < 150; ++y) for (int z = 0; z < 40; ++z) if (uni(rng) < 0.8f) { add(60, y, z); add(-75 + z / 8, y, z); }
    for (int i = 0; i < 20000; ++i) add((int)(uni(rng) * 300) - 150, (int)(uni(rng) * 300) - 150, (int)(uni(rng) * 30));
    // shuffle rows: scan order of a LiDAR sweep is not raster order either
    const int64_t n0 = coords.size() / 4;
    std::vector<int64_t> perm(n0); std::iota(perm.begin(), perm.end(), 0); std::shuffle(perm.begin(), perm.end(), rng);
    std::vector<int32_t> c2(coords.size());
    for (int64_t i = 0; i < n0; ++i) memcpy(&c2[i * 4], &coords[perm[i] * 4], 16);
    coords.swap(c2);
  }
  const int64_t n = coords.size() / 4;
  const int kv = 27;
  printf("voxels: %lld\n", (long long)n);
  int32_t* d_coords = upload(coords);
  int32_t* d_status = dalloc<int32_t>(1); CK(cudaMemset(d_status, 0, 4));

  // ---- kernel maps: point-wise table vs run table
  const int64_t cap = gcd_hash_capacity(n);
  uint64_t* d_keys = dalloc<uint64_t>(cap); int32_t* d_vals = dalloc<int32_t>(cap);
  void* d_slots; CK(cudaMalloc(&d_slots, (size_t)cap * gcd_runtable_slot_bytes()));
  int32_t* d_nbr = dalloc<int32_t>((size_t)kv * n); int32_t* d_nbr_runs = dalloc<int32_t>((size_t)kv * n);
  GCD(gcd_hash_build(d_coords, n, d_keys, d_vals, cap, d_status, nullptr));
  GCD(gcd_kmap_subm(d_coords, n, d_keys, d_vals, cap, 3, 1, d_nbr, nullptr));
  GCD(gcd_runtable_build(d_coords, n, 1, d_slots, cap, d_status, nullptr));
  GCD(gcd_kmap_subm_runs(d_coords, n, d_slots, cap, 3, 1, d_nbr_runs, nullptr));
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> nbr = download(d_nbr, (size_t)kv * n), nbr_runs = download(d_nbr_runs, (size_t)kv * n);
  int32_t status = download(d_status, 1)[0];
  char buf[256];
  {
    const float t_pts = time_ms([&] { GCD(gcd_kmap_subm(d_coords, n, d_keys, d_vals, cap, 3, 1, d_nbr, nullptr)); });
    const float t_runs = time_ms([&] { GCD(gcd_kmap_subm_runs(d_coords, n, d_slots, cap, 3, 1, d_nbr_runs, nullptr)); });
    const float t_b0 = time_ms([&] { GCD(gcd_hash_build(d_coords, n, d_keys, d_vals, cap, d_status, nullptr)); });
    const float t_b1 = time_ms([&] { GCD(gcd_runtable_build(d_coords, n, 1, d_slots, cap, d_status, nullptr)); });
    snprintf(buf, sizeof buf, "status %d; K=3 search %.1f us vs %.1f us, build %.1f us vs %.1f us (points vs runs)", status, t_pts * 1e3, t_runs * 1e3, t_b0 * 1e3, t_b1 * 1e3);
    bool centre = true;
    for (int64_t o = 0; o < n; ++o) centre = centre && nbr[(size_t)13 * n + o] == (int32_t)o;
    report("run table kernel map == point table kernel map (K=3)", status == 0 && centre && nbr == nbr_runs, buf);
    int32_t* d_n5a = dalloc<int32_t>((size_t)125 * n); int32_t* d_n5b = dalloc<int32_t>((size_t)125 * n);
    GCD(gcd_kmap_subm(d_coords, n, d_keys, d_vals, cap, 5, 1, d_n5a, nullptr));
    GCD(gcd_kmap_subm_runs(d_coords, n, d_slots, cap, 5, 1, d_n5b, nullptr));
    const float t5a = time_ms([&] { GCD(gcd_kmap_subm(d_coords, n, d_keys, d_vals, cap, 5, 1, d_n5a, nullptr)); });
    const float t5b = time_ms([&] { GCD(gcd_kmap_subm_runs(d_coords, n, d_slots, cap, 5, 1, d_n5b, nullptr)); });
    snprintf(buf, sizeof buf, "K=5 search %.1f us vs %.1f us", t5a * 1e3, t5b * 1e3);
    report("run table kernel map == point table kernel map (K=5)", download(d_n5a, (size_t)125 * n) == download(d_n5b, (size_t)125 * n), buf);
    CK(cudaFree(d_n5a)); CK(cudaFree(d_n5b));
  }

  // ---- pair lists (kept for the weight-gradient check below)
  int32_t* d_pi = dalloc<int32_t>((size_t)kv * n); int32_t* d_po = dalloc<int32_t>((size_t)kv * n); int32_t* d_off = dalloc<int32_t>(kv + 1);
  {
    const size_t ws_bytes = gcd_pairs_workspace_bytes(n, kv);
    void* ws; CK(cudaMalloc(&ws, ws_bytes));
    GCD(gcd_pairs_from_table(d_nbr, n, kv, d_pi, d_po, d_off, ws, ws_bytes, nullptr));
    const float t = time_ms([&] { GCD(gcd_pairs_from_table(d_nbr, n, kv, d_pi, d_po, d_off, ws, ws_bytes, nullptr)); });
    std::vector<int32_t> off = download(d_off, kv + 1), ref_off(kv + 1), ref_in, ref_out;
    for (int k = 0; k < kv; ++k) {
      ref_off[k] = (int32_t)ref_in.size();
      for (int64_t o = 0; o < n; ++o) if (nbr[(size_t)k * n + o] >= 0) { ref_in.push_back(nbr[(size_t)k * n + o]); ref_out.push_back((int32_t)o); }
    }
    ref_off[kv] = (int32_t)ref_in.size();
    bool ok = off == ref_off;
    if (ok) ok = download(d_pi, ref_in.size()) == ref_in && download(d_po, ref_out.size()) == ref_out;
    snprintf(buf, sizeof buf, "%zu pairs, %.1f us", ref_in.size(), t * 1e3);
    report("pair lists == host", ok, buf);
    CK(cudaFree(ws));
  }

  // ---- tile sort
  int32_t* d_sorted = dalloc<int32_t>((size_t)kv * n); int32_t* d_rows = dalloc<int32_t>(n);
  uint32_t* d_masks = dalloc<uint32_t>((size_t)((n + 127) / 128));
  std::vector<int32_t> rows;
  {
    const size_t ws_bytes = gcd_tile_sort_workspace_bytes(n);
    void* ws; CK(cudaMalloc(&ws, ws_bytes));
    GCD(gcd_kmap_tile_sort(d_nbr, n, kv, d_sorted, d_rows, d_masks, ws, ws_bytes, nullptr));
    const float t = time_ms([&] { GCD(gcd_kmap_tile_sort(d_nbr, n, kv, d_sorted, d_rows, d_masks, ws, ws_bytes, nullptr)); });
    rows = download(d_rows, n);
    std::vector<int32_t> sorted = download(d_sorted, (size_t)kv * n);
    std::vector<uint64_t> key(n, 0);
    int bits[27];
    for (int k = 0; k < 27; ++k) bits[k] = sort_bit(k);
    for (int64_t o = 0; o < n; ++o) for (int k = 0; k < kv; ++k) key[o] |= (uint64_t)(nbr[(size_t)k * n + o] >= 0) << bits[k];
    std::vector<int32_t> ref_rows(n); std::iota(ref_rows.begin(), ref_rows.end(), 0);
    std::stable_sort(ref_rows.begin(), ref_rows.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
    bool ok = rows == ref_rows;
    for (int64_t i = 0; i < n && ok; ++i) for (int k = 0; k < kv; ++k) ok = ok && sorted[(size_t)k * n + i] == nbr[(size_t)k * n + rows[i]];
    snprintf(buf, sizeof buf, "%.1f us; offsets with a hit per 128-row tile %.1f -> %.1f", t * 1e3, offsets_per_tile(nbr, kv, n), offsets_per_tile(sorted, kv, n));
    report("tile sort == stable sort of the presence masks", ok, buf);
    CK(cudaFree(ws));
  }

  // ---- convolution: scan order vs tile-sorted table, bf16 tcgen05 path
  for (int cfg = 0; cfg < 2; ++cfg) {
    const int c_in = cfg == 0 ? 96 : 32, c_out = cfg == 0 ? 96 : 32;
    std::normal_distribution<float> gauss(0.f, 1.f);
    std::vector<__nv_bfloat16> x((size_t)n * c_in);
    std::vector<float> xf(x.size()), w((size_t)kv * c_in * c_out);
    for (size_t i = 0; i < x.size(); ++i) { x[i] = __float2bfloat16(gauss(rng)); xf[i] = __bfloat162float(x[i]); }
    for (auto& v : w) v = 0.05f * gauss(rng);
    __nv_bfloat16* d_x = upload(x); float* d_w = upload(w);
    void* d_packed; CK(cudaMalloc(&d_packed, gcd_conv_packed_weight_bytes(kv, c_in, c_out)));
    GCD(gcd_conv_pack_weights(d_w, kv, c_in, c_out, 0, 0, d_packed, nullptr));
    __nv_bfloat16* d_y0 = dalloc<__nv_bfloat16>((size_t)n * c_out); __nv_bfloat16* d_y1 = dalloc<__nv_bfloat16>((size_t)n * c_out);
    CK(cudaMemset(d_y1, 0xff, (size_t)n * c_out * 2));
    gcd_conv_args a; memset(&a, 0, sizeof a);
    a.in = d_x; a.ld_in = c_in; a.n_in = n; a.nbr = d_nbr; a.kv = kv; a.n_out = n; a.c_in = c_in; a.c_out = c_out;
    a.w = d_w; a.w_packed = d_packed; a.w_stride_k = (int64_t)c_in * c_out; a.w_stride_c = c_out; a.w_stride_n = 1;
    a.out = d_y0; a.ld_out = c_out; a.in_dtype = GCD_BF16; a.out_dtype = GCD_BF16; a.math_mode = GCD_MATH_BF16_TCGEN05;
    gcd_conv_args b = a; b.nbr = d_sorted; b.out_rows = d_rows; b.tile_masks = d_masks; b.out = d_y1;     // sorted table + per-tile offset masks
    GCD(gcd_conv_forward(&a, nullptr)); GCD(gcd_conv_forward(&b, nullptr));
    CK(cudaDeviceSynchronize());
    const float t0 = time_ms([&] { GCD(gcd_conv_forward(&a, nullptr)); });
    const float t1 = time_ms([&] { GCD(gcd_conv_forward(&b, nullptr)); });
    std::vector<__nv_bfloat16> y0 = download(d_y0, (size_t)n * c_out), y1 = download(d_y1, (size_t)n * c_out);
    double max_ref = 0, err0 = 0, err1 = 0, diff = 0, max_y = 0;
    for (size_t i = 0; i < y0.size(); ++i) {
      const double u = __bfloat162float(y0[i]), v = __bfloat162float(y1[i]);
      diff = std::max(diff, std::fabs(u - v)); max_y = std::max(max_y, std::fabs(u));
    }
    for (int s = 0; s < 300; ++s) {                            // host reference on sampled rows
      const int64_t o = (int64_t)(uni(rng) * n) % n;
      for (int j = 0; j < c_out; ++j) {
        double acc = 0;
        for (int k = 0; k < kv; ++k) {
          const int32_t i = nbr[(size_t)k * n + o];
          if (i < 0) continue;
          for (int c = 0; c < c_in; ++c) acc += (double)xf[(size_t)i * c_in + c] * (double)__bfloat162float(__float2bfloat16(w[((size_t)k * c_in + c) * c_out + j]));
        }
        max_ref = std::max(max_ref, std::fabs(acc));
        err0 = std::max(err0, std::fabs(acc - __bfloat162float(y0[(size_t)o * c_out + j])));
        err1 = std::max(err1, std::fabs(acc - __bfloat162float(y1[(size_t)o * c_out + j])));
      }
    }
    snprintf(buf, sizeof buf, "%d->%d: %.1f us scan order, %.1f us sorted; rel err vs host %.2e / %.2e, sorted vs scan %.2e", c_in, c_out, t0 * 1e3, t1 * 1e3,
             err0 / max_ref, err1 / max_ref, diff / max_y);
    report("conv through the tile-sorted table == scan-order conv", err0 / max_ref < 3e-2 && err1 / max_ref < 3e-2 && diff / max_y < 3e-2, buf);
    {
      // dynamic tile schedule (gcd_conv_args.sched): the same bits whichever CTA claims a tile, counters zero afterwards
      int32_t* d_sched = dalloc<int32_t>(2); CK(cudaMemset(d_sched, 0, 8));
      __nv_bfloat16* d_y2 = dalloc<__nv_bfloat16>((size_t)n * c_out);
      gcd_conv_args d = b; d.out = d_y2; d.sched = d_sched;
      GCD(gcd_conv_forward(&d, nullptr)); GCD(gcd_conv_forward(&d, nullptr));
      const float t2 = time_ms([&] { GCD(gcd_conv_forward(&d, nullptr)); });
      std::vector<__nv_bfloat16> y2 = download(d_y2, (size_t)n * c_out);
      std::vector<int32_t> sc = download(d_sched, 2);
      snprintf(buf, sizeof buf, "%d->%d: %.1f us dynamic vs %.1f us static", c_in, c_out, t2 * 1e3, t1 * 1e3);
      report("dynamic tile schedule == static striding (bit-exact), counters reset", memcmp(y2.data(), y1.data(), y1.size() * 2) == 0 && sc[0] == 0 && sc[1] == 0, buf);
      // weight gradient over the pair lists, static and dynamic schedule, against the host on sampled entries
      std::vector<__nv_bfloat16> gy((size_t)n * c_out);
      std::vector<float> gf(gy.size());
      for (size_t i = 0; i < gy.size(); ++i) { gy[i] = __float2bfloat16(gauss(rng)); gf[i] = __bfloat162float(gy[i]); }
      __nv_bfloat16* d_g = upload(gy);
      float* d_dw0 = dalloc<float>(w.size()); float* d_dw1 = dalloc<float>(w.size());
      CK(cudaMemset(d_dw0, 0, w.size() * 4)); CK(cudaMemset(d_dw1, 0, w.size() * 4));
      gcd_wgrad_args g; memset(&g, 0, sizeof g);
      g.in = d_x; g.ld_in = c_in; g.gout = d_g; g.ld_gout = c_out; g.pair_in = d_pi; g.pair_out = d_po; g.pair_off = d_off;
      g.n_pairs = (int64_t)kv * n; g.kv = kv; g.c_in = c_in; g.c_out = c_out; g.dw = d_dw0; g.n_out = n;
      g.in_dtype = GCD_BF16; g.gout_dtype = GCD_BF16; g.math_mode = GCD_MATH_BF16_TCGEN05;
      GCD(gcd_conv_wgrad(&g, nullptr));
      gcd_wgrad_args gd = g; gd.dw = d_dw1; gd.sched = d_sched;
      GCD(gcd_conv_wgrad(&gd, nullptr));
      std::vector<float> dw0 = download(d_dw0, w.size()), dw1 = download(d_dw1, w.size());
      sc = download(d_sched, 2);
      double werr = 0, wmax = 0, wdiff = 0;
      for (size_t i = 0; i < dw0.size(); ++i) wdiff = std::max(wdiff, (double)std::fabs(dw0[i] - dw1[i]));
      for (int s = 0; s < 40; ++s) {
        const int k = (int)(uni(rng) * kv) % kv, ci = (int)(uni(rng) * c_in) % c_in, co = (int)(uni(rng) * c_out) % c_out;
        double acc = 0;
        for (int64_t o = 0; o < n; ++o) {
          const int32_t i = nbr[(size_t)k * n + o];
          if (i >= 0) acc += (double)xf[(size_t)i * c_in + ci] * (double)gf[(size_t)o * c_out + co];
        }
        wmax = std::max(wmax, std::fabs(acc));
        werr = std::max(werr, std::fabs(acc - dw0[((size_t)k * c_in + ci) * c_out + co]));
      }
      snprintf(buf, sizeof buf, "%d->%d: rel err vs host %.2e, dynamic vs static %.2e (abs)", c_in, c_out, werr / wmax, wdiff);
      report("weight gradient (tcgen05, pair lists) == host, dynamic == static", werr / wmax < 1e-3 && wdiff < 1e-2 * wmax && sc[0] == 0 && sc[1] == 0, buf);
      CK(cudaFree(d_sched)); CK(cudaFree(d_y2)); CK(cudaFree(d_g)); CK(cudaFree(d_dw0)); CK(cudaFree(d_dw1));
    }
    CK(cudaFree(d_x)); CK(cudaFree(d_w)); CK(cudaFree(d_packed)); CK(cudaFree(d_y0)); CK(cudaFree(d_y1));
  }

  // ---- devoxelise gather
  {
    const int c = 96; const int64_t p = 3 * n + 17;
    std::vector<float> src((size_t)n * c); for (auto& v : src) v = uni(rng);
    std::vector<int64_t> idx(p); for (auto& v : idx) v = (int64_t)(uni(rng) * n) % n;
    float* d_src = upload(src); int64_t* d_idx = upload(idx); float* d_out = dalloc<float>((size_t)p * c);
    GCD(gcd_rows_gather(d_src, c, d_idx, p, c, d_out, c, nullptr));
    const float t = time_ms([&] { GCD(gcd_rows_gather(d_src, c, d_idx, p, c, d_out, c, nullptr)); });
    std::vector<float> out = download(d_out, (size_t)p * c);
    bool ok = true;
    for (int64_t i = 0; i < p && ok; ++i) ok = memcmp(&out[(size_t)i * c], &src[(size_t)idx[i] * c], c * 4) == 0;
    snprintf(buf, sizeof buf, "%.1f us, %.0f GB/s", t * 1e3, ((double)p * (8 + c * 4) + (double)n * c * 4) / (t * 1e-3) / 1e9);
    report("rows gather == host gather", ok, buf);
    CK(cudaFree(d_src)); CK(cudaFree(d_idx)); CK(cudaFree(d_out));
  }

  // ---- consistency terms
  {
    const int c = 17; const int64_t m = 50001;
    std::normal_distribution<float> gauss(0.f, 3.f);
    std::vector<float> ls((size_t)m * c), lt((size_t)m * c);
    for (auto& v : ls) v = gauss(rng);
    for (auto& v : lt) v = gauss(rng);
    float* d_ls = upload(ls); float* d_lt = upload(lt);
    float* d_sq = dalloc<float>(m); float* d_mp = dalloc<float>(m); int64_t* d_lab = dalloc<int64_t>(m); float* d_g = dalloc<float>((size_t)m * c);
    GCD(gcd_consistency_rows(d_ls, c, d_lt, c, m, c, 0.9f, d_sq, d_mp, d_lab, d_g, c, nullptr));
    std::vector<float> sq = download(d_sq, m), mp = download(d_mp, m), g = download(d_g, (size_t)m * c);
    std::vector<int64_t> lab = download(d_lab, m);
    double worst = 0; bool labels_ok = true;
    for (int64_t i = 0; i < m; ++i) {
      double ps[64], pt[64], zs = 0, zt = 0, ms = -1e30, mt = -1e30; int arg = 0;
      for (int j = 0; j < c; ++j) { ms = std::max(ms, (double)ls[i * c + j]); if (lt[i * c + j] > mt) { mt = lt[i * c + j]; arg = j; } }
      for (int j = 0; j < c; ++j) { ps[j] = std::exp(ls[i * c + j] - ms); zs += ps[j]; pt[j] = std::exp(lt[i * c + j] - mt); zt += pt[j]; }
      double err = 0, dot = 0;
      for (int j = 0; j < c; ++j) { ps[j] /= zs; pt[j] /= zt; err += (ps[j] - pt[j]) * (ps[j] - pt[j]); dot += ps[j] * (ps[j] - pt[j]); }
      worst = std::max(worst, std::fabs(err - sq[i]));
      worst = std::max(worst, std::fabs(1.0 / zt - mp[i]));
      for (int j = 0; j < c; ++j) worst = std::max(worst, std::fabs(2 * ps[j] * (ps[j] - pt[j] - dot) - g[i * c + j]));
      if (std::fabs(1.0 / zt - 0.9) > 1e-5) labels_ok = labels_ok && lab[i] == (1.0 / zt < 0.9 ? -1 : arg);
    }
    snprintf(buf, sizeof buf, "max abs error %.2e", worst);
    report("consistency terms == host double precision", worst < 2e-6 && labels_ok, buf);
  }

  printf(g_failed ? "FAILED %d\n" : "ALL OK\n", g_failed);
  return g_failed ? 1 : 0;
}
