// block.cu — whole residual blocks / conv-bn-act triples as ONE C-ABI call each way.
//
// The MinkUNet step is ~470 launches; issued one by one from Python the host needs longer per step than the GPU.
// These two entry points sequence the existing launches (gcd_conv_forward, gcd_bn_*, gcd_conv_wgrad) of
//   conv3 -> bn -> relu [-> conv3 -> bn (+ shortcut: identity | 1x1 conv -> bn) -> add -> relu]
// on the caller's stream from C, so one block costs the host one call instead of ~6 (forward) / ~12 (backward).
// Mirrors MinkowskiEngine's BasicBlock (ref MinkowskiEngine/modules/resnet_block.py BasicBlock.forward, imported at
// ref models/minkunet.py:30) and the conv->bn->relu triples of ref models/minkunet.py:140-147, 169-199.
#include "common.cuh"

namespace gcd {
int32_t bn_forward_train(const void* x, int64_t ld_x, int64_t n, int32_t c, double* stats, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, float* mean, float* invstd, const void* res,
                         int64_t ld_res, int32_t relu, void* y, int64_t ld_y, int32_t dtype, void* stream);
int32_t bn_backward_train(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, const void* y, int64_t ld_y, int64_t n, int32_t c,
                          const float* mean, const float* invstd, const float* gamma, double* sums, int32_t relu, void* dx, int64_t ld_dx,
                          void* dres, int64_t ld_dres, float* dgamma, float* dbeta, int32_t dtype, void* stream, const float* beta_mask);
namespace {

template <typename T>
__global__ void add_inplace_kernel(T* __restrict__ dst, const T* __restrict__ src, int64_t n_vec) {
  // 16-byte vectors
  pdl_trigger(); pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec) return;
  uint4 a = reinterpret_cast<const uint4*>(dst)[i];
  const uint4 b = reinterpret_cast<const uint4*>(src)[i];
  if constexpr (sizeof(T) == 4) {
    float* fa = reinterpret_cast<float*>(&a);
    const float* fb = reinterpret_cast<const float*>(&b);
#pragma unroll
    for (int j = 0; j < 4; ++j) fa[j] += fb[j];
  } else {
    __nv_bfloat162* ha = reinterpret_cast<__nv_bfloat162*>(&a);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = __bfloat1622float2(ha[j]), y = __bfloat1622float2(hb[j]);
      ha[j] = __floats2bfloat162_rn(x.x + y.x, x.y + y.y);
    }
  }
  reinterpret_cast<uint4*>(dst)[i] = a;
}

int32_t add_inplace(void* dst, const void* src, int64_t n, int32_t c, int32_t dtype, cudaStream_t st) {
  const int64_t numel = n * (int64_t)c;
  if (numel == 0) return GCD_OK;
  const int per = dtype == GCD_F32 ? 4 : 8;
  GCD_REQUIRE(numel % per == 0, "gcd_block: channel count %d not a multiple of the vector width", c);
  const int64_t n_vec = numel / per;
  const unsigned g = (unsigned)ceil_div(n_vec, 256);
  if (dtype == GCD_F32) launch_pdl(add_inplace_kernel<float>, dim3(g), dim3(256), 0, st, (float*)dst, (const float*)src, n_vec);
  else launch_pdl(add_inplace_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, st, (__nv_bfloat16*)dst, (const __nv_bfloat16*)src, n_vec);
  GCD_LAUNCH_CHECK("gcd_block add");
  return GCD_OK;
}

// dst[i, 0:c] (=|+=) src[i, 0:c] in 16-byte vectors; `vpr` vectors per row.
template <typename T, bool kAdd>
__global__ void cols_kernel(T* __restrict__ dst, int64_t ld_dst, const T* __restrict__ src, int64_t ld_src, int64_t n, int vpr) {
  pdl_trigger(); pdl_wait();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * vpr) return;
  const int64_t row = t / vpr;
  const int v = (int)(t - row * vpr);
  constexpr int per = 16 / (int)sizeof(T);
  const uint4 b = *reinterpret_cast<const uint4*>(src + row * ld_src + (int64_t)v * per);
  uint4* d = reinterpret_cast<uint4*>(dst + row * ld_dst + (int64_t)v * per);
  if constexpr (!kAdd) { *d = b; return; }
  uint4 a = *d;
  if constexpr (sizeof(T) == 4) {
    float* fa = reinterpret_cast<float*>(&a);
    const float* fb = reinterpret_cast<const float*>(&b);
#pragma unroll
    for (int j = 0; j < 4; ++j) fa[j] += fb[j];
  } else {
    __nv_bfloat162* ha = reinterpret_cast<__nv_bfloat162*>(&a);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = __bfloat1622float2(ha[j]), y = __bfloat1622float2(hb[j]);
      ha[j] = __floats2bfloat162_rn(x.x + y.x, x.y + y.y);
    }
  }
  *d = a;
}

int32_t cols_op(const gcd_op* o, bool add, cudaStream_t st) {
  if (o->n == 0 || o->c == 0) return GCD_OK;
  const int esz = o->dtype == GCD_F32 ? 4 : 2, per = 16 / esz;
  GCD_REQUIRE(o->dst && o->src && o->n > 0 && o->c > 0, "gcd_run_ops: bad copy / add operands");
  GCD_REQUIRE(o->c % per == 0 && o->ld_dst % per == 0 && o->ld_src % per == 0 && o->ld_dst >= o->c && o->ld_src >= o->c &&
              (reinterpret_cast<uintptr_t>(o->dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(o->src) & 15) == 0,
              "gcd_run_ops: column copies need 16-byte aligned rows (c = %d)", o->c);
  const int vpr = o->c / per;
  const unsigned g = (unsigned)ceil_div(o->n * vpr, 256);
  if (o->dtype == GCD_F32) {
    if (add) launch_pdl(cols_kernel<float, true>, dim3(g), dim3(256), 0, st, (float*)o->dst, o->ld_dst, (const float*)o->src, o->ld_src, o->n, vpr);
    else launch_pdl(cols_kernel<float, false>, dim3(g), dim3(256), 0, st, (float*)o->dst, o->ld_dst, (const float*)o->src, o->ld_src, o->n, vpr);
  } else {
    if (add) launch_pdl(cols_kernel<__nv_bfloat16, true>, dim3(g), dim3(256), 0, st, (__nv_bfloat16*)o->dst, o->ld_dst, (const __nv_bfloat16*)o->src, o->ld_src, o->n, vpr);
    else launch_pdl(cols_kernel<__nv_bfloat16, false>, dim3(g), dim3(256), 0, st, (__nv_bfloat16*)o->dst, o->ld_dst, (const __nv_bfloat16*)o->src, o->ld_src, o->n, vpr);
  }
  GCD_LAUNCH_CHECK("gcd_run_ops(copy/add)");
  return GCD_OK;
}

#define GCD_TRY(expr)                 \
  do {                                \
    int32_t rc__ = (expr);            \
    if (rc__ != GCD_OK) return rc__;  \
  } while (0)

// Execution context of gcd_run_ops_exec (caller-owned, see the header): second stream + events + scheduler counters.
struct Exec {
  int device = -1;
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int32_t* counters = nullptr;      // device: [0..1] tile schedule of the caller's stream, [2..3] of the side stream; zero between launches
  bool side_used = false;           // something was issued on the side stream during the current call
};
inline int32_t* sched_main(const Exec* e) { return e && option(GCD_OPT_DYN_TILES) ? e->counters : nullptr; }
inline int32_t* sched_side(const Exec* e) { return e && option(GCD_OPT_DYN_TILES) ? e->counters + 2 : nullptr; }

int32_t unit_conv(const gcd_convbn* u, const void* in, int64_t ld_in, void* out, int32_t dtype, void* stream, const Exec* ex) {
  gcd_conv_args a{};
  a.in = in; a.ld_in = ld_in; a.n_in = u->n_in;
  a.nbr = u->nbr; a.kv = u->kv; a.n_out = u->n_out; a.c_in = u->c_in; a.c_out = u->c_out;
  a.w = u->w; a.w_packed = u->w_packed_fwd;
  a.w_stride_k = (int64_t)u->c_in * u->c_out; a.w_stride_c = u->c_out; a.w_stride_n = 1;
  a.mirror = 0; a.bias = nullptr; a.out = out; a.ld_out = u->c_out;
  a.in_dtype = dtype; a.out_dtype = dtype; a.stats = nullptr;
  a.math_mode = u->w_packed_fwd ? GCD_MATH_BF16_TCGEN05 : GCD_MATH_FP32_SIMT;
  a.out_rows = u->out_rows;
  a.tile_masks = u->tile_masks;
  a.sched = sched_main(ex);
  return gcd_conv_forward(&a, stream);
}

int32_t unit_dgrad(const gcd_convbn* u, const void* dy, void* dx, int32_t dtype, void* stream, const Exec* ex) {
  gcd_conv_args a{};
  a.in = dy; a.ld_in = u->c_out; a.n_in = u->n_out;
  a.nbr = u->back_nbr; a.kv = u->kv; a.n_out = u->n_in; a.c_in = u->c_out; a.c_out = u->c_in;
  a.w = u->w; a.w_packed = u->w_packed_bwd;
  a.w_stride_k = (int64_t)u->c_in * u->c_out; a.w_stride_c = 1; a.w_stride_n = u->c_out;
  a.mirror = u->back_mirror; a.bias = nullptr; a.out = dx; a.ld_out = u->c_in;
  a.in_dtype = dtype; a.out_dtype = dtype; a.stats = nullptr;
  a.math_mode = u->w_packed_bwd ? GCD_MATH_BF16_TCGEN05 : GCD_MATH_FP32_SIMT;
  a.out_rows = u->back_out_rows;
  a.tile_masks = u->back_tile_masks;
  a.sched = sched_main(ex);
  return gcd_conv_forward(&a, stream);
}

// `stream` is where the launch goes (the caller's stream, or the context's side stream); `sched` the matching counters
int32_t unit_wgrad(const gcd_convbn* u, const void* in, int64_t ld_in, const void* dy, int32_t dtype, void* stream, int32_t* sched) {
  gcd_wgrad_args a{};
  a.in = in; a.ld_in = ld_in; a.gout = dy; a.ld_gout = u->c_out;
  a.pair_in = u->pair_in; a.pair_out = u->pair_out; a.pair_off = u->pair_off;
  a.n_pairs = u->pair_in ? u->n_pairs : u->n_out;
  a.kv = u->kv; a.c_in = u->c_in; a.c_out = u->c_out;
  a.dw = u->dw; a.dbias = nullptr; a.n_out = u->n_out;
  a.in_dtype = dtype; a.gout_dtype = dtype;
  const bool tc = u->w_packed_fwd != nullptr && dtype == GCD_BF16 && u->c_out <= 256;
  a.math_mode = tc ? GCD_MATH_BF16_TCGEN05 : GCD_MATH_FP32_SIMT;
  a.sched = sched;
  return gcd_conv_wgrad(&a, stream);
}

// Weight gradient of a unit inside a backward block.  With a context and GCD_OPT_WGRAD_SIDE it goes to the side stream behind
// an event recorded on the caller's stream: mode 1 records it where the unit's output gradient is complete (call this BEFORE
// issuing the input-gradient kernel: the two then run side by side where the level leaves SMs idle), mode 2 where the
// input-gradient kernel has been issued (call it AFTER).  `phase` says which of the two call sites this is.
int32_t unit_wgrad_overlapped(const gcd_convbn* u, const void* in, int64_t ld_in, const void* dy, int32_t dtype, void* stream, Exec* ex,
                              int phase) {
  const int mode = ex ? option(GCD_OPT_WGRAD_SIDE) : 0;
  if (mode != 1 && mode != 2) return phase == 2 ? unit_wgrad(u, in, ld_in, dy, dtype, stream, sched_main(ex)) : GCD_OK;
  if (phase != mode) return GCD_OK;
  cudaError_t e = cudaEventRecord(ex->fork, as_stream(stream));
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ex->side, ex->fork, 0);
  if (e != cudaSuccess) return cuda_fail(e, "gcd_run_ops_exec(fork to the side stream)");
  ex->side_used = true;
  return unit_wgrad(u, in, ld_in, dy, dtype, ex->side, sched_side(ex));
}

int32_t unit_bn_fwd(const gcd_convbn* u, const void* x, const void* res, int32_t relu, void* y, int32_t dtype, void* stream) {
  return bn_forward_train(x, u->c_out, u->n_out, u->c_out, u->stats, u->gamma, u->beta, u->eps, u->momentum, u->running_mean,
                          u->running_var, u->mean, u->invstd, res, res ? u->c_out : 0, relu, y, u->c_out, dtype, stream);
}

int32_t unit_bn_bwd(const gcd_convbn* u, const void* dy, int64_t ld_dy, const void* x, const void* y, int32_t relu, void* dx, void* dres,
                    int32_t dtype, void* stream) {
  const int64_t ld = u->c_out;
  if (ld_dy == 0) ld_dy = ld;
  // conv -> BN -> ReLU without a residual (dres == NULL): the mask can be re-derived from x, see bn_backward_train
  return bn_backward_train(dy, ld_dy, x, ld, y, y ? ld : 0, u->n_out, u->c_out, u->mean, u->invstd, u->gamma, u->sums, relu, dx, ld,
                           dres, dres ? ld : 0, u->dgamma, u->dbeta, dtype, stream, (relu && !dres) ? u->beta : nullptr);
}

int32_t check_unit(const gcd_convbn* u, const char* who) {
  GCD_REQUIRE(u->w && u->gamma && u->beta && u->running_mean && u->running_var && u->mean && u->invstd, "%s: null parameter pointer", who);
  GCD_REQUIRE(u->kv >= 1 && u->c_in >= 1 && u->c_out >= 1, "%s: bad shape", who);
  GCD_REQUIRE(u->nbr || (u->kv == 1 && u->n_in == u->n_out), "%s: identity map needs kv == 1 and n_in == n_out", who);
  return GCD_OK;
}

int32_t block_forward(gcd_block_args* b, void* stream, Exec* ex);
int32_t block_backward(gcd_block_args* b, void* stream, Exec* ex);
}  // namespace
}  // namespace gcd

using namespace gcd;

extern "C" int32_t gcd_block_forward(gcd_block_args* b, void* stream) { return block_forward(b, stream, nullptr); }
extern "C" int32_t gcd_block_backward(gcd_block_args* b, void* stream) { return block_backward(b, stream, nullptr); }

namespace gcd {
namespace {
int32_t block_forward(gcd_block_args* b, void* stream, Exec* ex) {
  GCD_REQUIRE(b != nullptr, "gcd_block_forward: null args");
  GCD_REQUIRE(b->x && b->y1 && b->a1, "gcd_block_forward: null activation pointer");
  GCD_TRY(check_unit(&b->u1, "gcd_block_forward(u1)"));
  GCD_REQUIRE(b->u1.stats, "gcd_block_forward: null stats scratch");
  const int32_t dt = b->dtype;
  const int32_t unit_launches = option(GCD_OPT_BN_FUSED) ? 2 : 3;     // convolution + batch norm (one two-phase launch, or two)
  int32_t n = 0;
  b->launches = 0;
  if (b->u1.n_out == 0) return GCD_OK;
  GCD_TRY(unit_conv(&b->u1, b->x, b->ld_x, b->y1, dt, stream, ex));
  GCD_TRY(unit_bn_fwd(&b->u1, b->y1, nullptr, b->has_u2 ? 1 : b->relu1, b->a1, dt, stream));
  n += unit_launches;
  if (b->has_u2) {
    GCD_TRY(check_unit(&b->u2, "gcd_block_forward(u2)"));
    GCD_REQUIRE(b->y2 && b->out && b->u2.stats, "gcd_block_forward: null pointer for the second unit");
    GCD_REQUIRE(b->u2.n_in == b->u1.n_out && b->u2.n_out == b->u1.n_out && b->u2.c_in == b->u1.c_out, "gcd_block_forward: unit shapes do not chain");
    GCD_TRY(unit_conv(&b->u2, b->a1, b->u1.c_out, b->y2, dt, stream, ex));
    const void* res = b->x;
    if (b->has_ud) {
      GCD_TRY(check_unit(&b->ud, "gcd_block_forward(ud)"));
      GCD_REQUIRE(b->yd && b->rd && b->ud.stats, "gcd_block_forward: null pointer for the shortcut unit");
      GCD_TRY(unit_conv(&b->ud, b->x, b->ld_x, b->yd, dt, stream, ex));
      GCD_TRY(unit_bn_fwd(&b->ud, b->yd, nullptr, 0, b->rd, dt, stream));
      res = b->rd;
      n += unit_launches;
    } else {
      GCD_REQUIRE(b->u1.c_in == b->u2.c_out && b->ld_x == b->u1.c_in && b->u1.n_in == b->u1.n_out,
                  "gcd_block_forward: identity shortcut needs matching shapes and a dense input");
    }
    GCD_TRY(unit_bn_fwd(&b->u2, b->y2, res, 1, b->out, dt, stream));
    n += unit_launches;
  }
  b->launches = n;
  return GCD_OK;
}

int32_t block_backward(gcd_block_args* b, void* stream, Exec* ex) {
  GCD_REQUIRE(b != nullptr, "gcd_block_backward: null args");
  GCD_REQUIRE(b->gout && b->dy1 && b->u1.sums && b->u1.dw && b->u1.dgamma && b->u1.dbeta, "gcd_block_backward: null pointer");
  GCD_REQUIRE(!b->need_dx || b->dx, "gcd_block_backward: dx requested but NULL");
  const int32_t dt = b->dtype;
  const int32_t bn_launches = option(GCD_OPT_BN_FUSED) ? 1 : 2;
  int32_t n = 0;
  b->launches = 0;
  if (b->u1.n_out == 0) return GCD_OK;
  const void* g1 = b->gout;       // gradient arriving at the first unit's activation
  int64_t ld_g1 = b->ld_gout;
  if (b->has_u2) {
    GCD_REQUIRE(b->dy2 && b->dres && b->da1 && b->u2.sums && b->u2.dw && b->u2.dgamma && b->u2.dbeta, "gcd_block_backward: null pointer (u2)");
    GCD_TRY(unit_bn_bwd(&b->u2, b->gout, b->ld_gout, b->y2, b->out, 1, b->dy2, b->dres, dt, stream));
    GCD_TRY(unit_wgrad_overlapped(&b->u2, b->a1, b->u1.c_out, b->dy2, dt, stream, ex, 1));
    GCD_TRY(unit_dgrad(&b->u2, b->dy2, b->da1, dt, stream, ex));
    GCD_TRY(unit_wgrad_overlapped(&b->u2, b->a1, b->u1.c_out, b->dy2, dt, stream, ex, 2));
    g1 = b->da1;
    ld_g1 = 0;
    n += 2 + bn_launches;
  }
  const int32_t relu1 = b->has_u2 ? 1 : b->relu1;
  GCD_TRY(unit_bn_bwd(&b->u1, g1, ld_g1, b->y1, relu1 ? b->a1 : nullptr, relu1, b->dy1, nullptr, dt, stream));
  GCD_TRY(unit_wgrad_overlapped(&b->u1, b->x, b->ld_x, b->dy1, dt, stream, ex, 1));
  if (b->need_dx) { GCD_TRY(unit_dgrad(&b->u1, b->dy1, b->dx, dt, stream, ex)); ++n; }
  GCD_TRY(unit_wgrad_overlapped(&b->u1, b->x, b->ld_x, b->dy1, dt, stream, ex, 2));
  n += 1 + bn_launches;
  if (b->has_u2) {
    if (b->has_ud) {
      GCD_REQUIRE(b->dyd && b->ud.sums && b->ud.dw && b->ud.dgamma && b->ud.dbeta, "gcd_block_backward: null pointer (ud)");
      GCD_TRY(unit_bn_bwd(&b->ud, b->dres, 0, b->yd, nullptr, 0, b->dyd, nullptr, dt, stream));
      n += bn_launches;
      if (b->need_dx) {
        GCD_REQUIRE(b->dxd, "gcd_block_backward: null dxd");
        GCD_TRY(unit_wgrad_overlapped(&b->ud, b->x, b->ld_x, b->dyd, dt, stream, ex, 1));
        GCD_TRY(unit_dgrad(&b->ud, b->dyd, b->dxd, dt, stream, ex));
        GCD_TRY(add_inplace(b->dx, b->dxd, b->u1.n_in, b->u1.c_in, dt, as_stream(stream)));
        n += 2;
      } else {
        GCD_TRY(unit_wgrad_overlapped(&b->ud, b->x, b->ld_x, b->dyd, dt, stream, ex, 1));
      }
      GCD_TRY(unit_wgrad_overlapped(&b->ud, b->x, b->ld_x, b->dyd, dt, stream, ex, 2));
      ++n;
    } else if (b->need_dx) {
      GCD_TRY(add_inplace(b->dx, b->dres, b->u1.n_in, b->u1.c_in, dt, as_stream(stream)));
      ++n;
    }
  }
  b->launches = n;
  return GCD_OK;
}

int32_t run_ops(Exec* ex, gcd_op* ops, int32_t n_ops, void* stream, int32_t* launches) {
  GCD_REQUIRE(ops != nullptr && n_ops >= 0, "gcd_run_ops: bad arguments");
  int32_t total = 0;
  int32_t rc = GCD_OK;
  for (int32_t i = 0; i < n_ops && rc == GCD_OK; ++i) {
    gcd_op* o = &ops[i];
    switch (o->op) {
      case GCD_OP_BLOCK_FORWARD:
        if (o->block == nullptr) { set_error("gcd_run_ops: operation %d has no block", i); rc = GCD_ERR_INVALID_ARG; break; }
        rc = block_forward(o->block, stream, ex);
        total += o->block->launches;
        break;
      case GCD_OP_BLOCK_BACKWARD:
        if (o->block == nullptr) { set_error("gcd_run_ops: operation %d has no block", i); rc = GCD_ERR_INVALID_ARG; break; }
        rc = block_backward(o->block, stream, ex);
        total += o->block->launches;
        break;
      case GCD_OP_COPY_COLS:
        rc = cols_op(o, false, as_stream(stream));
        ++total;
        break;
      case GCD_OP_ADD_COLS:
        rc = cols_op(o, true, as_stream(stream));
        ++total;
        break;
      case GCD_OP_RECORD_EVENT: {
        // everything issued so far by this call -- on the caller's stream and on the context's side stream -- precedes the event
        cudaEvent_t ev = reinterpret_cast<cudaEvent_t>(o->dst);
        if (ev == nullptr) { set_error("gcd_run_ops: operation %d has no event", i); rc = GCD_ERR_INVALID_ARG; break; }
        cudaError_t e;
        if (ex && ex->side_used) {
          e = cudaEventRecord(ex->fork, as_stream(stream));
          if (e == cudaSuccess) e = cudaStreamWaitEvent(ex->side, ex->fork, 0);
          if (e == cudaSuccess) e = cudaEventRecord(ev, ex->side);
        } else {
          e = cudaEventRecord(ev, as_stream(stream));
        }
        if (e != cudaSuccess) rc = cuda_fail(e, "gcd_run_ops(record event)");
        break;
      }
      default:
        set_error("gcd_run_ops: unknown operation %d at index %d", o->op, i);
        rc = GCD_ERR_INVALID_ARG;
    }
  }
  // join: whatever the call put on the side stream is ordered before everything the caller issues next (also on errors)
  if (ex && ex->side_used) {
    ex->side_used = false;
    cudaError_t e = cudaEventRecord(ex->join, ex->side);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(as_stream(stream), ex->join, 0);
    if (e != cudaSuccess && rc == GCD_OK) rc = cuda_fail(e, "gcd_run_ops_exec(join)");
  }
  if (launches) *launches = total;
  return rc;
}
}  // namespace
}  // namespace gcd

extern "C" int32_t gcd_run_ops(gcd_op* ops, int32_t n_ops, void* stream, int32_t* launches) { return run_ops(nullptr, ops, n_ops, stream, launches); }

extern "C" int32_t gcd_run_ops_exec(void* exec, gcd_op* ops, int32_t n_ops, void* stream, int32_t* launches) {
  GCD_REQUIRE(exec != nullptr, "gcd_run_ops_exec: null context");
  Exec* ex = static_cast<Exec*>(exec);
  int dev = -1;
  if (const cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return cuda_fail(e, "gcd_run_ops_exec(cudaGetDevice)");
  GCD_REQUIRE(dev == ex->device, "gcd_run_ops_exec: context belongs to device %d, current device is %d", ex->device, dev);
  return run_ops(ex, ops, n_ops, stream, launches);
}

extern "C" int32_t gcd_exec_create(void** exec) {
  GCD_REQUIRE(exec != nullptr, "gcd_exec_create: null output pointer");
  *exec = nullptr;
  Exec* ex = new Exec();
  cudaError_t e = cudaGetDevice(&ex->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ex->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ex->fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ex->join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&ex->counters, 4 * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemset(ex->counters, 0, 4 * sizeof(int32_t));
  if (e != cudaSuccess) {
    gcd_exec_destroy(ex);
    return cuda_fail(e, "gcd_exec_create");
  }
  *exec = ex;
  return GCD_OK;
}

extern "C" int32_t gcd_exec_destroy(void* exec) {
  if (exec == nullptr) return GCD_OK;
  Exec* ex = static_cast<Exec*>(exec);
  if (ex->side) { cudaStreamSynchronize(ex->side); cudaStreamDestroy(ex->side); }
  if (ex->fork) cudaEventDestroy(ex->fork);
  if (ex->join) cudaEventDestroy(ex->join);
  if (ex->counters) cudaFree(ex->counters);
  delete ex;
  return GCD_OK;
}
