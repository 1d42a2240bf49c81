// extern "C" surface of libb200g16 (see include/b200g16.h for the contract and the
// gnark / gnark-crypto routines each entry point replaces).
#include "common.cuh"
#include "ec.cuh"

namespace b200 {
template <class F>
int msm_device(b200g16_ctx* ctx, const Affine<F>* d_bases, const Fr* d_scalars, size_t n, Affine<F>* out);

template <class F>
int fixed_base_mul_device(b200g16_ctx* ctx, const Affine<F>& base, const Fr* d_scalars, size_t n, Affine<F>* d_out);
int modmul_probe(b200g16_ctx* ctx, int blocks_per_sm, int chains, int iters, double* modmul_per_s, float* ms_out);

// out[i] = k_i * base; result either copied to the host (out_host) or kept as a new bases handle
template <class F>
static int fixed_base_entry(b200g16_ctx* ctx, const uint64_t* base, const uint64_t* scalars, size_t n, int group,
                            uint64_t* out_host, b200g16_bases** out_bases) {
  if (!ctx || !base || (n && !scalars) || (!out_host && !out_bases)) return fail(B200G16_ERR_ARG, "fixed_base: null");
  std::lock_guard<std::mutex> lock(ctx->mu);
  B200_CUDA(cudaSetDevice(ctx->device));
  Affine<F> b;
  memcpy(&b, base, sizeof(b));
  B200_TRY(ctx->io_a.ensure((n ? n : 1) * sizeof(Fr)));
  if (n) B200_CUDA(cudaMemcpyAsync(ctx->io_a.p, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  Affine<F>* d_out = nullptr;
  b200g16_bases* h = nullptr;
  if (out_bases) {
    h = new b200g16_bases();
    h->group = group;
    h->n = n;
    h->device = ctx->device;
    cudaError_t e = cudaMalloc(&h->d_points, (n ? n : 1) * sizeof(Affine<F>));
    if (e != cudaSuccess) { delete h; return fail(B200G16_ERR_CUDA, "fixed_base: %s", cudaGetErrorString(e)); }
    d_out = reinterpret_cast<Affine<F>*>(h->d_points);
  } else {
    B200_TRY(ctx->io_b.ensure((n ? n : 1) * sizeof(Affine<F>)));
    d_out = ctx->io_b.as<Affine<F>>();
  }
  int st = fixed_base_mul_device<F>(ctx, b, ctx->io_a.as<Fr>(), n, d_out);
  if (st == 0 && out_host && n) {
    cudaError_t e = cudaMemcpy(out_host, d_out, n * sizeof(Affine<F>), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) st = fail(B200G16_ERR_CUDA, "fixed_base: %s", cudaGetErrorString(e));
  }
  if (st != 0 && h) { cudaFree(h->d_points); delete h; h = nullptr; }
  if (out_bases) *out_bases = h;
  return st;
}

template <class F>
static int msm_entry(b200g16_ctx* ctx, const b200g16_bases* bases, int group, size_t offset, const void* scalars,
                     bool scalars_on_device, size_t n, uint64_t* out) {
  if (!ctx || !bases || !out || (n && !scalars)) return fail(B200G16_ERR_ARG, "msm: null argument");
  if (bases->group != group) return fail(B200G16_ERR_ARG, "msm: bases are G%d, call is G%d", bases->group, group);
  if (bases->device != ctx->device) return fail(B200G16_ERR_STATE, "msm: bases live on another device");
  if (offset > bases->n || n > bases->n - offset)
    return fail(B200G16_ERR_ARG, "msm: range [%zu,%zu) exceeds %zu bases", offset, offset + n, bases->n);
  std::lock_guard<std::mutex> lock(ctx->mu);
  B200_CUDA(cudaSetDevice(ctx->device));
  const Fr* d_scalars = reinterpret_cast<const Fr*>(scalars);
  if (!scalars_on_device && n) {
    B200_TRY(ctx->msm.scalars.ensure(n * sizeof(Fr)));
    B200_CUDA(cudaMemcpyAsync(ctx->msm.scalars.p, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    d_scalars = ctx->msm.scalars.as<Fr>();
  }
  Affine<F> res;
  B200_TRY(msm_device<F>(ctx, reinterpret_cast<const Affine<F>*>(bases->d_points) + offset, d_scalars, n, &res));
  memcpy(out, &res, sizeof(res));
  return 0;
}

template <class F>
static int upload(b200g16_ctx* ctx, const uint64_t* points, size_t n, int group, b200g16_bases** out) {
  if (!ctx || !out || (n && !points)) return fail(B200G16_ERR_ARG, "bases_upload: null argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  B200_CUDA(cudaSetDevice(ctx->device));
  b200g16_bases* b = new b200g16_bases();
  b->group = group;
  b->n = n;
  b->device = ctx->device;
  size_t bytes = (n ? n : 1) * sizeof(Affine<F>);
  cudaError_t e = cudaMalloc(&b->d_points, bytes);
  if (e == cudaSuccess && n) e = cudaMemcpy(b->d_points, points, n * sizeof(Affine<F>), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (b->d_points) cudaFree(b->d_points);
    delete b;
    return fail(B200G16_ERR_CUDA, "bases_upload: %s", cudaGetErrorString(e));
  }
  *out = b;
  return 0;
}

template <class F>
static void host_add(const uint64_t* a, const uint64_t* b, uint64_t* out) {
  Affine<F> pa, pb;
  memcpy(&pa, a, sizeof(pa));
  memcpy(&pb, b, sizeof(pb));
  XYZZ<F> acc = XYZZ<F>::from_affine(pa);
  acc.madd(pb);
  Affine<F> r = acc.to_affine();
  memcpy(out, &r, sizeof(r));
}

template <class F>
static void host_scalar_mul(const uint64_t* p, const uint64_t* k, uint64_t* out) {
  Affine<F> pa;
  Fr s;
  memcpy(&pa, p, sizeof(pa));
  memcpy(&s, k, sizeof(s));
  s = Fr::from_mont(s);
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = 255; i >= 0; i--) {
    acc.dbl();
    if ((s.l[i >> 5] >> (i & 31)) & 1) acc.madd(pa);
  }
  Affine<F> r = acc.to_affine();
  memcpy(out, &r, sizeof(r));
}
}  // namespace b200

using namespace b200;

extern "C" {

int b200g16_version(void) { return 100; }

const char* b200g16_last_error(void) { return last_error_buf(); }

int b200g16_init(int device, b200g16_ctx** out) {
  if (!out) return fail(B200G16_ERR_ARG, "init: out is null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(B200G16_ERR_NO_DEVICE, "init: no CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(B200G16_ERR_ARG, "init: device %d of %d", device, count);
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(B200G16_ERR_NO_DEVICE, "init: device %d is sm_%d%d; kernels are built for sm_100a only", device,
                prop.major, prop.minor);
  B200_CUDA(cudaSetDevice(device));
  b200g16_ctx* ctx = new b200g16_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  B200_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  for (auto& ev : ctx->ev) B200_CUDA(cudaEventCreate(&ev));
  *out = ctx;
  return 0;
}

void b200g16_destroy(b200g16_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  DevBuf* bufs[] = {&ctx->msm.scalars, &ctx->msm.digits, &ctx->msm.entries, &ctx->msm.counts, &ctx->msm.partials,
                    &ctx->msm.buckets, &ctx->msm.chunks,  &ctx->msm.windows, &ctx->msm.misc,   &ctx->msm.tasks,
                    &ctx->ntt.a,       &ctx->ntt.b,       &ctx->ntt.c,       &ctx->ntt.tw,     &ctx->io_a,
                    &ctx->io_b,        &ctx->io_c};
  for (DevBuf* b : bufs) b->release();
  if (ctx->msm.pinned) cudaFreeHost(ctx->msm.pinned);
  for (auto& ev : ctx->ev) cudaEventDestroy(ev);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

uint64_t b200g16_launch_count(const b200g16_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b200g16_last_timings(const b200g16_ctx* ctx, float* out_ms, int cap) {
  if (!ctx || !out_ms) return 0;
  int n = ctx->timings.n < cap ? ctx->timings.n : cap;
  for (int i = 0; i < n; i++) out_ms[i] = ctx->timings.ms[i];
  return n;
}

int b200g16_set_msm_window(b200g16_ctx* ctx, int c) {
  if (!ctx || c < 0 || c > 24 || c == 1) return fail(B200G16_ERR_ARG, "set_msm_window: bad argument");
  ctx->msm_window_override = c;
  return 0;
}

int b200g16_bases_upload_g1(b200g16_ctx* ctx, const uint64_t* points, size_t n, b200g16_bases** out) {
  return upload<Fp>(ctx, points, n, 1, out);
}
int b200g16_bases_upload_g2(b200g16_ctx* ctx, const uint64_t* points, size_t n, b200g16_bases** out) {
  return upload<Fp2>(ctx, points, n, 2, out);
}
void b200g16_bases_free(b200g16_bases* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->d_points) cudaFree(b->d_points);
  delete b;
}
size_t b200g16_bases_len(const b200g16_bases* b) { return b ? b->n : 0; }

int b200g16_bases_download(const b200g16_bases* b, size_t offset, size_t n, uint64_t* out) {
  if (!b || !out) return fail(B200G16_ERR_ARG, "bases_download: null");
  if (offset > b->n || n > b->n - offset) return fail(B200G16_ERR_ARG, "bases_download: range");
  size_t sz = b->group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
  B200_CUDA(cudaSetDevice(b->device));
  B200_CUDA(cudaMemcpy(out, (const char*)b->d_points + offset * sz, n * sz, cudaMemcpyDeviceToHost));
  return 0;
}

int b200g16_fixed_base_mul_g1(b200g16_ctx* ctx, const uint64_t base[8], const uint64_t* scalars, size_t n,
                              uint64_t* out_points) {
  return fixed_base_entry<Fp>(ctx, base, scalars, n, 1, out_points, nullptr);
}
int b200g16_fixed_base_mul_g2(b200g16_ctx* ctx, const uint64_t base[16], const uint64_t* scalars, size_t n,
                              uint64_t* out_points) {
  return fixed_base_entry<Fp2>(ctx, base, scalars, n, 2, out_points, nullptr);
}
int b200g16_bases_from_scalars_g1(b200g16_ctx* ctx, const uint64_t base[8], const uint64_t* scalars, size_t n,
                                  b200g16_bases** out) {
  return fixed_base_entry<Fp>(ctx, base, scalars, n, 1, nullptr, out);
}
int b200g16_bases_from_scalars_g2(b200g16_ctx* ctx, const uint64_t base[16], const uint64_t* scalars, size_t n,
                                  b200g16_bases** out) {
  return fixed_base_entry<Fp2>(ctx, base, scalars, n, 2, nullptr, out);
}

int b200g16_modmul_probe(b200g16_ctx* ctx, int blocks_per_sm, int chains, int iters, double* modmul_per_s,
                         float* ms) {
  if (!ctx || !modmul_per_s || !ms || blocks_per_sm < 1 || iters < 1) return fail(B200G16_ERR_ARG, "probe: bad arg");
  std::lock_guard<std::mutex> lock(ctx->mu);
  B200_CUDA(cudaSetDevice(ctx->device));
  return modmul_probe(ctx, blocks_per_sm, chains, iters, modmul_per_s, ms);
}

int b200g16_msm_g1(b200g16_ctx* ctx, const b200g16_bases* bases, size_t offset, const uint64_t* scalars, size_t n,
                   uint64_t out[8]) {
  return msm_entry<Fp>(ctx, bases, 1, offset, scalars, false, n, out);
}
int b200g16_msm_g2(b200g16_ctx* ctx, const b200g16_bases* bases, size_t offset, const uint64_t* scalars, size_t n,
                   uint64_t out[16]) {
  return msm_entry<Fp2>(ctx, bases, 2, offset, scalars, false, n, out);
}
int b200g16_msm_g1_dev(b200g16_ctx* ctx, const b200g16_bases* bases, size_t offset, const void* d_scalars, size_t n,
                       uint64_t out[8]) {
  return msm_entry<Fp>(ctx, bases, 1, offset, d_scalars, true, n, out);
}
int b200g16_msm_g2_dev(b200g16_ctx* ctx, const b200g16_bases* bases, size_t offset, const void* d_scalars, size_t n,
                       uint64_t out[16]) {
  return msm_entry<Fp2>(ctx, bases, 2, offset, d_scalars, true, n, out);
}

int b200g16_g1_add(const uint64_t a[8], const uint64_t b[8], uint64_t out[8]) {
  if (!a || !b || !out) return fail(B200G16_ERR_ARG, "g1_add: null");
  host_add<Fp>(a, b, out);
  return 0;
}
int b200g16_g2_add(const uint64_t a[16], const uint64_t b[16], uint64_t out[16]) {
  if (!a || !b || !out) return fail(B200G16_ERR_ARG, "g2_add: null");
  host_add<Fp2>(a, b, out);
  return 0;
}
int b200g16_g1_scalar_mul(const uint64_t p[8], const uint64_t k[4], uint64_t out[8]) {
  if (!p || !k || !out) return fail(B200G16_ERR_ARG, "g1_scalar_mul: null");
  host_scalar_mul<Fp>(p, k, out);
  return 0;
}
int b200g16_g2_scalar_mul(const uint64_t p[16], const uint64_t k[4], uint64_t out[16]) {
  if (!p || !k || !out) return fail(B200G16_ERR_ARG, "g2_scalar_mul: null");
  host_scalar_mul<Fp2>(p, k, out);
  return 0;
}

}  // extern "C"
