// displace.cu -- displacement.ApplyDisplacementMap (internal/displacement/displacement.go:145-280) on the device.
//
// The reference tessellates triangle by triangle on one host thread: split 1 -> 4 (tessellate, :36-103) until a
// triangle spans at most 4 texels of the displacement map in U and V and its displacement variation is below the
// adaptive threshold (isTessellatedEnough, :120-141), then moves every vertex along the triangle's own normal
// through its TBN matrix (applyDisplacement, :201-280).  Here every level of the subdivision is one launch over
// all triangles still being refined: thread = (parent, child), children that are fine enough are appended to the
// output in (parent, child) order by an exclusive scan, the others form the next level.  That is the order the
// reference's loops produce; with `per_triangle` the result is regrouped by input triangle (stable radix sort on
// the input index), which is the concatenation transport.go:633-646 builds by calling the function once per
// triangle.  Arithmetic is the reference's fp64 sequence without FMA: the output is bit-identical to the oracle.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

#include "dscene.cuh"
#include "shade_textures.cuh"

using namespace izpi;

namespace {

struct alignas(16) DTri {  // minimalTriangle (displacement.go:19-33) + bookkeeping, 128 bytes
  double v[9];    // vertex0 vertex1 vertex2
  double uv[6];   // u0 v0 u1 v1 u2 v2
  int32_t mat;
  int32_t base;   // index of the input triangle this one descends from
};
static_assert(sizeof(DTri) == 128, "DTri layout");

struct DispMap {
  const double* pixels;
  int32_t w, h;
  double mn, mx, max_du, max_dv;
};

// ImageTxt.Value(u, v).Z (texture/image.go:73-101): nearest texel, blue channel
__device__ __forceinline__ double disp_value(const DispMap& m, double u, double v) {
  long long i = go_int(u * (double)m.w);
  long long j = go_int((1 - v) * ((double)m.h - 0.001));
  if (i < 0) i = 0;
  if (j < 0) j = 0;
  if (i > m.w - 1) i = m.w - 1;
  if (j > m.h - 1) j = m.h - 1;
  return __ldg(m.pixels + ((size_t)j * m.w + (size_t)i) * 4 + 2);
}

__device__ __forceinline__ d3 mid(d3 a, d3 b) { return (a + b) / 2.0; }

// one child of tessellate(); thread = 4 * parent + child
__global__ void tessellate_kernel(const DTri* __restrict__ in, long long n_in, DTri* __restrict__ children, int32_t* __restrict__ enough,
                                  DispMap m) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= 4 * n_in) return;
  const DTri p = in[idx >> 2];
  const int c = (int)(idx & 3);
  d3 v0 = mk(p.v[0], p.v[1], p.v[2]), v1 = mk(p.v[3], p.v[4], p.v[5]), v2 = mk(p.v[6], p.v[7], p.v[8]);
  d3 a = mid(v0, v1), b = mid(v1, v2), cc = mid(v2, v0);
  double u0 = p.uv[0], w0 = p.uv[1], u1 = p.uv[2], w1 = p.uv[3], u2 = p.uv[4], w2 = p.uv[5];
  double ua = (u0 + u1) / 2.0, va = (w0 + w1) / 2.0, ub = (u1 + u2) / 2.0, vb = (w1 + w2) / 2.0, uc = (u2 + u0) / 2.0, vc = (w2 + w0) / 2.0;
  d3 q0, q1, q2;
  double t[6];
  if (c == 0) { q0 = v0; q1 = a; q2 = cc; t[0] = u0; t[1] = w0; t[2] = ua; t[3] = va; t[4] = uc; t[5] = vc; }
  else if (c == 1) { q0 = a; q1 = b; q2 = cc; t[0] = ua; t[1] = va; t[2] = ub; t[3] = vb; t[4] = uc; t[5] = vc; }
  else if (c == 2) { q0 = a; q1 = v1; q2 = b; t[0] = ua; t[1] = va; t[2] = u1; t[3] = w1; t[4] = ub; t[5] = vb; }
  else { q0 = cc; q1 = b; q2 = v2; t[0] = uc; t[1] = vc; t[2] = ub; t[3] = vb; t[4] = u2; t[5] = w2; }
  DTri o;
  o.v[0] = q0.x; o.v[1] = q0.y; o.v[2] = q0.z; o.v[3] = q1.x; o.v[4] = q1.y; o.v[5] = q1.z; o.v[6] = q2.x; o.v[7] = q2.y; o.v[8] = q2.z;
  for (int k = 0; k < 6; k++) o.uv[k] = t[k];
  o.mat = p.mat; o.base = p.base;
  children[idx] = o;
  // isTessellatedEnough (displacement.go:120-141)
  bool uv_ok = fabs(t[2] - t[0]) <= m.max_du && fabs(t[4] - t[2]) <= m.max_du && fabs(t[0] - t[4]) <= m.max_du &&
               fabs(t[3] - t[1]) <= m.max_dv && fabs(t[5] - t[3]) <= m.max_dv && fabs(t[1] - t[5]) <= m.max_dv;
  bool ok = false;
  if (uv_ok) {
    double d0 = disp_value(m, t[0], t[1]), d1 = disp_value(m, t[2], t[3]), d2 = disp_value(m, t[4], t[5]);
    double lo = d0 < (d1 < d2 ? d1 : d2) ? d0 : (d1 < d2 ? d1 : d2);
    double hi = d0 > (d1 > d2 ? d1 : d2) ? d0 : (d1 > d2 ? d1 : d2);
    ok = (hi - lo) * fabs(m.mx - m.mn) <= 2.0;  // adaptiveThreshold (displacement.go:183)
  }
  enough[idx] = ok ? 1 : 0;
}

__global__ void scatter_kernel(const DTri* __restrict__ children, const int32_t* __restrict__ enough, const int32_t* __restrict__ pos,
                               long long n, DTri* __restrict__ done, long long done_base, DTri* __restrict__ next_in) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  if (enough[idx]) done[done_base + pos[idx]] = children[idx];
  else next_in[idx - pos[idx]] = children[idx];
}

__global__ void keys_kernel(const DTri* __restrict__ done, long long n, int32_t* keys, int32_t* vals) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = done[i].base; vals[i] = (int32_t)i;
}

// applyDisplacement (displacement.go:201-280) for triangle order[i] (or i), written as 15 doubles + material
__global__ void displace_kernel(const DTri* __restrict__ done, const int32_t* __restrict__ order, long long n, DispMap m,
                                double* __restrict__ out, int32_t* __restrict__ out_mat) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const DTri t = done[order ? order[i] : i];
  d3 v0 = mk(t.v[0], t.v[1], t.v[2]), v1 = mk(t.v[3], t.v[4], t.v[5]), v2 = mk(t.v[6], t.v[7], t.v[8]);
  d3 e1 = v1 - v0, e2 = v2 - v0;
  d3 normal = unit(cross(e1, e2));
  double dU1 = t.uv[2] - t.uv[0], dU2 = t.uv[4] - t.uv[0], dV1 = t.uv[3] - t.uv[1], dV2 = t.uv[5] - t.uv[1];
  double f = 1.0 / (dU1 * dV2 - dU2 * dV1);
  d3 tg = unit(mk(f * (dV2 * e1.x - dV1 * e2.x), f * (dV2 * e1.y - dV1 * e2.y), f * (dV2 * e1.z - dV1 * e2.z)));
  d3 bt = unit(mk(f * (-dU2 * e1.x + dU1 * e2.x), f * (-dU2 * e1.y + dU1 * e2.y), f * (-dU2 * e1.z + dU1 * e2.z)));
  d3 vs[3] = {v0, v1, v2};
  double* o = out + 15 * i;
  for (int k = 0; k < 3; k++) {
    double z = m.mn + ((m.mx - m.mn) * disp_value(m, t.uv[2 * k], t.uv[2 * k + 1]));
    // mat3.MatrixVectorMul(tbn, (0, 0, z)) (mat3.go:34-40): all three products of every row are formed
    d3 d = mk(tg.x * 0.0 + bt.x * 0.0 + normal.x * z, tg.y * 0.0 + bt.y * 0.0 + normal.y * z, tg.z * 0.0 + bt.z * 0.0 + normal.z * z);
    d3 p = vs[k] + d;
    o[3 * k] = p.x; o[3 * k + 1] = p.y; o[3 * k + 2] = p.z;
  }
  for (int k = 0; k < 6; k++) o[9 + k] = t.uv[k];
  out_mat[i] = t.mat;
}

template <typename T>
int grow(T** p, size_t* cap, size_t need, size_t keep, cudaStream_t st) {
  if (need <= *cap) return IZPI_OK;
  size_t ncap = std::max(need, *cap * 2);
  T* q = nullptr;
  IZ_CUDA(cudaMalloc(&q, ncap * sizeof(T)));
  if (keep) IZ_CUDA(cudaMemcpyAsync(q, *p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
  IZ_CUDA(cudaStreamSynchronize(st));
  cudaFree(*p);
  *p = q; *cap = ncap;
  return IZPI_OK;
}

}  // namespace

struct DisplaceResult {
  double* d_out = nullptr;
  int32_t* d_mat = nullptr;
  int64_t n = 0;
};

void displace_result_free(izpi_ctx* ctx) {
  auto* r = static_cast<DisplaceResult*>(ctx->displace);
  if (!r) return;
  cudaFree(r->d_out); cudaFree(r->d_mat);
  delete r;
  ctx->displace = nullptr;
}

namespace {

// Working buffers of one izpi_displace call.  Owned here so that EVERY exit path -- in particular the IZ_CUDA early returns
// in the middle of a GB-scale tessellation -- releases them.
struct DisplaceWork {
  double* d_pix = nullptr;
  DTri *d_in = nullptr, *d_next = nullptr, *d_child = nullptr, *d_done = nullptr;
  int32_t *d_flag = nullptr, *d_pos = nullptr;
  void* d_tmp = nullptr;
  int32_t *d_k0 = nullptr, *d_k1 = nullptr, *d_v0 = nullptr, *d_v1 = nullptr;
  ~DisplaceWork() {
    cudaFree(d_pix); cudaFree(d_in); cudaFree(d_next); cudaFree(d_child); cudaFree(d_done); cudaFree(d_flag); cudaFree(d_pos); cudaFree(d_tmp);
    cudaFree(d_k0); cudaFree(d_k1); cudaFree(d_v0); cudaFree(d_v1);
  }
};

int displace_impl(izpi_ctx* ctx, int64_t n, const double* tris15, const int32_t* materials, int32_t tex_w, int32_t tex_h,
                  const double* pixels_rgba, double mn, double mx, int per_triangle, int64_t* n_out) {
  if (!ctx || n < 0 || (n > 0 && !tris15) || !pixels_rgba || tex_w < 2 || tex_h < 2 || !n_out || n > (1 << 28)) {
    set_error("izpi_displace: bad argument (the displacement map must be an image of at least 2x2 texels)");
    return IZPI_EINVAL;
  }
  IZ_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  displace_result_free(ctx);
  auto* res = new DisplaceResult();
  ctx->displace = res;
  *n_out = 0;
  if (n == 0) return IZPI_OK;
  // inputs
  std::vector<DTri> h_in((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    std::memcpy(h_in[i].v, tris15 + 15 * i, 15 * sizeof(double));
    h_in[i].mat = materials ? materials[i] : 0;
    h_in[i].base = (int32_t)i;
  }
  DisplaceWork w;
  double*& d_pix = w.d_pix;
  IZ_CUDA(cudaMalloc(&d_pix, (size_t)tex_w * tex_h * 32));
  IZ_CUDA(cudaMemcpyAsync(d_pix, pixels_rgba, (size_t)tex_w * tex_h * 32, cudaMemcpyHostToDevice, st));
  DispMap m{d_pix, tex_w, tex_h, mn, mx, 4.0 / (double)(tex_w - 1), 4.0 / (double)(tex_h - 1)};  // displacement.go:176-178
  DTri *&d_in = w.d_in, *&d_next = w.d_next, *&d_child = w.d_child, *&d_done = w.d_done;
  int32_t *&d_flag = w.d_flag, *&d_pos = w.d_pos;
  void*& d_tmp = w.d_tmp;
  size_t cap_in = 0, cap_next = 0, cap_child = 0, cap_done = 0, cap_flag = 0, cap_pos = 0, cap_tmp = 0;
  int rc;
  if ((rc = grow(&d_in, &cap_in, (size_t)n, 0, st)) != IZPI_OK) return rc;  // ~DisplaceWork releases everything
  IZ_CUDA(cudaMemcpyAsync(d_in, h_in.data(), (size_t)n * sizeof(DTri), cudaMemcpyHostToDevice, st));
  long long n_in = n, n_done = 0;
  for (int level = 0; n_in > 0; level++) {
    if (level > 40 || 4 * n_in > (1ll << 30)) { set_error("izpi_displace: subdivision does not terminate / too many triangles"); return IZPI_EINVAL; }
    long long nc = 4 * n_in;
    if ((rc = grow(&d_child, &cap_child, (size_t)nc, 0, st)) != IZPI_OK || (rc = grow(&d_flag, &cap_flag, (size_t)nc, 0, st)) != IZPI_OK ||
        (rc = grow(&d_pos, &cap_pos, (size_t)nc, 0, st)) != IZPI_OK) return rc;  // ~DisplaceWork releases everything
    tessellate_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(d_in, n_in, d_child, d_flag, m);
    ctx->launches++;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag, d_pos, (int)nc, st);
    if (tmp_bytes > cap_tmp) { cudaFree(d_tmp); d_tmp = nullptr; IZ_CUDA(cudaMalloc(&d_tmp, tmp_bytes)); cap_tmp = tmp_bytes; }
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flag, d_pos, (int)nc, st);
    ctx->launches++;
    int32_t last_pos = 0, last_flag = 0;
    IZ_CUDA(cudaMemcpyAsync(&last_pos, d_pos + nc - 1, 4, cudaMemcpyDeviceToHost, st));
    IZ_CUDA(cudaMemcpyAsync(&last_flag, d_flag + nc - 1, 4, cudaMemcpyDeviceToHost, st));
    IZ_CUDA(cudaStreamSynchronize(st));
    long long n_ok = (long long)last_pos + last_flag, n_todo = nc - n_ok;
    if ((rc = grow(&d_done, &cap_done, (size_t)(n_done + n_ok), (size_t)n_done, st)) != IZPI_OK ||
        (rc = grow(&d_next, &cap_next, (size_t)std::max<long long>(n_todo, 1), 0, st)) != IZPI_OK) return rc;  // ~DisplaceWork releases everything
    scatter_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(d_child, d_flag, d_pos, nc, d_done, n_done, d_next);
    ctx->launches++;
    IZ_CUDA(cudaGetLastError());
    n_done += n_ok;
    std::swap(d_in, d_next); std::swap(cap_in, cap_next);
    n_in = n_todo;
  }
  // order: as produced (levels, parents, children) or regrouped by input triangle (stable sort)
  int32_t* d_order = nullptr;
  int32_t *&d_k0 = w.d_k0, *&d_k1 = w.d_k1, *&d_v0 = w.d_v0, *&d_v1 = w.d_v1;
  if (per_triangle && n_done > 0 && n > 1) {
    IZ_CUDA(cudaMalloc(&d_k0, (size_t)n_done * 4)); IZ_CUDA(cudaMalloc(&d_k1, (size_t)n_done * 4));
    IZ_CUDA(cudaMalloc(&d_v0, (size_t)n_done * 4)); IZ_CUDA(cudaMalloc(&d_v1, (size_t)n_done * 4));
    keys_kernel<<<(unsigned)((n_done + 255) / 256), 256, 0, st>>>(d_done, n_done, d_k0, d_v0);
    ctx->launches++;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_k0, d_k1, d_v0, d_v1, (int)n_done, 0, 32, st);
    if (tmp_bytes > cap_tmp) { cudaFree(d_tmp); d_tmp = nullptr; IZ_CUDA(cudaMalloc(&d_tmp, tmp_bytes)); cap_tmp = tmp_bytes; }
    cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_k0, d_k1, d_v0, d_v1, (int)n_done, 0, 32, st);  // LSD radix sort: stable
    ctx->launches++;
    d_order = d_v1;
  }
  if (n_done > 0) {
    IZ_CUDA(cudaMalloc(&res->d_out, (size_t)n_done * 15 * sizeof(double)));
    IZ_CUDA(cudaMalloc(&res->d_mat, (size_t)n_done * sizeof(int32_t)));
    displace_kernel<<<(unsigned)((n_done + 127) / 128), 128, 0, st>>>(d_done, d_order, n_done, m, res->d_out, res->d_mat);
    ctx->launches++;
    IZ_CUDA(cudaGetLastError());
  }
  IZ_CUDA(cudaStreamSynchronize(st));
  res->n = n_done;
  *n_out = n_done;
  return IZPI_OK;
}

}  // namespace

extern "C" {

int izpi_displace(izpi_ctx* ctx, int64_t n, const double* tris15, const int32_t* materials, int32_t tex_w, int32_t tex_h,
                  const double* pixels_rgba, double mn, double mx, int per_triangle, int64_t* n_out) {
  int rc = displace_impl(ctx, n, tris15, materials, tex_w, tex_h, pixels_rgba, mn, mx, per_triangle, n_out);
  if (rc != IZPI_OK && ctx) {  // no half-built result: a later izpi_displace_fetch must fail, not return n = 0
    cudaSetDevice(ctx->device);
    displace_result_free(ctx);
  }
  return rc;
}

int izpi_displace_fetch(izpi_ctx* ctx, double* out_tris15, int32_t* out_materials) {
  if (!ctx) { set_error("izpi_displace_fetch: bad argument"); return IZPI_EINVAL; }
  auto* r = static_cast<DisplaceResult*>(ctx->displace);
  if (!r) { set_error("izpi_displace_fetch: izpi_displace has not been called"); return IZPI_ESTATE; }
  IZ_CUDA(cudaSetDevice(ctx->device));
  if (r->n > 0 && out_tris15) IZ_CUDA(cudaMemcpy(out_tris15, r->d_out, (size_t)r->n * 15 * sizeof(double), cudaMemcpyDeviceToHost));
  if (r->n > 0 && out_materials) IZ_CUDA(cudaMemcpy(out_materials, r->d_mat, (size_t)r->n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return IZPI_OK;
}

}  // extern "C"
