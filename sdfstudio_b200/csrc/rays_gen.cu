// The step before the hot path: camera ray generation, scene colliders, and the lattice
// generator of the meshing grid (row 2).  Pure elementwise fp32 work, one thread per ray / point; the expression trees follow
// the reference so that origins / directions / nears / fars agree to the last bits:
//   nerfstudio/cameras/cameras.py:459-695 (_generate_rays_from_coords: perspective + fisheye, no distortion parameters)
//   nerfstudio/model_components/scene_colliders.py:47-163 (AABBBoxCollider, NearFarCollider, SphereCollider)
//   nerfstudio/utils/marching_cubes.py:49-56 (np.linspace lattice, 'ij' order)
#include "common.cuh"

namespace sdfb200 {

struct RayGenArgs {
  const float *fx, *fy, *cx, *cy, *c2w;
  const int32_t* cam_type;
  const int32_t* cam_idx;
  const float* coords;
  int64_t n;
  int32_t n_cameras;
  float *origins, *directions, *pixel_area, *directions_norm;
};

__device__ __forceinline__ void cam_dir(int type, float u, float v, float (&d)[3]) {
  if (type == SDFB200_CAMERA_FISHEYE) {
    float theta = __fsqrt_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
    theta = fminf(fmaxf(theta, 0.0f), 3.14159265358979323846f);
    const float st = sinf(theta);
    d[0] = __fdiv_rn(__fmul_rn(u, st), theta);
    d[1] = __fdiv_rn(__fmul_rn(v, st), theta);
    d[2] = -cosf(theta);
  } else {
    d[0] = u; d[1] = v; d[2] = -1.0f;
  }
}

__global__ void __launch_bounds__(256) k_generate_rays(const RayGenArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int c = a.cam_idx[i];
  c = c < 0 ? 0 : (c >= a.n_cameras ? a.n_cameras - 1 : c);
  const float y = a.coords[i * 2], x = a.coords[i * 2 + 1];                // coords are (row, col) = (y, x)  cameras.py:549-550
  const float fx = a.fx[c], fy = a.fy[c], cx = a.cx[c], cy = a.cy[c];
  const int type = a.cam_type ? a.cam_type[c] : SDFB200_CAMERA_PERSPECTIVE;
  // image-plane coordinates of the pixel and of its +1 neighbours in x and y  (cameras.py:574-576)
  const float xc = __fsub_rn(x, cx), yc = __fsub_rn(y, cy);
  const float u0 = __fdiv_rn(xc, fx), v0 = -__fdiv_rn(yc, fy);
  const float u1 = __fdiv_rn(__fadd_rn(xc, 1.0f), fx), v2 = -__fdiv_rn(__fadd_rn(yc, 1.0f), fy);
  const float uu[3] = {u0, u1, u0}, vv[3] = {v0, v0, v2};
  const float* m = a.c2w + (int64_t)c * 12;                                // [3][4] row major
  float dn[3][3];
  float norm0 = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float d[3];
    cam_dir(type, uu[k], vv[k], d);
    float w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)                                            // sum(d[None, :] * R, -1)  cameras.py:662-664
      w[r] = __fadd_rn(__fadd_rn(__fmul_rn(d[0], m[r * 4 + 0]), __fmul_rn(d[1], m[r * 4 + 1])), __fmul_rn(d[2], m[r * 4 + 2]));
    // torch's CPU norm over a contiguous last dimension accumulates with fused multiply-adds (verified against the reference
    // goldens); its stand-alone vectorised torch.sqrt is not correctly rounded on AVX-512 hosts, so dx / dy / pixel_area and the
    // sphere collider agree with the reference to 1 ulp only (tests/test_gpu_raygen.py)
    const float nrm = __fsqrt_rn(__fmaf_rn(w[2], w[2], __fmaf_rn(w[1], w[1], __fmul_rn(w[0], w[0]))));
    if (k == 0) norm0 = nrm;
    const float den = fmaxf(nrm, 1e-12f);                                  // F.normalize
#pragma unroll
    for (int r = 0; r < 3; ++r) dn[k][r] = __fdiv_rn(w[r], den);
  }
  float dx = 0.f, dy = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float ex = __fsub_rn(dn[0][r], dn[1][r]), ey = __fsub_rn(dn[0][r], dn[2][r]);
    dx = __fadd_rn(dx, __fmul_rn(ex, ex)); dy = __fadd_rn(dy, __fmul_rn(ey, ey));
  }
  a.origins[i * 3] = m[3]; a.origins[i * 3 + 1] = m[7]; a.origins[i * 3 + 2] = m[11];
  a.directions[i * 3] = dn[0][0]; a.directions[i * 3 + 1] = dn[0][1]; a.directions[i * 3 + 2] = dn[0][2];
  if (a.pixel_area) a.pixel_area[i] = __fmul_rn(__fsqrt_rn(dx), __fsqrt_rn(dy));
  if (a.directions_norm) a.directions_norm[i] = norm0;
}

struct CollideArgs {
  const float *origins, *directions;
  int64_t n;
  int type;
  float p[6];          // aabb min xyz, max xyz | near, far | radius, soft, radius^2
  float near_plane;
  float *nears, *fars;
};

__global__ void __launch_bounds__(256) k_collide(const CollideArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  float nr, fr;
  if (a.type == SDFB200_COLLIDER_NEAR_FAR) {
    nr = a.p[0]; fr = a.p[1];
  } else {
    const float o[3] = {a.origins[i * 3], a.origins[i * 3 + 1], a.origins[i * 3 + 2]};
    const float d[3] = {a.directions[i * 3], a.directions[i * 3 + 1], a.directions[i * 3 + 2]};
    if (a.type == SDFB200_COLLIDER_AABB) {
      nr = -INFINITY; fr = INFINITY;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float inv = __fdiv_rn(1.0f, __fadd_rn(d[k], 1e-6f));          // scene_colliders.py:72
        const float t1 = __fmul_rn(__fsub_rn(a.p[k], o[k]), inv), t2 = __fmul_rn(__fsub_rn(a.p[3 + k], o[k]), inv);
        nr = fmaxf(nr, fminf(t1, t2)); fr = fminf(fr, fmaxf(t1, t2));
      }
      nr = fmaxf(nr, a.near_plane);
      fr = fmaxf(fr, __fadd_rn(nr, 1e-6f));
    } else {                                                              // sphere, scene_colliders.py:143-163
      const float rc = __fadd_rn(__fadd_rn(__fmul_rn(d[0], o[0]), __fmul_rn(d[1], o[1])), __fmul_rn(d[2], o[2]));
      const float on = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(o[0], o[0]), __fmul_rn(o[1], o[1])), __fmul_rn(o[2], o[2])));
      float us = __fsub_rn(__fmul_rn(rc, rc), __fsub_rn(__fmul_rn(on, on), a.p[2]));         // p[2] = radius**2 evaluated in double on the host
      us = fmaxf(us, 0.01f);
      if (a.p[1] != 0.f) us = a.p[0];                                      // soft_intersection: ones * radius
      const float sq = __fsqrt_rn(us);
      nr = fmaxf(__fsub_rn(-sq, rc), 0.01f);
      fr = fmaxf(__fsub_rn(sq, rc), 0.01f);
    }
  }
  a.nears[i] = nr; a.fars[i] = fr;
}

// lattice point `idx` of np.meshgrid(linspace(min, max, res), indexing="ij"): x slowest, z fastest
__global__ void __launch_bounds__(256) k_lattice_points(double x0, double y0, double z0, double sx, double sy, double sz, double x1, double y1, double z1,
                                                        int rx, int ry, int rz, int64_t start, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t idx = start + i;
  const int iz = (int)(idx % rz);
  const int64_t t = idx / rz;
  const int iy = (int)(t % ry), ix = (int)(t / ry);
  // np.linspace: start + i * step in float64, last sample pinned to `stop`
  out[i * 3 + 0] = (float)(ix == rx - 1 && rx > 1 ? x1 : __dadd_rn(__dmul_rn((double)ix, sx), x0));   // no fma contraction
  out[i * 3 + 1] = (float)(iy == ry - 1 && ry > 1 ? y1 : __dadd_rn(__dmul_rn((double)iy, sy), y0));
  out[i * 3 + 2] = (float)(iz == rz - 1 && rz > 1 ? z1 : __dadd_rn(__dmul_rn((double)iz, sz), z0));
}

}  // namespace sdfb200

using namespace sdfb200;

extern "C" int sdfb200_generate_rays(const float* fx, const float* fy, const float* cx, const float* cy, const int32_t* camera_type,
                                     const float* camera_to_worlds, int32_t n_cameras, const int32_t* camera_indices, const float* coords,
                                     int64_t n_rays, float* origins, float* directions, float* pixel_area, float* directions_norm, void* stream) {
  SDFB_REQUIRE(n_rays >= 0 && n_cameras >= 1, "bad sizes");
  if (n_rays == 0) return 0;
  SDFB_REQUIRE(fx && fy && cx && cy && camera_to_worlds && camera_indices && coords && origins && directions, "NULL pointer");
  RayGenArgs a;
  a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.c2w = camera_to_worlds; a.cam_type = camera_type; a.cam_idx = camera_indices; a.coords = coords;
  a.n = n_rays; a.n_cameras = n_cameras; a.origins = origins; a.directions = directions; a.pixel_area = pixel_area; a.directions_norm = directions_norm;
  k_generate_rays<<<(unsigned)ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(a);
  SDFB_LAUNCHED("k_generate_rays");
  return 0;
}

extern "C" int sdfb200_collide(const float* origins, const float* directions, int64_t n_rays, int32_t collider_type, const float* params,
                               float near_plane, float* nears, float* fars, void* stream) {
  SDFB_REQUIRE(n_rays >= 0, "bad sizes");
  if (n_rays == 0) return 0;
  SDFB_REQUIRE(params && nears && fars, "NULL pointer");
  SDFB_REQUIRE(collider_type == SDFB200_COLLIDER_AABB || collider_type == SDFB200_COLLIDER_NEAR_FAR || collider_type == SDFB200_COLLIDER_SPHERE, "collider_type");
  if (collider_type != SDFB200_COLLIDER_NEAR_FAR) SDFB_REQUIRE(origins && directions, "NULL pointer");
  CollideArgs a;
  a.origins = origins; a.directions = directions; a.n = n_rays; a.type = collider_type; a.near_plane = near_plane; a.nears = nears; a.fars = fars;
  const int np = collider_type == SDFB200_COLLIDER_AABB ? 6 : (collider_type == SDFB200_COLLIDER_SPHERE ? 3 : 2);
  for (int k = 0; k < 6; ++k) a.p[k] = k < np ? params[k] : 0.f;        // `params` is a HOST array
  k_collide<<<(unsigned)ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(a);
  SDFB_LAUNCHED("k_collide");
  return 0;
}

extern "C" int sdfb200_lattice_points(const double* bbox_min, const double* bbox_max, const int32_t* resolution, int64_t start, int64_t n,
                                      float* points, void* stream) {
  SDFB_REQUIRE(bbox_min && bbox_max && resolution, "NULL pointer");
  SDFB_REQUIRE(resolution[0] >= 1 && resolution[1] >= 1 && resolution[2] >= 1, "resolution");
  const int64_t total = (int64_t)resolution[0] * resolution[1] * resolution[2];
  SDFB_REQUIRE(start >= 0 && n >= 0 && start + n <= total, "range outside the lattice");
  if (n == 0) return 0;
  SDFB_REQUIRE(points != nullptr, "NULL pointer");
  double st[3];
  for (int k = 0; k < 3; ++k) st[k] = resolution[k] > 1 ? (bbox_max[k] - bbox_min[k]) / (double)(resolution[k] - 1) : 0.0;
  k_lattice_points<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(bbox_min[0], bbox_min[1], bbox_min[2], st[0], st[1], st[2], bbox_max[0],
                                                                                  bbox_max[1], bbox_max[2], resolution[0], resolution[1],
                                                                                  resolution[2], start, n, points);
  SDFB_LAUNCHED("k_lattice_points");
  return 0;
}
