// b200mj_render.cu — rendering hand-off: a batched ray caster over the model's primitives (sm_100a).
//
// Stands where Physics.render / Camera.render stand in the reference (dm_control/mujoco/engine.py:178-233, 840-946),
// for the three images a task can observe: rgb, depth, segmentation. One thread per pixel, one CTA row per environment;
// the environment's objects (type, size, frame, colour) are staged once per CTA in shared memory and every thread
// walks them with ray / primitive intersections in the primitive's frame (plane, sphere, capsule, ellipsoid, cylinder,
// box; fp64: the images are compared pixel by pixel with the numpy restatement, oracle/render_oracle.py). The pixel
// <-> ray map inverts the reference's camera matrix (engine.py:759-810). It is NOT MuJoCo's OpenGL renderer: no
// textures, shadows, reflections or skybox — depth and segmentation are geometric and match, rgb is a headlight shade
// of geom_rgba.
//
// Memory: per environment nobj * (12 + 3 + 3) doubles are read once per CTA (coalesced), each pixel writes 3 + 4 + 8
// bytes; a 64 x 64 egocentric image of the 98-object CMU corridor scene is 4096 pixels x 98 tests.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200mj.h"
#include "../../include/b200mj_model_fields.h"

namespace {

struct Obj { double pos[3], mat[9], size[3], rb2; float rgb[3]; int type, kind, id; };      // rb2: squared bounding radius, < 0 for planes

__device__ __forceinline__ void pick(double t1, double t2, double tmin, double& t) {
  const double a = t1 > tmin ? t1 : INFINITY, b = t2 > tmin ? t2 : INFINITY;
  t = fmin(a, b);
}

// ray (o, d) in the primitive's frame -> nearest t beyond tmin (INFINITY: miss) and the outward normal there
__device__ double ray_primitive(int type, const double* s, const double* o, const double* d, double tmin, double* n) {
  double t = INFINITY;
  n[0] = n[1] = 0; n[2] = 1;
  if (type == BMJ_GEOM_PLANE) {
    if (!(d[2] < -1e-15)) return INFINITY;
    const double tt = -o[2] / d[2];
    if (!(tt > tmin)) return INFINITY;
    const double px = o[0] + tt * d[0], py = o[1] + tt * d[1];
    if (s[0] > 0 && fabs(px) > s[0]) return INFINITY;
    if (s[1] > 0 && fabs(py) > s[1]) return INFINITY;
    return tt;
  }
  if (type == BMJ_GEOM_SPHERE || type == BMJ_GEOM_ELLIPSOID) {
    double os[3], ds[3];
    const bool ell = type == BMJ_GEOM_ELLIPSOID;
    for (int k = 0; k < 3; k++) { const double sc = ell ? s[k] : 1.0; os[k] = o[k] / sc; ds[k] = d[k] / sc; }
    const double r = ell ? 1.0 : s[0];
    const double a = ds[0]*ds[0] + ds[1]*ds[1] + ds[2]*ds[2], b = os[0]*ds[0] + os[1]*ds[1] + os[2]*ds[2];
    const double c = os[0]*os[0] + os[1]*os[1] + os[2]*os[2] - r * r;
    const double disc = b * b - a * c;
    if (disc < 0) return INFINITY;
    const double sq = sqrt(disc);
    pick((-b - sq) / a, (-b + sq) / a, tmin, t);
    if (t == INFINITY) return t;
    for (int k = 0; k < 3; k++) { const double p = o[k] + t * d[k]; n[k] = ell ? p / (s[k] * s[k]) : p; }
    return t;
  }
  if (type == BMJ_GEOM_CAPSULE || type == BMJ_GEOM_CYLINDER) {
    const double r = s[0], h = s[1];
    const double a = d[0]*d[0] + d[1]*d[1], b = o[0]*d[0] + o[1]*d[1], c = o[0]*o[0] + o[1]*o[1] - r * r;
    const double disc = b * b - a * c;
    if (disc >= 0 && a > 1e-30) {
      const double sq = sqrt(disc);
      double t1 = (-b - sq) / a, t2 = (-b + sq) / a;
      if (!(t1 > tmin && fabs(o[2] + t1 * d[2]) <= h)) t1 = INFINITY;
      if (!(t2 > tmin && fabs(o[2] + t2 * d[2]) <= h)) t2 = INFINITY;
      t = fmin(t1, t2);
      if (t != INFINITY) { n[0] = o[0] + t * d[0]; n[1] = o[1] + t * d[1]; n[2] = 0; }
    }
    for (int e = 0; e < 2; e++) {
      const double sgn = e == 0 ? 1.0 : -1.0;
      if (type == BMJ_GEOM_CAPSULE) {
        const double oc[3] = {o[0], o[1], o[2] - sgn * h};
        const double aa = d[0]*d[0] + d[1]*d[1] + d[2]*d[2], bb = oc[0]*d[0] + oc[1]*d[1] + oc[2]*d[2];
        const double cc = oc[0]*oc[0] + oc[1]*oc[1] + oc[2]*oc[2] - r * r;
        const double dd = bb * bb - aa * cc;
        if (dd < 0) continue;
        const double sq = sqrt(dd);
        for (int root = 0; root < 2; root++) {
          const double tc = root == 0 ? (-bb - sq) / aa : (-bb + sq) / aa;
          const double zc = o[2] + tc * d[2];
          if (tc > tmin && sgn * zc >= h && tc < t) { t = tc; n[0] = oc[0] + tc * d[0]; n[1] = oc[1] + tc * d[1]; n[2] = oc[2] + tc * d[2]; }
        }
      } else {
        const double tc = (sgn * h - o[2]) / d[2];
        if (!isfinite(tc)) continue;
        const double px = o[0] + tc * d[0], py = o[1] + tc * d[1];
        if (tc > tmin && px * px + py * py <= r * r && tc < t) { t = tc; n[0] = 0; n[1] = 0; n[2] = sgn; }
      }
    }
    return t;
  }
  if (type == BMJ_GEOM_BOX) {
    double tn = -INFINITY, tf = INFINITY;
    for (int k = 0; k < 3; k++) {
      double lo, hi;
      if (fabs(d[k]) < 1e-300) {
        const bool inside = fabs(o[k]) <= s[k];
        lo = inside ? -INFINITY : INFINITY; hi = inside ? INFINITY : -INFINITY;
      } else {
        const double inv = 1.0 / d[k];
        const double ta = (-s[k] - o[k]) * inv, tb = (s[k] - o[k]) * inv;
        lo = fmin(ta, tb); hi = fmax(ta, tb);
      }
      tn = fmax(tn, lo); tf = fmin(tf, hi);
    }
    if (!(tn <= tf)) return INFINITY;
    const double tt = tn > tmin ? tn : tf;
    if (!(tt > tmin)) return INFINITY;
    int ax = 0; double best = -1;
    double p[3];
    for (int k = 0; k < 3; k++) { p[k] = o[k] + tt * d[k]; const double q = fabs(p[k]) / s[k]; if (q > best) { best = q; ax = k; } }
    n[0] = n[1] = n[2] = 0; n[ax] = p[ax] > 0 ? 1.0 : (p[ax] < 0 ? -1.0 : 0.0);
    return tt;
  }
  return INFINITY;
}

}  // namespace

extern "C" __global__ void __launch_bounds__(128)
b200mj_render_kernel(const __grid_constant__ b200mj_render_scene sc, int height, int width, uint8_t* rgb, float* depth, int32_t* seg) {
  extern __shared__ __align__(16) unsigned char raw[];
  Obj* objs = reinterpret_cast<Obj*>(raw);
  __shared__ int nvis;
  const int env = blockIdx.y;
  // stage the visible objects of this environment (in object order)
  if (threadIdx.x == 0) {
    int k = 0;
    for (int i = 0; i < sc.nobj; i++) if (!sc.visible || sc.visible[i]) {
      Obj& ob = objs[k++];
      ob.type = sc.obj_type[i]; ob.kind = sc.obj_kind[i]; ob.id = sc.obj_id[i];
      for (int c = 0; c < 3; c++) ob.rgb[c] = sc.rgba[4 * i + c];
      reinterpret_cast<int*>(&ob.pos[0])[0] = i;      // source index, replaced by the frame below
    }
    nvis = k;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nvis; k += blockDim.x) {
    Obj& ob = objs[k];
    const int i = reinterpret_cast<int*>(&ob.pos[0])[0];
    const double* p = sc.pos + ((size_t)env * sc.nobj + i) * 3;
    const double* m = sc.mat + ((size_t)env * sc.nobj + i) * 9;
    const double* s = sc.size + (size_t)env * sc.size_stride + (size_t)i * 3;
    double pp[3] = {p[0], p[1], p[2]};
    for (int c = 0; c < 9; c++) ob.mat[c] = m[c];
    for (int c = 0; c < 3; c++) { ob.size[c] = s[c]; ob.pos[c] = pp[c]; }
    // bounding sphere for the world-frame rejection test ahead of the frame change + primitive test (most rays miss most objects)
    double rb;
    switch (ob.type) {
      case BMJ_GEOM_SPHERE: rb = s[0]; break;
      case BMJ_GEOM_CAPSULE: rb = s[0] + s[1]; break;
      case BMJ_GEOM_CYLINDER: rb = sqrt(s[0] * s[0] + s[1] * s[1]); break;
      case BMJ_GEOM_ELLIPSOID: rb = fmax(s[0], fmax(s[1], s[2])); break;
      case BMJ_GEOM_BOX: rb = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]); break;
      default: rb = -1; break;
    }
    ob.rb2 = rb < 0 ? -1.0 : rb * rb * (1.0 + 1e-9);
  }
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= height * width) return;
  const int v = pix / width, u = pix - v * width;
  const double f = (height / 2.0) / tan(sc.fovy * (3.14159265358979323846 / 180.0) / 2);
  const double dc[3] = {(u - (width - 1) / 2.0) / f, -(v - (height - 1) / 2.0) / f, -1.0};
  const double* cm = sc.cam_xmat + (size_t)env * 9;
  const double* cp = sc.cam_xpos + (size_t)env * 3;
  double d[3], o[3] = {cp[0], cp[1], cp[2]};
  for (int r = 0; r < 3; r++) d[r] = cm[3 * r] * dc[0] + cm[3 * r + 1] * dc[1] + cm[3 * r + 2] * dc[2];
  double best = sc.zfar; int bid = -1, bkind = -1; float col[3] = {0, 0, 0};
  const double dd2 = d[0]*d[0] + d[1]*d[1] + d[2]*d[2], dnorm = sqrt(dd2);
  for (int k = 0; k < nvis; k++) {
    const Obj& ob = objs[k];
    double ol[3], dl[3], rel[3] = {o[0] - ob.pos[0], o[1] - ob.pos[1], o[2] - ob.pos[2]};
    if (ob.rb2 >= 0) {      // the ray's closest approach to the centre lies outside the bounding sphere, or the sphere is behind / beyond the best hit
      const double b = rel[0] * d[0] + rel[1] * d[1] + rel[2] * d[2], cc = rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2] - ob.rb2;
      const double disc = b * b - dd2 * cc;
      if (disc < 0 || (cc > 0 && b > 0)) continue;
      if (cc > 0 && (-b - sqrt(disc)) > best * dd2) continue;
    }
    for (int c = 0; c < 3; c++) {      // R^T v
      ol[c] = ob.mat[c] * rel[0] + ob.mat[3 + c] * rel[1] + ob.mat[6 + c] * rel[2];
      dl[c] = ob.mat[c] * d[0] + ob.mat[3 + c] * d[1] + ob.mat[6 + c] * d[2];
    }
    double nl[3];
    const double t = ray_primitive(ob.type, ob.size, ol, dl, sc.znear, nl);
    if (!(t < best)) continue;
    best = t; bid = ob.id; bkind = ob.kind;
    double n[3];
    for (int r = 0; r < 3; r++) n[r] = ob.mat[3 * r] * nl[0] + ob.mat[3 * r + 1] * nl[1] + ob.mat[3 * r + 2] * nl[2];
    const double nn = fmax(sqrt(n[0]*n[0] + n[1]*n[1] + n[2]*n[2]), 1e-300);
    const double cosv = -(n[0] * d[0] + n[1] * d[1] + n[2] * d[2]) / (nn * dnorm);
    const double shade = 0.4 + 0.6 * fmax(0.0, cosv);
    for (int c = 0; c < 3; c++) col[c] = (float)((double)ob.rgb[c] * shade);
  }
  const size_t at = (size_t)env * height * width + pix;
  if (rgb) for (int c = 0; c < 3; c++) { const double q = floor((double)col[c] * 255.0 + 0.5); rgb[at * 3 + c] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q)); }
  if (depth) depth[at] = (float)best;
  if (seg) { seg[at * 2] = bid; seg[at * 2 + 1] = bkind; }
}

extern "C" int b200mj_render(const b200mj_render_scene* scene, int batch, int height, int width, uint8_t* rgb, float* depth,
                             int32_t* seg, void* stream) {
  if (!scene || batch <= 0 || height <= 0 || width <= 0 || scene->nobj < 0) return -1;
  if (!scene->cam_xpos || !scene->cam_xmat || (scene->nobj > 0 && (!scene->obj_type || !scene->obj_kind || !scene->obj_id ||
      !scene->rgba || !scene->size || !scene->pos || !scene->mat))) return -1;
  const size_t smem = (size_t)(scene->nobj > 0 ? scene->nobj : 1) * sizeof(Obj);
  if (smem > 200 * 1024) return -3;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(b200mj_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  const dim3 grid((unsigned)((height * width + 127) / 128), (unsigned)batch);
  b200mj_render_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(*scene, height, width, rgb, depth, seg);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
