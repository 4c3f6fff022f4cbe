// 3-DoF vehicle tracking models (one thread = one sample): bicycle-model step, analytic reference
// trajectories, ego-frame observation and their hand-derived adjoints.
//   pyth_veh3dofconti      env_ocp/env_model/pyth_veh3dofconti_model.py:24-203, resources/ref_traj_model.py
//   veh3dof_tracking       env_gen_ocp/env_model/veh3dof_tracking_model.py:11-102, robot/veh3dof_model.py:24-58
#pragma once
#include "rollout.cuh"

namespace gops {

// vehicle parameters, pyth_veh3dofconti.py:24-33 (doubles folded exactly as the python expressions)
struct VehC {
  float dt, m, Iz, c_lk, dkf, dm, dlfkf, dkk, dll;
};
__device__ __forceinline__ VehC veh_const() {
  constexpr double k_f = -128915.5, k_r = -85943.6, l_f = 1.06, l_r = 1.85, m = 1412.0, I_z = 1536.7, dt = 0.1;
  VehC c;
  c.dt = (float)dt;
  c.m = (float)m;
  c.Iz = (float)I_z;
  c.c_lk = (float)(dt * (l_f * k_f - l_r * k_r));    // delta_t * (l_f*k_f - l_r*k_r)
  c.dkf = (float)(dt * k_f);                         // delta_t * k_f
  c.dm = (float)(dt * m);                            // delta_t * m
  c.dlfkf = (float)(dt * l_f * k_f);                 // delta_t * l_f * k_f
  c.dkk = (float)(dt * (k_f + k_r));                 // delta_t * (k_f + k_r)
  c.dll = (float)(dt * (l_f * l_f * k_f + l_r * l_r * k_r));
  return c;
}

// s = [x, y, phi, u, v, w], a = [steer, a_x]  ->  next state (VehicleDynamicsModel.f_xu :24-61)
__device__ __forceinline__ void veh_step(const VehC& c, float* s, const float* a) {
  const float x = s[0], y = s[1], phi = s[2], u = s[3], v = s[4], w = s[5];
  float sn, cs;
  sincosf(phi, &sn, &cs);
  const float nx = x + c.dt * (u * cs - v * sn);
  const float ny = y + c.dt * (u * sn + v * cs);
  const float nphi = phi + c.dt * w;
  const float nu = u + c.dt * a[1];
  const float nv = (((c.m * v) * u + c.c_lk * w) - (c.dkf * a[0]) * u - (c.dm * (u * u)) * w) / (c.m * u - c.dkk);
  const float nw = (((c.Iz * w) * u + c.c_lk * v) - (c.dlfkf * a[0]) * u) / (c.Iz * u - c.dll);
  s[0] = nx; s[1] = ny; s[2] = angle_normalize(nphi); s[3] = nu; s[4] = nv; s[5] = nw;
}

// lam: adjoint of next state (in) -> adjoint of s (out); abar += dL/da
__device__ __forceinline__ void veh_step_bwd(const VehC& c, const float* s, const float* a, float* lam, float* abar) {
  const float phi = s[2], u = s[3], v = s[4], w = s[5];
  float sn, cs;
  sincosf(phi, &sn, &cs);
  const float Dv = c.m * u - c.dkk, Dw = c.Iz * u - c.dll;
  const float Nv = ((c.m * v) * u + c.c_lk * w) - (c.dkf * a[0]) * u - (c.dm * (u * u)) * w;
  const float Nw = ((c.Iz * w) * u + c.c_lk * v) - (c.dlfkf * a[0]) * u;
  const float iDv = 1.f / Dv, iDw = 1.f / Dw;
  const float lx = lam[0], ly = lam[1], lp = lam[2], lu = lam[3], lv = lam[4], lw = lam[5];
  const float dv_du = (c.m * v - c.dkf * a[0] - 2.f * c.dm * u * w) * iDv - Nv * c.m * iDv * iDv;
  const float dw_du = (c.Iz * w - c.dlfkf * a[0]) * iDw - Nw * c.Iz * iDw * iDw;
  lam[0] = lx;
  lam[1] = ly;
  lam[2] = lp + lx * c.dt * (-u * sn - v * cs) + ly * c.dt * (u * cs - v * sn);   // angle_normalize has unit slope
  lam[3] = lu + lx * c.dt * cs + ly * c.dt * sn + lv * dv_du + lw * dw_du;
  lam[4] = lx * (-c.dt * sn) + ly * (c.dt * cs) + lv * (c.m * u * iDv) + lw * (c.c_lk * iDw);
  lam[5] = lp * c.dt + lv * ((c.c_lk - c.dm * u * u) * iDv) + lw * (c.Iz * u * iDw);
  abar[0] += lv * (-c.dkf * u * iDv) + lw * (-c.dlfkf * u * iDw);
  abar[1] += lu * c.dt;
}

// ---- analytic reference trajectories (resources/ref_traj_model.py:27-232) ----------------------
// path: 0 sine, 1 double lane, 2 triangle, 3 circle; speed: 0 sine, 1 constant.  Evaluating only the
// selected path/speed equals the reference's masked sum (it adds exact zeros).
__device__ __forceinline__ float rt_int_u(const RtC& r, float t, int sp) {
  if (sp == 0) return (r.sp_c1 * cosf(r.sp_omega * t + r.sp_phi) + r.sp_b * t) + r.sp_c3;
  return r.sp_const * t;
}
__device__ __forceinline__ float rt_u(const RtC& r, float t, int sp) {
  return sp == 0 ? r.sp_A * sinf(r.sp_omega * t + r.sp_phi) + r.sp_b : r.sp_const;
}
__device__ __forceinline__ float rt_x(const RtC& r, float t, int path, int sp) {
  const float arc = rt_int_u(r, t, sp);
  return path == 3 ? r.circ_r * sinf(arc / r.circ_r) : arc;
}
__device__ __forceinline__ float torch_remainder(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.f && ((m < 0.f) != (b < 0.f))) m += b;
  return m;
}
__device__ __forceinline__ float rt_y(const RtC& r, float t, int path, int sp) {
  if (path == 0) return r.sine_A * sinf(r.sine_omega * t + r.sine_phi);
  if (path == 1) {
    if (t <= r.dl_t1) return r.dl_y1;
    if (t <= r.dl_t2) return r.dl_k1 * (t - r.dl_t1) + r.dl_y1;
    if (t <= r.dl_t3) return r.dl_y2;
    if (t <= r.dl_t4) return r.dl_k2 * (t - r.dl_t3) + r.dl_y2;
    return r.dl_y1;
  }
  if (path == 2) {
    const float s = torch_remainder(t, r.tri_T);
    if (s <= r.tri_half) return r.tri_k1 * s;
    if (s < r.tri_T) return r.tri_k2 * (s - r.tri_T);
    return 0.f;
  }
  return r.circ_r * (cosf(rt_int_u(r, t, sp) / r.circ_r) - 1.f);
}
__device__ __forceinline__ float rt_phi(const RtC& r, float t, int path, int sp) {
  const float t2 = t + 0.001f;
  const float dx = rt_x(r, t2, path, sp) - rt_x(r, t, path, sp);
  const float dy = rt_y(r, t2, path, sp) - rt_y(r, t, path, sp);
  return atan2f(dy, dx);
}

// One reference point -> ego-frame observation entries (ego_vehicle_coordinate_transform :189-203)
__device__ __forceinline__ void ego_obs(const float* s, float cs, float sn, float rx, float ry, float rphi, float ru,
                                        float* o4) {
  const float dx = rx - s[0], dy = ry - s[1];
  o4[0] = dx * cs - dy * sn;
  o4[1] = dx * sn + dy * cs;
  o4[2] = angle_normalize(rphi - s[2]);
  o4[3] = ru - s[3];
}

}  // namespace gops
