"""Agent::step bodies written as CUDA C snippets for kg_field2d_step_custom (shared by the CPU and GPU tests)."""

# Bird::step (tests/model/flockers/bird.rs:39-155) exactly as the built-in generic kernel evaluates it
# (csrc/boids_device.cuh boids_pair / boids_finish): acc = (x_avoid, y_avoid, x_cohe, y_cohe, x_cons, y_cons),
# c = (cohesion, avoidance, randomness, consistency, momentum, jump)
BIRD_PAIR = """
if (sid != oid) {                                   // bird.rs:63
  cnt += 1;
  float sq = fadd(fmul(dx, dx), fmul(dy, dy));
  float den = fadd(fmul(sq, sq), 1.0f);
  acc[0] = fadd(acc[0], fdiv(dx, den));             // :70-71
  acc[1] = fadd(acc[1], fdiv(dy, den));
  acc[2] = fadd(acc[2], dx);                        // :74-75
  acc[3] = fadd(acc[3], dy);
  acc[4] = fadd(acc[4], oa);                        // :78-79
  acc[5] = fadd(acc[5], ob);
}
"""
BIRD_FINISH = """
float avx = 0.f, avy = 0.f, cox = 0.f, coy = 0.f, rax = 0.f, ray = 0.f, csx = 0.f, csy = 0.f;
if (nvec != 0) {                                    // bird.rs:52
  float xa = acc[0], ya = acc[1], xc = acc[2], yc = acc[3], xs = acc[4], ys = acc[5];
  if (cnt > 0) {
    float cf = (float)cnt;
    xa = fdiv(xa, cf); ya = fdiv(ya, cf); xc = fdiv(xc, cf); yc = fdiv(yc, cf); xs = fdiv(xs, cf); ys = fdiv(ys, cf);
    csx = fdiv(xs, cf); csy = fdiv(ys, cf);         // divided by count twice, :88-91
  } else { csx = xs; csy = ys; }
  avx = fmul(400.0f, xa); avy = fmul(400.0f, ya);
  cox = fdiv(-xc, 10.0f); coy = fdiv(-yc, 10.0f);
  float xr = fsub(fmul(u0, 2.0f), 1.0f), yr = fsub(fmul(u1, 2.0f), 1.0f);
  float sq = fsqrt(fadd(fmul(xr, xr), fmul(yr, yr)));
  rax = fdiv(fmul(0.05f, xr), sq); ray = fdiv(fmul(0.05f, yr), sq);
}
float ddx = fadd(fadd(fadd(fadd(fmul(c[0], cox), fmul(c[1], avx)), fmul(c[3], csx)), fmul(c[2], rax)), fmul(c[4], sa));
float ddy = fadd(fadd(fadd(fadd(fmul(c[0], coy), fmul(c[1], avy)), fmul(c[3], csy)), fmul(c[2], ray)), fmul(c[4], sb));
float dis = fsqrt(fadd(fmul(ddx, ddx), fmul(ddy, ddy)));
if (dis > 0.0f) { ddx = fmul(fdiv(ddx, dis), c[5]); ddy = fmul(fdiv(ddy, dis), c[5]); }
nx = toroidal_transform(fadd(sx, ddx), w);
ny = toroidal_transform(fadd(sy, ddy), w);          // `width` for both axes, :146-147
na = ddx; nb = ddy;
"""

# A model the library does not ship: every agent drifts towards the centroid of its neighbours at constant speed
# c[0]; an agent with more than c[1] neighbours stops (dies).  Written with ordinary operators.
CENTROID_PAIR = """
if (sid != oid) { cnt += 1; acc[0] = acc[0] + dx; acc[1] = acc[1] + dy; }
"""
CENTROID_FINISH = """
float mx = 0.f, my = 0.f;
if (cnt > 0) { mx = -acc[0] / (float)cnt; my = -acc[1] / (float)cnt; }
float len = sqrtf(mx * mx + my * my);
if (len > 0.f) { mx = mx / len * c[0]; my = my / len * c[0]; }
nx = toroidal_transform(sx + mx, w);
ny = toroidal_transform(sy + my, h);
na = mx; nb = my;
stopped = (float)cnt > c[1];
"""
