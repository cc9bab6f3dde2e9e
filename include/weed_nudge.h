/*
 * weed_nudge.h — the documented deviation for exactly-coincident colliders.
 *
 * The reference separates two colliders at distance 0 along a direction drawn from the
 * worker's sequential mulberry32 stream (src/workers/physics_worker.js:460-507,
 * src/core/utils.js:333-342).  A sequential stream has no meaning in a parallel sweep, so
 * the GPU path (and the oracle's J-order mode, which defines what the GPU must match)
 * derives the direction from a counter-based hash of (min id, max id, frame, substep,
 * seed).  Both sides of a pair evaluate the same function, so the nudge stays
 * antisymmetric.  Everything here is integer arithmetic or individually rounded binary64
 * multiplies/adds (no FMA), so gcc (-ffp-contract=off) and nvcc give identical bits.
 */
#ifndef WEED_NUDGE_H
#define WEED_NUDGE_H

#include <stdint.h>

#ifdef __CUDA_ARCH__
#define WEED_HD __host__ __device__ __forceinline__
#define WEED_MUL(a, b) __dmul_rn((a), (b))
#define WEED_ADD(a, b) __dadd_rn((a), (b))
#else
#ifdef __CUDACC__
#define WEED_HD __host__ __device__ __forceinline__
#else
#define WEED_HD static inline
#endif
#define WEED_MUL(a, b) ((a) * (b))
#define WEED_ADD(a, b) ((a) + (b))
#endif

WEED_HD uint32_t weed_mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu;
  h ^= h >> 13; h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

WEED_HD uint32_t weed_nudge_hash(uint32_t lo_id, uint32_t hi_id, uint32_t frame,
                                 uint32_t substep, uint32_t seed) {
  uint32_t h = weed_mix32(lo_id ^ 0x9e3779b9u);
  h = weed_mix32(h ^ (hi_id * 0x7feb352du + 0x846ca68bu));
  h = weed_mix32(h ^ (frame * 0x2c1b3c6du + substep * 0x297a2d39u));
  h = weed_mix32(h ^ seed);
  return h;
}

/* (cos, sin) of the angle 2*pi*h/2^32, by quadrant reduction + a fixed Taylor polynomial. */
WEED_HD void weed_nudge_dir(uint32_t h, double* c_out, double* s_out) {
  const uint32_t hq = h + 0x20000000u;           /* + 1/8 turn, wraps mod 1 turn */
  const uint32_t q = hq >> 30;                   /* quadrant 0..3 */
  const double rem = (double)(hq & 0x3FFFFFFFu) * (1.0 / 1073741824.0); /* [0,1) exact */
  const double a = WEED_MUL(WEED_ADD(rem, -0.5), 1.5707963267948966);    /* [-pi/4,pi/4) */
  const double a2 = WEED_MUL(a, a);
  /* cos a = 1 - a2/2 + a2^2/24 - ... (to a^14), Horner */
  double c = -1.1470745597729725e-11;            /* -1/14! */
  c = WEED_ADD(WEED_MUL(c, a2), 2.08767569878681e-09);    /*  1/12! */
  c = WEED_ADD(WEED_MUL(c, a2), -2.755731922398589e-07);  /* -1/10! */
  c = WEED_ADD(WEED_MUL(c, a2), 2.48015873015873e-05);    /*  1/8!  */
  c = WEED_ADD(WEED_MUL(c, a2), -0.001388888888888889);   /* -1/6!  */
  c = WEED_ADD(WEED_MUL(c, a2), 0.041666666666666664);    /*  1/4!  */
  c = WEED_ADD(WEED_MUL(c, a2), -0.5);
  c = WEED_ADD(WEED_MUL(c, a2), 1.0);
  /* sin a = a (1 - a2/6 + a2^2/120 - ... (to a^15)) */
  double s = -7.647163731819816e-13;             /* -1/15! */
  s = WEED_ADD(WEED_MUL(s, a2), 1.6059043836821613e-10);  /*  1/13! */
  s = WEED_ADD(WEED_MUL(s, a2), -2.505210838544172e-08);  /* -1/11! */
  s = WEED_ADD(WEED_MUL(s, a2), 2.7557319223985893e-06);  /*  1/9!  */
  s = WEED_ADD(WEED_MUL(s, a2), -0.0001984126984126984);  /* -1/7!  */
  s = WEED_ADD(WEED_MUL(s, a2), 0.008333333333333333);    /*  1/5!  */
  s = WEED_ADD(WEED_MUL(s, a2), -0.16666666666666666);    /* -1/3!  */
  s = WEED_ADD(WEED_MUL(s, a2), 1.0);
  s = WEED_MUL(s, a);
  /* the quadrant centre is q*pi/2 - pi/4 + pi/4 ... : angle = q*pi/2 + a - pi/4 + pi/4.
     With the +1/8 turn offset the centre of quadrant q sits at angle q*pi/2, so rotate. */
  double cc, ss;
  if (q == 0)      { cc = c;  ss = s;  }
  else if (q == 1) { cc = -s; ss = c;  }
  else if (q == 2) { cc = -c; ss = -s; }
  else             { cc = s;  ss = -c; }
  *c_out = cc; *s_out = ss;
}

#endif /* WEED_NUDGE_H */
