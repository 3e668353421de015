/*
 * epi_oracle.c — CPU restatement of the epi-mcmc log-posterior path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.  The product
 * (libepimcmc.so) never links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED versus the reference's executed output: the reference is R
 * (absent here), its ODE solver is deSolve::lsoda (CRAN, un-vendored, version
 * unpinned) and its sampler is drjacoby (un-vendored); the reference ships no
 * tests, seeds or golden vectors.  What pins this file instead:
 *   - the compiled RHS is checked against the reference's own model-agen.cpp,
 *     built from where it lies into oracle/_ref/ (oracle/Makefile, tests);
 *   - the integrator is a fixed-step classical RK4 with `substeps` steps per
 *     day, cut additionally at the beta(t) breakpoints (see ode_rk4); tests
 *     cross-check it against scipy LSODA(rtol=atol=1e-6,max_step=1), i.e.
 *     deSolve's defaults, to solver tolerance;
 *   - pgamma/pnorm/dnbinom/dnorm/dlnorm are cross-checked against
 *     scipy.special / scipy.stats / mpmath.
 *
 * Every function cites the reference lines it follows (paths relative to the
 * reference root).  R vectors are 1-based: arrays here are allocated with one
 * spare leading slot and indexed 1-based so that index arithmetic reads like
 * the R source.  NA is represented by NaN.  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/epimcmc.h"
#include "epi_oracle.h"

#define NA_REAL NAN
#define INITIAL 1.0 /* Initial <- 1, model-age-groups2q.R:6 / model-simple-ode-drj-cmp.R:6 */

/* ------------------------------------------------------------------ nmath -- */

/* Regularised lower incomplete gamma P(a, x): restates R's
 * pgamma(q, shape, scale) = P(shape, q/scale) (used at
 * R/models/model-age-groups2q.R:31-32).  Series for x < a+1, modified Lentz
 * continued fraction otherwise (Numerical-Recipes style, iterated to 1e-17). */
double oracle_pgamma(double q, double shape, double scale) {
  if (isnan(q) || isnan(shape) || isnan(scale)) return NA_REAL;
  double x = q / scale, a = shape;
  if (x <= 0.0) return 0.0;
  double lpre = -x + a * log(x) - lgamma(a);
  if (x < a + 1.0) {
    double ap = a, del = 1.0 / a, sum = del;
    for (int n = 0; n < 100000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return sum * exp(lpre);
  } else {
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 100000; ++i) {
      double an = -(double)i * ((double)i - a);
      b += 2.0;
      d = an * d + b;
      if (fabs(d) < tiny) d = tiny;
      c = b + an / c;
      if (fabs(c) < tiny) c = tiny;
      d = 1.0 / d;
      double del = d * c;
      h *= del;
      if (fabs(del - 1.0) < 1e-17) break;
    }
    return 1.0 - exp(lpre) * h;
  }
}

/* pnorm(q, mean, sd), used at R/models/model-simple-ode-drj-cmp.R:25-26 */
double oracle_pnorm(double q, double mean, double sd) {
  double z = (q - mean) / sd;
  return 0.5 * erfc(-z / M_SQRT2);
}

#define LN_SQRT_2PI 0.918938533204672741780329736406 /* R's M_LN_SQRT_2PI */

/* dnorm(x, mean, sd, log=TRUE): R nmath dnorm4 */
double oracle_dnorm_log(double x, double mean, double sd) {
  double z = (x - mean) / sd;
  return -(LN_SQRT_2PI + 0.5 * z * z + log(sd));
}

/* dlnorm(x, meanlog, sdlog, log=TRUE): R nmath dlnorm */
double oracle_dlnorm_log(double x, double meanlog, double sdlog) {
  if (isnan(x)) return NA_REAL;
  if (x <= 0) return -INFINITY;
  double y = (log(x) - meanlog) / sdlog;
  return -(LN_SQRT_2PI + 0.5 * y * y + log(x * sdlog));
}

/* dnbinom(x, mu=, size=, log=TRUE): R nmath dnbinom_mu, in its closed form
 * lgamma(x+size)-lgamma(size)-lgamma(x+1)+size*log(size/(size+mu))+x*log(mu/(size+mu))
 * with R's x == 0 branch. */
double oracle_dnbinom_mu_log(double x, double size, double mu) {
  if (isnan(x) || isnan(size) || isnan(mu)) return NA_REAL;
  if (x < 0 || !isfinite(x) || fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x))) return -INFINITY;
  x = nearbyint(x);
  if (x == 0)
    return size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu)));
  /* the data-only lgamma part cancels heavily for large x: long double keeps it
   * at the accuracy R's saddle-point dbinom_raw() has.  It depends on (x, size) only, i.e. on the data: a small
   * per-thread direct-mapped cache keeps the CPU baseline from paying three lgammal() per term and evaluation
   * (the GPU path folds the same constants into its term table once, at handle creation). */
  static __thread struct { double x, size, lg; int used; } cache[4096];
  union { double d; uint64_t u; } kx = {x}, ks = {size};
  const unsigned slot = (unsigned)((kx.u * 0x9E3779B97F4A7C15ull ^ ks.u * 0xC2B2AE3D27D4EB4Full) >> 52);
  double lg;
  if (cache[slot].used && cache[slot].x == x && cache[slot].size == size) {
    lg = cache[slot].lg;
  } else {
    lg = (double)(lgammal((long double)x + size) - lgammal((long double)size) - lgammal((long double)x + 1.0L));
    cache[slot].x = x; cache[slot].size = size; cache[slot].lg = lg; cache[slot].used = 1;
  }
  return lg + size * log(size / (size + mu)) + x * log(mu / (size + mu));
}

static double pmax_(double lo, double v) { /* R pmax propagates NA */
  if (isnan(v)) return NA_REAL;
  return v < lo ? lo : v;
}
static double pmin_(double hi, double v) {
  if (isnan(v)) return NA_REAL;
  return v > hi ? hi : v;
}

/* ------------------------------------------------------- latency profiles -- */

typedef struct {
  int kbegin, kend; /* R's (non-positive) result$kbegin / result$kend */
  int n;            /* length(result$values) */
  double values[EPI_MAX_TAPS + 2]; /* 1-based */
} profile_t;

static void jitter_weights(double *v1, int n); /* conditioning probe, see oracle_set_jitter */

/* calcGammaProfile(mean, sd): R/models/model-age-groups2q.R:17-40.
 * Returns 0, or -1 if the profile needs more than EPI_MAX_TAPS taps. */
static int calc_gamma_profile(double mean, double sd, profile_t *r) {
  double scale = sd * sd / mean;
  double shape = mean / scale;
  double kb = fmax(0.0, ceil(mean - sd * 3));
  double ke = fmax(kb + 1, floor(mean + sd * 3));
  if (!(ke - kb + 1 <= EPI_MAX_TAPS) || !(kb >= 0)) return -1;
  int kbegin = (int)kb, kend = (int)ke;
  r->kbegin = -kend;
  r->kend = -kbegin;
  r->n = kend - kbegin + 1; /* numeric(L) grown to L+1 by the loop, :28-34 */
  double sum = 0;
  int i = 1;
  for (int k = kbegin; k <= kend; ++k) {
    r->values[i] = oracle_pgamma(k - 0.5, shape, scale) - oracle_pgamma(k + 0.5, shape, scale);
    sum += r->values[i];
    ++i;
  }
  for (i = 1; i <= r->n; ++i) r->values[i] = r->values[i] / sum;
  for (i = 1; i <= r->n / 2; ++i) { /* rev(), :37 */
    double t = r->values[i];
    r->values[i] = r->values[r->n + 1 - i];
    r->values[r->n + 1 - i] = t;
  }
  jitter_weights(r->values, r->n);
  return 0;
}

/* calcConvolveProfile(latency, latency_sd): R/models/model-simple-ode-drj-cmp.R:15-34
 * (called with a NEGATIVE latency, :56-57; not reversed). */
static int calc_convolve_profile(double latency, double latency_sd, profile_t *r) {
  double ke = fmin(0.0, ceil(latency + latency_sd * 1.5));
  double kb = fmin(ke - 1, floor(latency - latency_sd * 1.5));
  if (!(ke - kb + 1 <= EPI_MAX_TAPS) || !(ke <= 0)) return -1;
  int kbegin = (int)kb, kend = (int)ke;
  r->kbegin = kbegin;
  r->kend = kend;
  r->n = kend - kbegin + 1;
  double sum = 0;
  int i = 1;
  for (int k = kbegin; k <= kend; ++k) {
    r->values[i] = oracle_pnorm(k - 0.5, latency, latency_sd) - oracle_pnorm(k + 0.5, latency, latency_sd);
    sum += r->values[i];
    ++i;
  }
  for (i = 1; i <= r->n; ++i) r->values[i] = r->values[i] / sum;
  jitter_weights(r->values, r->n);
  return 0;
}

/* convolute(values, i1, i2, profile): R/models/model-age-groups2q.R:42-45 =
 * stats::filter(values, profile$values, method="convolution", sides=1)
 *   [(i1 + profile$kend):(i2 + profile$kend)]
 * filter (R's C cfilter, sides=1): out[i] = sum_{j=1..n} f[j]*x[i-j+1], plain
 * double accumulation in ascending j, NA while i < n.  x is 1-based length T.
 * out is 1-based, out[k] for k = 1..(i2-i1+1). */
static void convolute(const double *x, int T, int i1, int i2, const profile_t *p, double *out) {
  for (int t = i1; t <= i2; ++t) {
    int i = t + p->kend;
    double z;
    if (i < p->n || i > T || i < 1) {
      z = NA_REAL;
    } else {
      z = 0;
      for (int j = 1; j <= p->n; ++j) {
        double tmp = x[i - j + 1];
        if (isnan(tmp)) { z = NA_REAL; break; }
        z += p->values[j] * tmp;
      }
    }
    out[t - i1 + 1] = z;
  }
}

/* ------------------------------------------------------------- RHS + RK4 -- */

/* Piecewise schedule with INDEX priority: the highest k with t > tk wins
 * (R/models/model-agen.cpp:73-100).  kind[k]: 1 = linear towards k+1, 0 = hold. */
typedef struct {
  int nbp;
  double t[16];
  int kind[16];
} sched_t;

#define SEG_LITERAL (-2) /* choose the segment from t itself, exactly as the reference RHS does */

/* highest k with t > tk, or -1 (R/models/model-agen.cpp:78-99) */
static int active_segment(const sched_t *s, double t) {
  for (int k = s->nbp - 1; k >= 0; --k)
    if (t > s->t[k]) return k;
  return -1;
}

static double interpolate(const sched_t *s, int seg, double t, const double *v) {
  const int k = seg == SEG_LITERAL ? active_segment(s, t) : seg;
  if (k < 0) return v[0];
  if (s->kind[k]) return v[k] + (t - s->t[k]) / (s->t[k + 1] - s->t[k]) * (v[k + 1] - v[k]);
  return v[k];
}

typedef struct {
  /* age flavour, R/models/model-agen.cpp:7-51 */
  double Ny, No, a1, a2, gamma;
  double eta, tuncertain, funcertain; /* model-ager.cpp:12-14 (age-2x only) */
  sched_t s;
  double by[16], bo[16], byo[16];
  /* simple / Ne flavour (inferred RHS; model-simple-ode-cmp.cpp is absent —
   * parameters from R/models/model-simple-ode-drj-cmp.R:75-78,98-102) */
  double N, a, ldts, ldte, beta0, Tinf0, betat, Tinft;
  double Ne; /* Ne flavour: Nef*N; simple: == N */
  int flavour;
} rhs_t;

/* derivs(): R/models/model-agen.cpp:104-163 (states Sy,E1y,E2y,Iy,Ry,So,E1o,E2o,Io,Ro) */
static void derivs_age(const rhs_t *p, int seg, double t, const double *y, double *ydot) {
  const double Sy = y[0], E1y = y[1], E2y = y[2], Iy = y[3], Ry = y[4];
  const double So = y[5], E1o = y[6], E2o = y[7], Io = y[8], Ro = y[9];
  const double Ny = p->Ny, No = p->No;
  if (Sy < 0 || E1y < 0 || E2y < 0 || Iy < 0 || Ry < 0 || Sy > Ny || E1y > Ny || E2y > Ny ||
      Iy > Ny || Ry > Ny || So < 0 || E1o < 0 || E2o < 0 || Io < 0 || Ro < 0 || So > No ||
      E1o > No || E2o > No || Io > No || Ro > No) {
    for (int i = 0; i < 10; ++i) ydot[i] = 0;
    return;
  }
  const double betay = interpolate(&p->s, seg, t, p->by);
  const double betao = interpolate(&p->s, seg, t, p->bo);
  const double betayo = interpolate(&p->s, seg, t, p->byo);

  const double ygot_infected = (betay * Iy) / Ny * Sy;
  const double ygot_latent2 = p->a1 * E1y;
  const double ygot_infectious = p->a2 * E2y;
  const double ygot_removed = p->gamma * Iy;

  const double ogot_infected = ((betao * Io) / No + (betayo * Iy) / Ny) * So;
  const double ogot_latent2 = p->a1 * E1o;
  const double ogot_infectious = p->a2 * E2o;
  const double ogot_removed = p->gamma * Io;

  ydot[0] = -ygot_infected;
  ydot[1] = ygot_infected - ygot_latent2;
  ydot[2] = ygot_latent2 - ygot_infectious;
  ydot[3] = ygot_infectious - ygot_removed;
  ydot[4] = ygot_removed;
  ydot[5] = -ogot_infected;
  ydot[6] = ogot_infected - ogot_latent2;
  ydot[7] = ogot_latent2 - ogot_infectious;
  ydot[8] = ogot_infectious - ogot_removed;
  ydot[9] = ogot_removed;
}

/* derivs() of R/models/model-ager.cpp:135-214: waning immunity (eta * R flows back into S), the younger group's
 * infections run on the effective susceptibles Syeff = max(0, 0.8 Ny - (Ny - Sy)), the three beta curves pass
 * through uncertain() (:126-132).  yout (optional, 9 values): Re, Rt, y.Re, o.Re, y.i, o.i, betay, betao, betayo. */
static void derivs_ager(const rhs_t *p, int seg, double t, const double *y, double *ydot, double *yout) {
  const double Sy = y[0], E1y = y[1], E2y = y[2], Iy = y[3], Ry = y[4];
  const double So = y[5], E1o = y[6], E2o = y[7], Io = y[8], Ro = y[9];
  const double Ny = p->Ny, No = p->No;
  if (Sy < 0 || E1y < 0 || E2y < 0 || Iy < 0 || Ry < 0 || Sy > Ny || E1y > Ny || E2y > Ny ||
      Iy > Ny || Ry > Ny || So < 0 || E1o < 0 || E2o < 0 || Io < 0 || Ro < 0 || So > No ||
      E1o > No || E2o > No || Io > No || Ro > No) {
    if (ydot) for (int i = 0; i < 10; ++i) ydot[i] = 0;
    if (yout) for (int i = 0; i < 9; ++i) yout[i] = NA_REAL; /* left stale by the reference: reported as NA */
    return;
  }
  const double mbetay = interpolate(&p->s, seg, t, p->by);
  const double mbetao = interpolate(&p->s, seg, t, p->bo);
  const double mbetayo = interpolate(&p->s, seg, t, p->byo);
  const double betay = t > p->tuncertain ? mbetay * p->funcertain : mbetay;
  const double betao = t > p->tuncertain ? mbetao * p->funcertain : mbetao;
  const double betayo = t > p->tuncertain ? mbetayo * p->funcertain : mbetayo;

  const double Syeff = fmax(0.0, Ny * 0.8 - (Ny - Sy));

  const double ygot_infected = (betay * Iy) / Ny * Syeff;
  const double ygot_latent2 = p->a1 * E1y;
  const double ygot_infectious = p->a2 * E2y;
  const double ygot_removed = p->gamma * Iy;
  const double ygot_reverted = p->eta * Ry;

  const double ogot_infected = ((betao * Io) / No + (betayo * Iy) / Ny) * So;
  const double ogot_latent2 = p->a1 * E1o;
  const double ogot_infectious = p->a2 * E2o;
  const double ogot_removed = p->gamma * Io;
  const double ogot_reverted = p->eta * Ro;

  if (ydot) {
    ydot[0] = -ygot_infected + ygot_reverted;
    ydot[1] = ygot_infected - ygot_latent2;
    ydot[2] = ygot_latent2 - ygot_infectious;
    ydot[3] = ygot_infectious - ygot_removed;
    ydot[4] = ygot_removed - ygot_reverted;
    ydot[5] = -ogot_infected + ogot_reverted;
    ydot[6] = ogot_infected - ogot_latent2;
    ydot[7] = ogot_latent2 - ogot_infectious;
    ydot[8] = ogot_infectious - ogot_removed;
    ydot[9] = ogot_removed - ogot_reverted;
  }
  if (yout) {
    yout[0] = 1 / p->gamma * (ygot_infected + ogot_infected) / (Iy + Io);
    yout[1] = yout[0] * (Ny + No) / (Sy + So);
    yout[2] = 1 / p->gamma * ((betay * Iy) / Ny * Sy + (betayo * Iy) / Ny * So) / Iy;
    yout[3] = 1 / p->gamma * (betao * Io) / No * So / Io;
    yout[4] = ygot_infected;
    yout[5] = ogot_infected;
    yout[6] = betay;
    yout[7] = betao;
    yout[8] = betayo;
  }
}

/* Simple-flavour derivs.  INFERRED: model-simple-ode-cmp.cpp is not in the
 * snapshot.  Restated from its call site (parms N,a,ldts,ldte,beta0,Tinf0,
 * betat,Tinft; states S,E,I,R: R/models/model-simple-ode-drj-cmp.R:75-86) by
 * analogy with the single-E SEIR of R/models/model-age3.cpp:142-184 and the
 * strict `t > tk` schedule + bounds guard every shipped RHS uses
 * (R/models/model-agen.cpp:73-124).  Ne flavour (README.md:78-95): infectious
 * and susceptible populations are both taken relative to Ne = Nef*N, with the
 * effective-susceptible idiom of R/models/model-ager.cpp:176-178. */
static void derivs_simple(const rhs_t *p, int seg, double t, const double *y, double *ydot) {
  const double S = y[0], E = y[1], I = y[2], R = y[3];
  const double N = p->N;
  if (S < 0 || E < 0 || I < 0 || R < 0 || S > N || E > N || I > N || R > N) {
    for (int i = 0; i < 4; ++i) ydot[i] = 0;
    return;
  }
  double beta, Tinf;
  if (seg == SEG_LITERAL) seg = t > p->ldte ? 1 : t > p->ldts ? 0 : -1;
  if (seg == 1) {
    beta = p->betat;
    Tinf = p->Tinft;
  } else if (seg == 0) {
    double f = (t - p->ldts) / (p->ldte - p->ldts);
    beta = p->beta0 + f * (p->betat - p->beta0);
    Tinf = p->Tinf0 + f * (p->Tinft - p->Tinf0);
  } else {
    beta = p->beta0;
    Tinf = p->Tinf0;
  }
  double got_infected;
  if (p->flavour == EPI_FLAVOUR_NE) {
    const double Seff = fmax(0.0, p->Ne - (N - S));
    got_infected = (beta * I) / p->Ne * Seff;
  } else {
    got_infected = (beta * I) / N * S;
  }
  const double got_infectious = p->a * E;
  const double got_removed = I / Tinf;
  ydot[0] = -got_infected;
  ydot[1] = got_infected - got_infectious;
  ydot[2] = got_infectious - got_removed;
  ydot[3] = got_removed;
}

static void jitter_rows(double *out, int nday, int neq);

static void derivs(const rhs_t *p, int neq, int seg, double t, const double *y, double *ydot) {
  if (neq == 10 && p->flavour == EPI_FLAVOUR_AGE2X)
    derivs_ager(p, seg, t, y, ydot, 0);
  else if (neq == 10)
    derivs_age(p, seg, t, y, ydot);
  else
    derivs_simple(p, seg, t, y, ydot);
}

/* The nout = 4 output variables Re, Rt, y.Re, o.Re that derivs() fills next to ydot
 * (R/models/model-agen.cpp:159-163, model-agef.cpp:146-150), evaluated the way deSolve does for
 * the output matrix: one derivs() call at the output time with the LITERAL `t > tk` schedule.
 * When the bounds guard returns early the reference leaves yout untouched (stale values of an
 * earlier call): reported as NA here. */
static void yout_age(const rhs_t *p, double t, const double *y, double *yout) {
  const double Sy = y[0], E1y = y[1], E2y = y[2], Iy = y[3], Ry = y[4];
  const double So = y[5], E1o = y[6], E2o = y[7], Io = y[8], Ro = y[9];
  const double Ny = p->Ny, No = p->No;
  if (Sy < 0 || E1y < 0 || E2y < 0 || Iy < 0 || Ry < 0 || Sy > Ny || E1y > Ny || E2y > Ny ||
      Iy > Ny || Ry > Ny || So < 0 || E1o < 0 || E2o < 0 || Io < 0 || Ro < 0 || So > No ||
      E1o > No || E2o > No || Io > No || Ro > No) {
    for (int i = 0; i < 4; ++i) yout[i] = NA_REAL;
    return;
  }
  const double betay = interpolate(&p->s, SEG_LITERAL, t, p->by);
  const double betao = interpolate(&p->s, SEG_LITERAL, t, p->bo);
  const double betayo = interpolate(&p->s, SEG_LITERAL, t, p->byo);
  const double ygot_infected = (betay * Iy) / Ny * Sy;
  const double ogot_infected = ((betao * Io) / No + (betayo * Iy) / Ny) * So;
  yout[0] = 1 / p->gamma * (ygot_infected + ogot_infected) / (Iy + Io);
  yout[1] = yout[0] * (Ny + No) / (Sy + So);
  yout[2] = 1 / p->gamma * ((betay * Iy) / Ny * Sy + (betayo * Iy) / Ny * So) / Iy;
  yout[3] = 1 / p->gamma * (betao * Io) / No * So / Io;
}

/* simple flavour, nout = 2 (Re, Rt).  INFERRED like derivs_simple (the RHS file is absent):
 * Re = Tinf * beta * S / N, Rt = beta * Tinf; Ne flavour: S/N -> Seff/Ne. */
static void yout_simple(const rhs_t *p, double t, const double *y, double *yout) {
  const double S = y[0], E = y[1], I = y[2], R = y[3];
  const double N = p->N;
  if (S < 0 || E < 0 || I < 0 || R < 0 || S > N || E > N || I > N || R > N) {
    yout[0] = yout[1] = NA_REAL;
    return;
  }
  double beta, Tinf;
  if (t > p->ldte) {
    beta = p->betat; Tinf = p->Tinft;
  } else if (t > p->ldts) {
    double f = (t - p->ldts) / (p->ldte - p->ldts);
    beta = p->beta0 + f * (p->betat - p->beta0);
    Tinf = p->Tinf0 + f * (p->Tinft - p->Tinf0);
  } else {
    beta = p->beta0; Tinf = p->Tinf0;
  }
  if (p->flavour == EPI_FLAVOUR_NE)
    yout[0] = Tinf * beta * fmax(0.0, p->Ne - (N - S)) / p->Ne;
  else
    yout[0] = Tinf * beta * S / N;
  yout[1] = beta * Tinf;
}

/* yout rows for an ode() output matrix: out is [nday][neq] at times t0 .. t0+nday-1 */
static void yout_rows(const rhs_t *p, int neq, const double *out, int t0, int nday, double *yo) {
  for (int d = 0; d < nday; ++d) {
    if (neq == 10) yout_age(p, (double)(t0 + d), out + (size_t)d * 10, yo + (size_t)d * 4);
    else yout_simple(p, (double)(t0 + d), out + (size_t)d * 4, yo + (size_t)d * 2);
  }
}

/* Replaces deSolve::ode(Y, times = (t0):(t0+nday-1), ...)
 * (R/models/model-age-groups2q.R:163-165,211-213; LSODA, adaptive) by a
 * deterministic fixed-step scheme, the SAME scheme the CUDA kernels implement:
 *
 *   - each day is cut into m sub-steps [a, b], a = tday + s/m (b = tday + 1 for
 *     the last);
 *   - a sub-step that contains schedule breakpoints strictly inside (a, b) is
 *     cut again at those breakpoints (LSODA likewise never steps across a
 *     discontinuity unnoticed; without the cut RK4 degrades to first order at
 *     the beta(t) jumps/kinks and logl is off by O(0.1..1));
 *   - each piece [pa, pb] takes one classical RK4 step with the beta(t) segment
 *     frozen to the one active at its midpoint (index-priority rule of
 *     model-agen.cpp:78-99), the segment's formula being evaluated at the stage
 *     times pa, pa+hh/2, pb.
 *
 * State is reported at the integer times.  out is row-major [nday][neq], row
 * 0 = Y0 at t = t0. */
/* ---- mirrored arithmetic (tests only) -------------------------------------------------------------
 * oracle_set_mirror(1) makes ode_rk4 evaluate the SAME scheme with the operation order and the explicit fused
 * multiply-adds of the CUDA kernels (epi_mcmc_b200/csrc/epi_kernels.cu: rhs<>, rk4_try<>, select_segment<>):
 * curves pre-divided by the population they multiply, beta(t) = fma(t - tk, slope, b), every product-sum a
 * written-out fma, acc = k1 + 2 k2 + 2 k3 accumulated by fma, y' = fma(h/6, acc + k4, y).  Every operation is
 * correctly rounded IEEE on both sides, so the ODE states then agree BIT FOR BIT with the GPU and what is left of
 * a GPU <-> oracle difference comes from the latency profiles and the table-driven logarithms only.  The default
 * (mirror off) keeps the reference's own operation order, (beta * I) / N * S, which differs from the mirrored one by
 * a few ulp per stage — amplified by the cancellation in diff(N - filter(S)) where the epidemic has burnt out. */
static int g_mirror = 0;
void oracle_set_mirror(int on) { g_mirror = on; }

typedef struct { double tk, b0, b1, b2, s0, s1, s2; } mseg_t;

static int guard_hits(const rhs_t *p, int neq, const double *y) {
  if (neq == 10) {
    for (int i = 0; i < 5; ++i) if (y[i] < 0 || y[i] > p->Ny) return 1;
    for (int i = 5; i < 10; ++i) if (y[i] < 0 || y[i] > p->No) return 1;
    return 0;
  }
  for (int i = 0; i < 4; ++i) if (y[i] < 0 || y[i] > p->N) return 1;
  return 0;
}

/* select_segment<> of the kernels: segment k (highest breakpoint <= pa), curves divided by their population */
static mseg_t mirror_segment(const rhs_t *p, int neq, int k) {
  mseg_t g;
  memset(&g, 0, sizeof g);
  if (neq == 10) {
    const double inv0 = 1.0 / p->Ny, inv1 = 1.0 / p->No;
    const int kk = k < 0 ? 0 : k;
    g.b0 = p->by[kk]; g.b1 = p->bo[kk]; g.b2 = p->byo[kk];
    if (k >= 0 && p->s.kind[k]) {
      const double len = p->s.t[k + 1] - p->s.t[k];
      g.tk = p->s.t[k];
      g.s0 = (p->by[k + 1] - p->by[k]) / len; g.s1 = (p->bo[k + 1] - p->bo[k]) / len; g.s2 = (p->byo[k + 1] - p->byo[k]) / len;
    }
    g.b0 = g.b0 * inv0; g.s0 = g.s0 * inv0;
    g.b1 = g.b1 * inv1; g.s1 = g.s1 * inv1;
    g.b2 = g.b2 * inv0; g.s2 = g.s2 * inv0;
  } else {
    const double inv0 = 1.0 / (p->flavour == EPI_FLAVOUR_NE ? p->Ne : p->N);
    if (k < 0) { g.b0 = p->beta0; g.b1 = p->Tinf0; g.b2 = 1.0 / p->Tinf0; }
    else if (k == 1) { g.b0 = p->betat; g.b1 = p->Tinft; g.b2 = 1.0 / p->Tinft; }
    else {
      const double len = p->ldte - p->ldts;
      g.tk = p->ldts; g.b0 = p->beta0; g.b1 = p->Tinf0; g.b2 = 0.0;
      g.s0 = (p->betat - p->beta0) / len; g.s1 = (p->Tinft - p->Tinf0) / len;
      if (g.s1 == 0.0) g.b2 = 1.0 / p->Tinf0;
    }
    g.b0 = g.b0 * inv0; g.s0 = g.s0 * inv0;
  }
  return g;
}

static void derivs_mirror(const rhs_t *p, int neq, const mseg_t *g, double t, const double *y, double *dy) {
  if (guard_hits(p, neq, y)) { /* the reference's guard, model-agen.cpp:109-124 */
    for (int i = 0; i < neq; ++i) dy[i] = 0;
    return;
  }
  const double dt = t - g->tk;
  if (neq == 10 && p->flavour == EPI_FLAVOUR_AGE2X) { /* rhs<> of the kernels, ager branch */
    const double Sy = y[0], E1y = y[1], E2y = y[2], Iy = y[3], Ry = y[4], So = y[5], E1o = y[6], E2o = y[7], Io = y[8], Ro = y[9];
    const double betay = fma(dt, g->s0, g->b0), betao = fma(dt, g->s1, g->b1), betayo = fma(dt, g->s2, g->b2);
    const double Syeff = fmax(0.0, p->Ny * 0.8 - (p->Ny - Sy));
    const double yinf = (betay * Iy) * Syeff;
    const double oinf = fma(betao, Io, betayo * Iy) * So;
    const double yrev = p->eta * Ry, orev = p->eta * Ro;
    dy[0] = yrev - yinf;
    dy[1] = fma(-p->a1, E1y, yinf);
    dy[2] = fma(p->a1, E1y, -E2y);
    dy[3] = fma(-p->gamma, Iy, E2y);
    dy[4] = fma(p->gamma, Iy, -yrev);
    dy[5] = orev - oinf;
    dy[6] = fma(-p->a1, E1o, oinf);
    dy[7] = fma(p->a1, E1o, -E2o);
    dy[8] = fma(-p->gamma, Io, E2o);
    dy[9] = fma(p->gamma, Io, -orev);
  } else if (neq == 10) {
    const double Sy = y[0], E1y = y[1], E2y = y[2], Iy = y[3], So = y[5], E1o = y[6], E2o = y[7], Io = y[8];
    const double betay = fma(dt, g->s0, g->b0), betao = fma(dt, g->s1, g->b1), betayo = fma(dt, g->s2, g->b2);
    const double yinf = (betay * Iy) * Sy;
    const double oinf = fma(betao, Io, betayo * Iy) * So;
    dy[0] = -yinf;
    dy[1] = fma(-p->a1, E1y, yinf);
    dy[2] = fma(p->a1, E1y, -E2y);
    dy[3] = fma(-p->gamma, Iy, E2y);
    dy[4] = p->gamma * Iy;
    dy[5] = -oinf;
    dy[6] = fma(-p->a1, E1o, oinf);
    dy[7] = fma(p->a1, E1o, -E2o);
    dy[8] = fma(-p->gamma, Io, E2o);
    dy[9] = p->gamma * Io;
  } else {
    const double S = y[0], E = y[1], I = y[2];
    const double beta = fma(dt, g->s0, g->b0);
    const double rT = g->s1 == 0.0 ? g->b2 : 1.0 / fma(dt, g->s1, g->b1);
    double inf;
    if (p->flavour == EPI_FLAVOUR_NE) {
      const double Seff = fmax(0.0, p->Ne - (p->N - S));
      inf = (beta * I) * Seff;
    } else {
      inf = (beta * I) * S;
    }
    const double ir = I * rT;
    dy[0] = -inf;
    dy[1] = fma(-p->a, E, inf);
    dy[2] = fma(p->a, E, -ir);
    dy[3] = ir;
  }
}

/* one RK4 piece [pa, pb] in the kernels' arithmetic (rk4_try<>); hh2 / h6 are passed in because a regular sub-step
 * uses the kernel constants 0.5 * h and h / 6 */
static void rk4_piece_mirror(const rhs_t *p, int neq, int seg, double *y, double pa, double pb, double hh, double hh2,
                             double h6) {
  const mseg_t g = mirror_segment(p, neq, seg);
  const double tm = pa + hh2;
  double k[10], acc[10], yt[10];
  derivs_mirror(p, neq, &g, pa, y, k);
  for (int i = 0; i < neq; ++i) { acc[i] = k[i]; yt[i] = fma(hh2, k[i], y[i]); }
  derivs_mirror(p, neq, &g, tm, yt, k);
  for (int i = 0; i < neq; ++i) { acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh2, k[i], y[i]); }
  derivs_mirror(p, neq, &g, tm, yt, k);
  for (int i = 0; i < neq; ++i) { acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh, k[i], y[i]); }
  derivs_mirror(p, neq, &g, pb, yt, k);
  for (int i = 0; i < neq; ++i) y[i] = fma(h6, acc[i] + k[i], y[i]);
}

static void ode_rk4(const rhs_t *p, int neq, const double *Y0, int t0, int nday, int m, double *out) {
  double y[10], k1[10], k2[10], k3[10], k4[10], yt[10];
  const double h = 1.0 / (double)m;
  double brk[16];
  int nb;
  if (neq == 10) {
    nb = p->s.nbp;
    for (int k = 0; k < nb; ++k) brk[k] = p->s.t[k];
  } else {
    nb = 2;
    brk[0] = p->ldts;
    brk[1] = p->ldte;
  }
  memcpy(y, Y0, sizeof(double) * neq);
  memcpy(out, y, sizeof(double) * neq);
  for (int d = 1; d < nday; ++d) {
    const double tday = (double)(t0 + d - 1);
    for (int s = 0; s < m; ++s) {
      const double a = tday + (double)s * h;
      const double b = (s == m - 1) ? tday + 1.0 : tday + (double)(s + 1) * h;
      double pa = a;
      for (;;) {
        /* next cut: the smallest breakpoint strictly inside (pa, b), else b */
        double pb = b;
        for (int k = 0; k < nb; ++k)
          if (brk[k] > pa && brk[k] < pb) pb = brk[k];
        const double hh = pb - pa;
        const double tm = pa + 0.5 * hh;
        const int seg = neq == 10 ? active_segment(&p->s, tm) : (tm > p->ldte ? 1 : tm > p->ldts ? 0 : -1);
        if (g_mirror) {
          rk4_piece_mirror(p, neq, seg, y, pa, pb, hh, 0.5 * hh, hh / 6.0);
          if (pb >= b) break;
          pa = pb;
          continue;
        }
        derivs(p, neq, seg, pa, y, k1);
        for (int i = 0; i < neq; ++i) yt[i] = y[i] + 0.5 * hh * k1[i];
        derivs(p, neq, seg, tm, yt, k2);
        for (int i = 0; i < neq; ++i) yt[i] = y[i] + 0.5 * hh * k2[i];
        derivs(p, neq, seg, tm, yt, k3);
        for (int i = 0; i < neq; ++i) yt[i] = y[i] + hh * k3[i];
        derivs(p, neq, seg, pb, yt, k4);
        for (int i = 0; i < neq; ++i)
          y[i] = y[i] + hh / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
        if (pb >= b) break;
        pa = pb;
      }
    }
    memcpy(out + (size_t)d * neq, y, sizeof(double) * neq);
  }
  jitter_rows(out, nday, neq);
}

/* Conditioning probe (tests only): when set, every reported ODE row and every
 * latency-profile weight is multiplied by (1 + k*2^-52) with a pseudo-random
 * integer |k| <= ulps.  It
 * emulates the rounding noise any other correctly rounded FP64 implementation of
 * the same scheme carries after ~10^3 steps, so that |logl(jitter) - logl| measures
 * how ill-conditioned the REFERENCE algorithm (differences of N - filter(S)) is at
 * a given theta.  Never set outside tests/ and make_fixtures.py. */
static int g_jitter_ulps = 0;
static unsigned g_jitter_seed = 0;
void oracle_set_jitter(int ulps, unsigned seed) { g_jitter_ulps = ulps; g_jitter_seed = seed; }

static void jitter_rows(double *out, int nday, int neq) {
  if (g_jitter_ulps <= 0) return;
  for (int d = 1; d < nday; ++d)
    for (int i = 0; i < neq; ++i) {
      unsigned h = (unsigned)d * 2654435761u ^ (unsigned)(i + 1) * 40503u ^ g_jitter_seed * 2246822519u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
      const int k = (int)(h % (unsigned)(2 * g_jitter_ulps + 1)) - g_jitter_ulps;
      out[(size_t)d * neq + i] *= 1.0 + (double)k * 0x1p-52;
    }
}

static void jitter_weights(double *v1, int n) {
  if (g_jitter_ulps <= 0) return;
  for (int i = 1; i <= n; ++i) {
    unsigned h = (unsigned)i * 2654435761u ^ g_jitter_seed * 2246822519u ^ (unsigned)n * 97u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    const int k = (int)(h % (unsigned)(2 * g_jitter_ulps + 1)) - g_jitter_ulps;
    v1[i] *= 1.0 + (double)k * 0x1p-52;
  }
}

/* exported for the RHS cross-check against oracle/_ref (tests) */
void oracle_derivs_age(const double *parms45, double t, const double *y, double *ydot) {
  rhs_t p;
  memset(&p, 0, sizeof p);
  p.Ny = parms45[0]; p.No = parms45[1]; p.a1 = parms45[2]; p.a2 = parms45[3]; p.gamma = parms45[4];
  p.s.nbp = 10;
  static const int kind[10] = {1, 1, 1, 1, 0, 0, 0, 0, 1, 0};
  static const int tidx[10] = {5, 6, 7, 17, 21, 25, 29, 33, 37, 41};
  for (int k = 0; k < 10; ++k) {
    p.s.kind[k] = kind[k];
    p.s.t[k] = parms45[tidx[k]];
    int b = (k < 3) ? 8 + 3 * k : tidx[k] + 1;
    p.by[k] = parms45[b]; p.bo[k] = parms45[b + 1]; p.byo[k] = parms45[b + 2];
  }
  derivs_age(&p, SEG_LITERAL, t, y, ydot);
}

void oracle_yout_age(const double *parms45, double t, const double *y, double *yout) {
  rhs_t p;
  memset(&p, 0, sizeof p);
  p.Ny = parms45[0]; p.No = parms45[1]; p.a1 = parms45[2]; p.a2 = parms45[3]; p.gamma = parms45[4];
  p.s.nbp = 10;
  static const int kind[10] = {1, 1, 1, 1, 0, 0, 0, 0, 1, 0};
  static const int tidx[10] = {5, 6, 7, 17, 21, 25, 29, 33, 37, 41};
  for (int k = 0; k < 10; ++k) {
    p.s.kind[k] = kind[k];
    p.s.t[k] = parms45[tidx[k]];
    int b = (k < 3) ? 8 + 3 * k : tidx[k] + 1;
    p.by[k] = parms45[b]; p.bo[k] = parms45[b + 1]; p.byo[k] = parms45[b + 2];
  }
  yout_age(&p, t, y, yout);
}

/* the restated ager RHS on the reference's own 60-element parms vector (R/models/model-ager.cpp:4-66): ydot[10] and
 * the nout = 9 output variables */
void oracle_derivs_ager(const double *parms60, double t, const double *y, double *ydot, double *yout9) {
  rhs_t p;
  memset(&p, 0, sizeof p);
  p.flavour = EPI_FLAVOUR_AGE2X;
  p.Ny = parms60[0]; p.No = parms60[1]; p.a1 = parms60[2]; p.a2 = parms60[3]; p.gamma = parms60[4];
  p.eta = parms60[5]; p.tuncertain = parms60[6]; p.funcertain = parms60[7];
  p.s.nbp = 13;
  static const int kind[13] = {1, 1, 1, 1, 0, 0, 0, 0, 1, 0, 1, 1, 0};
  for (int k = 0; k < 13; ++k) {
    p.s.kind[k] = kind[k];
    const int ti = k < 3 ? 8 + k : 20 + 4 * (k - 3);
    const int b = k < 3 ? 11 + 3 * k : ti + 1;
    p.s.t[k] = parms60[ti];
    p.by[k] = parms60[b]; p.bo[k] = parms60[b + 1]; p.byo[k] = parms60[b + 2];
  }
  derivs_ager(&p, SEG_LITERAL, t, y, ydot, yout9);
}

/* same for the 8-breakpoint schedule of R/models/model-agef.cpp:7-43,66-87 (parms[37]):
 * segments 0..5 linear towards the next breakpoint, 6 and 7 hold */
void oracle_derivs_agef(const double *parms37, double t, const double *y, double *ydot) {
  rhs_t p;
  memset(&p, 0, sizeof p);
  p.Ny = parms37[0]; p.No = parms37[1]; p.a1 = parms37[2]; p.a2 = parms37[3]; p.gamma = parms37[4];
  p.s.nbp = 8;
  static const int kind[8] = {1, 1, 1, 1, 1, 1, 0, 0};
  static const int tidx[8] = {5, 6, 7, 17, 21, 25, 29, 33};
  for (int k = 0; k < 8; ++k) {
    p.s.kind[k] = kind[k];
    p.s.t[k] = parms37[tidx[k]];
    int b = (k < 3) ? 8 + 3 * k : tidx[k] + 1;
    p.by[k] = parms37[b]; p.bo[k] = parms37[b + 1]; p.byo[k] = parms37[b + 2];
  }
  derivs_age(&p, SEG_LITERAL, t, y, ydot);
}

void oracle_yout_agef(const double *parms37, double t, const double *y, double *yout) {
  rhs_t p;
  memset(&p, 0, sizeof p);
  p.Ny = parms37[0]; p.No = parms37[1]; p.a1 = parms37[2]; p.a2 = parms37[3]; p.gamma = parms37[4];
  p.s.nbp = 8;
  static const int kind[8] = {1, 1, 1, 1, 1, 1, 0, 0};
  static const int tidx[8] = {5, 6, 7, 17, 21, 25, 29, 33};
  for (int k = 0; k < 8; ++k) {
    p.s.kind[k] = kind[k];
    p.s.t[k] = parms37[tidx[k]];
    int b = (k < 3) ? 8 + 3 * k : tidx[k] + 1;
    p.by[k] = parms37[b]; p.bo[k] = parms37[b + 1]; p.byo[k] = parms37[b + 2];
  }
  yout_age(&p, t, y, yout);
}

/* ------------------------------------------------------------ helpers ----- */

static double *vec(int n, double fill) { /* 1-based vector of length n */
  double *v = (double *)malloc(sizeof(double) * (size_t)(n + 2));
  for (int i = 0; i <= n + 1; ++i) v[i] = fill;
  return v;
}
static double at(const double *v, int len, int i) { /* R indexing: beyond length -> NA */
  if (i < 1 || i > len) return NA_REAL;
  return v[i];
}

/* The three-slice rate map of R/models/model-age-groups2q.R:240-251 (same shape
 * repeated for the six series).  s2i is 1-based length P; rate 1-based length nr. */
static void rate_map(double *dst, const double *s2i, int pad, int P, int valid, double t1d,
                     const double *rate, int nr) {
  if (valid) {
    int t1 = (int)t1d;
    int t2 = (pad + P < t1 + nr - 1) ? pad + P : t1 + nr - 1;
    if (t1 > pad && t2 > t1) {
      for (int t = pad + 1; t <= t1; ++t) dst[t] = s2i[t - pad] * rate[1];
      for (int t = t1; t <= t2; ++t) dst[t] = s2i[t - pad] * rate[t - t1 + 1];
      for (int t = t2; t <= pad + P; ++t) dst[t] = s2i[t - pad] * rate[nr];
    }
  } else {
    for (int t = pad + 1; t <= pad + P; ++t) dst[t] = s2i[t - pad] * rate[1];
  }
}

/* sum(dnbinom(x, mu=pmax(floor, mu), size=size, log=T)) with R's recycling of
 * the three argument vectors to the longest (R/models/model-age-groups2q.R:514-526).
 * x = xv[x1..x2], mu = mv[m1..m2] (vector length mlen for NA-beyond-end), size
 * = sv[1..ns] (ns == 1 for a scalar). */
static double sum_dnbinom(const double *xv, int xlen, int x1, int x2, const double *mv, int mlen,
                          int m1, int m2, double floor_, const double *sv, int ns) {
  int lx = x2 - x1 + 1, lm = m2 - m1 + 1;
  if (lx <= 0 || lm <= 0) return NA_REAL; /* descending a:b ranges are rejected at create() */
  int len = lx > lm ? lx : lm;
  if (ns > len) len = ns;
  long double acc = 0; /* R's sum() accumulates in long double */
  for (int k = 0; k < len; ++k) {
    double x = at(xv, xlen, x1 + k % lx);
    double mu = pmax_(floor_, at(mv, mlen, m1 + k % lm));
    double size = sv[1 + k % ns];
    acc += oracle_dnbinom_mu_log(x, size, mu);
  }
  return (double)acc;
}

static void window(int offset, int n, int T, int *dstart, int *dend) {
  /* R/models/model-age-groups2q.R:497-512 */
  *dstart = offset;
  *dend = offset + n - 1;
  if (*dend > T) {
    *dend = T;
    *dstart = *dend - n + 1;
  }
  if (*dstart < 1) {
    *dstart = 1;
    *dend = *dstart + n - 1;
  }
}

/* ------------------------------------------------------------ age-2q ------ */

#define G_CONST 4.7
#define TINC1 3.0
#define TINC2 1.0

double oracle_age_gamma(void) { return 1.0 / ((G_CONST - (TINC1 + TINC2)) * 2.0); }

/* calculateModel(params, period): R/models/model-age-groups2q.R:50-337.
 * params is 1-based (params[1..45]).  Fills st (all series 1-based, length T). */
typedef struct {
  int pad, T, P, offset, na_profile;
  double *yS, *oS, *ycasei, *ocasei, *yhospi, *ohospi, *ydeadi, *odeadi;
  double *full; /* [P][10] state rows of the final ode() */
  double *yout; /* [P][4] Re, Rt, y.Re, o.Re of the final ode() */
} age_state_t;

static void age_state_free(age_state_t *s) {
  free(s->yS); free(s->oS); free(s->ycasei); free(s->ocasei);
  free(s->yhospi); free(s->ohospi); free(s->ydeadi); free(s->odeadi); free(s->full); free(s->yout);
  memset(s, 0, sizeof *s);
}

static int age_calculate_model(const epi_data *D, const double *params, int period, int m,
                               age_state_t *st) {
  memset(st, 0, sizeof *st);
  const double y_N = D->y_N, o_N = D->o_N;
  const double y_died_rate = D->y_ifr[0], o_died_rate = D->o_ifr[0]; /* :11-12 */
  const double a1 = 1 / TINC1, a2 = 1 / TINC2, gamma = oracle_age_gamma();

  /* :52-103 */
  double betay[10], betao[10], betayo[10];
  betay[0] = params[1]; betao[0] = params[2]; betayo[0] = params[3];
  betay[1] = params[4]; betao[1] = params[5]; betayo[1] = params[6];
  betay[2] = params[7]; betao[2] = params[8]; betayo[2] = params[9];
  const double ycase_rate = params[10], yhosp_rate = params[11], yhosp_latency = params[12],
               ydied_latency = params[13], ocase_rate = params[14], ohosp_rate = params[15],
               ohosp_latency = params[16], odied_latency = params[17], t0_morts = params[18],
               t0o = params[19], HLsd = params[20], DLsd = params[21], fyifr = params[22],
               fyhr = params[23], t3o = params[24], t4o = params[25];
  betay[3] = betay[2];
  betay[4] = params[26]; betao[4] = params[27]; betayo[4] = params[28];
  betao[3] = betao[4]; betayo[3] = betayo[4];
  betay[5] = params[29]; betao[5] = params[30]; betayo[5] = params[31];
  const double t6o = params[32];
  betay[6] = params[33]; betao[6] = params[34]; betayo[6] = params[35];
  const double t7o = params[36];
  betay[7] = params[37]; betao[7] = params[38]; betayo[7] = params[39];
  betay[8] = betay[7]; betao[8] = betao[7]; betayo[8] = betayo[7];
  betay[9] = params[40] * betay[8]; betao[9] = params[41] * betao[8]; betayo[9] = params[42] * betayo[8];
  const double ifrred = params[43], ycase_latency = params[44], ocase_latency = params[45];
  const double CLsd = 5;

  /* :106-111 */
  profile_t ycase, ocase, yhosp, ohosp, ydied, odied;
  if (calc_gamma_profile(ycase_latency, CLsd, &ycase) || calc_gamma_profile(ocase_latency, CLsd, &ocase) ||
      calc_gamma_profile(yhosp_latency, HLsd, &yhosp) || calc_gamma_profile(ohosp_latency, HLsd, &ohosp) ||
      calc_gamma_profile(ydied_latency, DLsd, &ydied) || calc_gamma_profile(odied_latency, DLsd, &odied))
    return -1;

  /* :113-114 */
  int padding = -ycase.kbegin;
  if (-ydied.kbegin > padding) padding = -ydied.kbegin;
  if (-ocase.kbegin > padding) padding = -ocase.kbegin;
  if (-odied.kbegin > padding) padding = -odied.kbegin;
  padding += 1;

  const int P = period, T = padding + period;
  st->pad = padding; st->T = T; st->P = P;
  st->yS = vec(T, y_N - INITIAL);   /* :118 */
  st->oS = vec(T, o_N);             /* :127 */
  st->ycasei = vec(T, 0); st->ocasei = vec(T, 0);
  st->yhospi = vec(T, 0); st->ohospi = vec(T, 0);
  st->ydeadi = vec(T, 0); st->odeadi = vec(T, 0);

  /* :143-158 */
  rhs_t rp;
  memset(&rp, 0, sizeof rp);
  rp.flavour = EPI_FLAVOUR_AGE2Q;
  rp.Ny = y_N; rp.No = o_N; rp.a1 = a1; rp.a2 = a2; rp.gamma = gamma;
  rp.s.nbp = 10;
  static const int kind[10] = {1, 1, 1, 1, 0, 0, 0, 0, 1, 0}; /* model-agen.cpp:78-99 */
  for (int k = 0; k < 10; ++k) {
    rp.s.kind[k] = kind[k];
    rp.s.t[k] = 1E10;
    rp.by[k] = betay[k]; rp.bo[k] = betao[k]; rp.byo[k] = betayo[k];
  }
  const double Y[10] = {y_N - INITIAL, INITIAL, 0, 0, 0, o_N, 0, 0, 0, 0};

  /* pass 1, :160-176 */
  const int period1 = 60;
  double *out = (double *)malloc(sizeof(double) * 10 * (size_t)(P > period1 ? P : period1));
  int out_rows = period1;
  ode_rk4(&rp, 10, Y, padding + 1, period1, m, out);
  const int T1 = (T > padding + period1) ? T : padding + period1; /* assignment may grow the vector */
  double *yS1 = vec(T1, y_N - INITIAL), *oS1 = vec(T1, o_N);
  for (int i = 1; i <= period1; ++i) {
    yS1[padding + i] = out[(i - 1) * 10 + 0];
    oS1[padding + i] = out[(i - 1) * 10 + 5];
  }
  double *s2 = vec(P > period1 ? P : period1, 0), *s2b = vec(period1, 0);
  convolute(yS1, T1, padding + 1, padding + period1, &ydied, s2);
  convolute(oS1, T1, padding + 1, padding + period1, &odied, s2b);
  double data_offset = EPI_INVALID_DATA_OFFSET; /* :178 */
  int lds1 = 0;
  for (int i = 1; i <= period1; ++i) { /* which(state$died > t0_morts), :180 (NA before padding+1) */
    double died = (y_N - s2[i]) * y_died_rate + (o_N - s2b[i]) * o_died_rate;
    if (died > t0_morts) { lds1 = padding + i; break; }
  }
  free(yS1); free(oS1); free(s2b);

  const int lo = D->lockdown_offset, ltp = D->lockdown_transition_period;
  if (lds1 > 0) { /* :181-214 */
    data_offset = lds1 - lo;
    double t2 = data_offset + lo + ltp;
    double t3 = t2 + t3o;
    double t4 = t3 + t4o;
    double t5 = data_offset + D->d5;
    double t6 = t5 + t6o;
    double t7 = t6 + t7o;
    double t8 = data_offset + D->d8;
    double t9 = t8 + 7;
    rp.s.t[0] = data_offset + lo + t0o;
    rp.s.t[1] = data_offset + lo + fmax(t0o, 0);
    rp.s.t[2] = t2; rp.s.t[3] = t3; rp.s.t[4] = t4; rp.s.t[5] = t5;
    rp.s.t[6] = t6; rp.s.t[7] = t7; rp.s.t[8] = t8; rp.s.t[9] = t9;
    ode_rk4(&rp, 10, Y, padding + 1, P, m, out);
    out_rows = P;
  }
  /* :216-223 — when pass 2 was skipped the 60-row `out` is recycled into the
   * P-length slices */
  st->full = (double *)malloc(sizeof(double) * 10 * (size_t)P);
  st->yout = (double *)malloc(sizeof(double) * 4 * (size_t)P);
  double *yo = (double *)malloc(sizeof(double) * 4 * (size_t)out_rows);
  yout_rows(&rp, 10, out, padding + 1, out_rows, yo);
  for (int i = 1; i <= P; ++i) {
    const double *row = out + (size_t)((i - 1) % out_rows) * 10;
    st->yS[padding + i] = row[0];
    st->oS[padding + i] = row[5];
    memcpy(st->full + (size_t)(i - 1) * 10, row, sizeof(double) * 10);
    memcpy(st->yout + (size_t)(i - 1) * 4, yo + (size_t)((i - 1) % out_rows) * 4, sizeof(double) * 4);
  }
  free(out); free(yo);

  /* :229-235 */
  const int nr = D->n_rate;
  double *gycr = vec(nr, 0), *gyifr = vec(nr, 0), *gyhr = vec(nr, 0), *gocr = vec(nr, 0),
         *goifr = vec(nr, 0), *gohr = vec(nr, 0);
  for (int i = 1; i <= nr; ++i) {
    const double yifr = D->y_ifr[i - 1], oifr = D->o_ifr[i - 1], yhr = D->y_hr[i - 1],
                 ohr = D->o_hr[i - 1], g = D->g[i - 1], gt = D->gtrimp[i - 1];
    gycr[i] = pmin_(1, ycase_rate * yifr * (1 + (fyifr - 1) * g));
    gyifr[i] = yifr * (1 + (fyifr - 1) * g) * (1 - ifrred * gt);
    gyhr[i] = yhosp_rate * yhr * (1 + (fyhr - 1) * g);
    gocr[i] = pmin_(1, ocase_rate * oifr);
    goifr[i] = oifr * (1 - ifrred * gt);
    gohr[i] = ohosp_rate * ohr;
  }

  const int valid = data_offset != EPI_INVALID_DATA_OFFSET;
  double *s2i = vec(P, 0);
  struct {
    const double *S; double Ng; const profile_t *prof; double t1; const double *rate; double *dst;
  } ser[6] = {
      /* :237-251 */ {st->yS, y_N, &ycase, data_offset + nearbyint(-ydied_latency + ycase_latency), gycr, st->ycasei},
      /* :254-268 */ {st->yS, y_N, &yhosp, data_offset + nearbyint(-ydied_latency + yhosp_latency), gyhr, st->yhospi},
      /* :270-283 */ {st->yS, y_N, &ydied, data_offset, gyifr, st->ydeadi},
      /* :286-299 */ {st->oS, o_N, &ocase, data_offset + nearbyint(-ydied_latency + ocase_latency), gocr, st->ocasei},
      /* :302-315 */ {st->oS, o_N, &ohosp, data_offset + nearbyint(-ydied_latency + ohosp_latency), gohr, st->ohospi},
      /* :317-330 */ {st->oS, o_N, &odied, data_offset, goifr, st->odeadi},
  };
  for (int q = 0; q < 6; ++q) {
    convolute(ser[q].S, T, padding + 1, padding + P, ser[q].prof, s2);
    for (int i = 1; i <= P; ++i) s2[i] = ser[q].Ng - s2[i];
    s2i[1] = s2[1];
    for (int i = 2; i <= P; ++i) s2i[i] = s2[i] - s2[i - 1]; /* c(s2[1], diff(s2)) */
    rate_map(ser[q].dst, s2i, padding, P, valid, ser[q].t1, ser[q].rate, nr);
  }
  free(s2); free(s2i);
  free(gycr); free(gyifr); free(gyhr); free(gocr); free(goifr); free(gohr);
  st->offset = (int)data_offset;
  return 0;
}

/* calclogp(params): R/models/model-age-groups2q.R:380-486 (params 1-based) */
static double age_calclogp(const epi_data *D, const double *params) {
  const double gamma = oracle_age_gamma();
  const double betay0 = params[1], betao0 = params[2], betayo0 = params[3], betay1 = params[4],
               betao1 = params[5], betayo1 = params[6], betay2 = params[7], betao2 = params[8],
               betayo2 = params[9], yhosp_rate = params[11], yhosp_latency = params[12],
               ydied_latency = params[13], ohosp_rate = params[15], ohosp_latency = params[16],
               odied_latency = params[17], t0o = params[19], HLsd = params[20], DLsd = params[21],
               fyifr = params[22], fyhr = params[23], t3o = params[24], t4o = params[25],
               betay4 = params[26], betao4 = params[27], betayo4 = params[28], betay5 = params[29],
               betao5 = params[30], betayo5 = params[31], t6o = params[32], betay6 = params[33],
               betao6 = params[34], betayo6 = params[35], t7o = params[36], betay7 = params[37],
               betao7 = params[38], betayo7 = params[39], fy9 = params[40], fo9 = params[41],
               fyo9 = params[42], ifrred = params[43], ycase_latency = params[44],
               ocase_latency = params[45];
  const double betay3 = betay2, betao8 = betao7;
  const double lo = D->lockdown_offset, ltp = D->lockdown_transition_period;
  double lp = 0;
  lp = lp + oracle_dnorm_log(t0o, 0, 10);
  lp = lp + oracle_dnorm_log(lo + ltp + t3o, D->d3, 10);
  lp = lp + oracle_dnorm_log(lo + ltp + t3o + t4o, D->d4, 10);
  lp = lp + oracle_dnorm_log(D->d5 + t6o, D->d6, 5);
  lp = lp + oracle_dnorm_log(D->d5 + t6o + t7o, D->d7, 4);
  const double SD = log(2), lSD = log(1.5);
  lp = lp + oracle_dlnorm_log(betay0 / betay1, 0, SD);
  lp = lp + oracle_dlnorm_log(betao0 / betao1, 0, SD);
  lp = lp + oracle_dlnorm_log(betayo0 / betayo1, 0, SD);
  lp = lp + oracle_dlnorm_log(betay1 / betay2, 0, SD);
  lp = lp + oracle_dlnorm_log(betao1 / betao2, 0, SD);
  lp = lp + oracle_dlnorm_log(betayo1 / betayo2, 0, SD);
  lp = lp + oracle_dlnorm_log(betay3 / betay4, 0, SD);
  lp = lp + oracle_dlnorm_log(betao2 / betao4, 0, SD);
  lp = lp + oracle_dlnorm_log(betayo2 / betayo4, 0, SD);
  lp = lp + oracle_dlnorm_log(betay5 / betay6, 0, SD);
  lp = lp + oracle_dlnorm_log(betao5 / betao6, 0, SD);
  lp = lp + oracle_dlnorm_log(betayo5 / betayo6, 0, SD);
  lp = lp + oracle_dlnorm_log(betay6 / betay7, 0, lSD);
  lp = lp + oracle_dlnorm_log(betao6 / betao7, 0, lSD);
  lp = lp + oracle_dlnorm_log(betayo6 / betayo7, 0, lSD);
  lp = lp + oracle_dnorm_log(betao7, 0.5 * gamma, 0.2 * gamma);
  lp = lp + oracle_dnorm_log(betao8 * fo9, 0.5 * gamma, 0.2 * gamma);
  lp = lp + oracle_dlnorm_log(fy9, log(0.85), log(1.07));
  lp = lp + oracle_dlnorm_log(fo9, log(0.85), log(1.07));
  lp = lp + oracle_dlnorm_log(fyo9, log(0.85), log(1.07));
  lp = lp + oracle_dnorm_log(HLsd, 5, 1);
  lp = lp + oracle_dnorm_log(DLsd, 5, 1);
  lp = lp + oracle_dnorm_log(ycase_latency, 10, 3);
  lp = lp + oracle_dnorm_log(ocase_latency, 10, 3);
  lp = lp + oracle_dnorm_log(yhosp_latency, 13, 3);
  lp = lp + oracle_dnorm_log(ohosp_latency, 13, 3);
  lp = lp + oracle_dnorm_log(ydied_latency, 21, 0.5);
  lp = lp + oracle_dnorm_log(odied_latency, 21, 0.5);
  lp = lp + oracle_dlnorm_log(fyifr, log(0.3), log(2));
  lp = lp + oracle_dlnorm_log(fyhr, log(0.3), log(2));
  lp = lp + oracle_dlnorm_log(ifrred, log(0.35), log(1.3));
  lp = lp + oracle_dlnorm_log(yhosp_rate, 0, lSD);
  lp = lp + oracle_dlnorm_log(ohosp_rate, 0, lSD);
  return lp;
}

/* calclogl(params, x): R/models/model-age-groups2q.R:488-602.  terms[0..5] =
 * y.loglC, o.loglC, loglH, loglHRatio, y.loglD, o.loglD */
static double age_calclogl(const epi_data *D, const double *params, int period, int m,
                           double *terms, int *offset_out, int *pad_out, int *status) {
  age_state_t st;
  if (age_calculate_model(D, params, period, m, &st)) {
    *status |= EPI_ST_UNSUPPORTED;
    for (int i = 0; i < 6; ++i) terms[i] = NA_REAL;
    *offset_out = 0; *pad_out = 0;
    return NA_REAL;
  }
  *offset_out = st.offset;
  *pad_out = st.pad;
  int offset = st.offset;
  if (offset == EPI_INVALID_DATA_OFFSET) { /* :493-494 */
    offset = 1;
    *status |= EPI_ST_INVALID_OFFSET;
  }
  const int T = st.T;
  int dstart, dend;
  const int r = D->d_reliable_cases, rh = D->d_reliable_hosp, nc = D->n_dcasei;
  double one[2];

  /* cases, :496-526 */
  window(offset, nc, T, &dstart, &dend);
  const double *yx = D->y_dcasei - 1, *ox = D->o_dcasei - 1; /* 1-based views */
  one[1] = D->case_nbinom_size1;
  double yC = sum_dnbinom(yx, nc, 1, r - 1, st.ycasei, T, dstart, dstart + r, 0.1, one, 1);
  double oC = sum_dnbinom(ox, nc, 1, r - 1, st.ocasei, T, dstart, dstart + r, 0.1, one, 1);
  one[1] = D->case_nbinom_sizey2;
  yC = yC + sum_dnbinom(yx, nc, r, nc, st.ycasei, T, dstart + r + 1, dend, 0.1, one, 1);
  one[1] = D->case_nbinom_sizeo2;
  oC = oC + sum_dnbinom(ox, nc, r, nc, st.ocasei, T, dstart + r + 1, dend, 0.1, one, 1);

  /* hosp, :528-554 */
  const int nh = D->n_dhospi, nhh = D->n_dhosp;
  window(offset, nh, T, &dstart, &dend);
  double *H = vec(T, 0);
  for (int t = 1; t <= T; ++t) H[t] = st.yhospi[t] + st.ohospi[t];
  const double *hx = D->dhospi - 1;
  one[1] = D->hosp_nbinom_size1;
  double lH = sum_dnbinom(hx, nh, 1, rh - 1, H, T, dstart, dstart + rh, 0.1, one, 1);
  one[1] = D->hosp_nbinom_size2;
  lH = lH + sum_dnbinom(hx, nh, rh, nhh, H, T, dstart + rh + 1, dend, 0.1, one, 1);
  one[1] = D->hosp_nbinom_size2 * D->hosp_last7;
  lH = lH + sum_dnbinom(hx, nh, nhh - 7, nhh, H, T, dend - 7, dend, 0.1, one, 1);
  free(H);

  /* :563-565 */
  long double yt = 0, ot = 0;
  for (int t = dstart + D->d_hosp_o1; t <= dstart + D->d_hosp_o2; ++t) {
    yt += at(st.yhospi, T, t);
    ot += at(st.ohospi, T, t);
  }
  const double ytotal = (double)yt, ototal = (double)ot;
  const double lHR = oracle_dlnorm_log(ytotal / (ytotal + ototal), log(56.7 / 100), log(1.05));

  /* deaths, :568-590 */
  const int nm = D->n_dmorti;
  window(offset, nm, T, &dstart, &dend);
  one[1] = D->mort_nbinom_size * 2;
  double yD = sum_dnbinom(D->y_dmorti - 1, nm, 1, nm, st.ydeadi, T, dstart, dend, 0.001, one, 1);
  double *sz = vec(nm, 0);
  for (int i = 1; i <= nm; ++i) sz[i] = D->o_wdmorti[i - 1] * D->mort_nbinom_size;
  double oD = sum_dnbinom(D->o_dmorti - 1, nm, 1, nm, st.odeadi, T, dstart, dend, 0.001, sz, nm);
  free(sz);

  terms[0] = yC; terms[1] = oC; terms[2] = lH; terms[3] = lHR; terms[4] = yD; terms[5] = oD;
  double result = yC + oC + lH + lHR + yD + oD; /* :592 */
  if (isnan(result)) *status |= EPI_ST_NA;
  age_state_free(&st);
  return result;
}

/* df_params, R/models/model-age-groups2q.R:624-663 */
static void age_df_params(const epi_data *D, double *mn, double *mx, double *init) {
  const double g = oracle_age_gamma();
  const double lo = D->lockdown_offset, ltp = D->lockdown_transition_period;
  const double i_[45] = {2.7, 1.7, 1.7, 2.3, 0.8, 0.8, 0.6, 0.4, 0.1, 2000, 1, 10, 15, 50, 2.5, 10, 21,
                         16, -10, 5, 5, 0.15, 0.5, D->d3 - lo - ltp, D->d4 - D->d3, 1, 0.2, 0.05,
                         0.7, 0.2, 0.01, D->d6 - D->d5, 1.1, 0.4, 0.03, D->d7 - D->d6, 1.4, 0.4, 0.08,
                         0.85, 0.85, 0.85, 0.4, 10, 10};
  const double mn_[45] = {2 * g, 2 * g, 2 * g, 1 * g, 1 * g, 1 * g, 0.05 * g, 0.01 * g, 0.01 * g,
                          3, 0.5, 6, 10, 3, 0.5, 6, 10, 0, -30, 2, 2, 0.05, 0.05, 60,
                          10, 0.05 * g, 0.01 * g, 0.002 * g, 0.05 * g, 0.01 * g, 0.002 * g,
                          10, 0.05 * g, 0.01 * g, 0.002 * g, 10, 0.05 * g, 0.01 * g, 0.002 * g,
                          0.6, 0.6, 0.6, 0.1, 4, 4};
  const double mx_[45] = {8 * g, 8 * g, 8 * g, 5 * g, 5 * g, 5 * g, 2 * g, 2 * g, 2 * g,
                          5000, 1.5, 15, 25, 100, 6, 15, 25,
                          fmax(D->dmort_last / 10, D->total_deaths_at_lockdown * 10),
                          30, 9, 9, 1.5, 1.5, 90,
                          50, 3 * g, 3 * g, 3 * g, 1.5 * g, 1.5 * g, 0.5 * g,
                          50, 3 * g, 3 * g, 0.5 * g, 50, 3 * g, 3 * g, 0.5 * g,
                          1.1, 1.1, 1.1, 0.7, 14, 14};
  for (int i = 0; i < 45; ++i) { mn[i] = mn_[i]; mx[i] = mx_[i]; init[i] = i_[i]; }
}

/* ------------------------------------------------------------ age-2i ------ */

/* three-slice rate map of R/models/model-age-groups2i.R:214-227 etc.: like rate_map(), but
 * the InvalidDataOffset branch multiplies by an explicit scalar (`:224`: ycase_rate *
 * y.died_rate, without the pmin of the valid branch), and a rate vector may carry a scalar
 * factor applied to s2i FIRST (`:236`: s2i * yhosp_rate * y.hr[...] is left-associative). */
static void rate_map2i(double *dst, const double *s2i, int pad, int P, int valid, double t1d, double pre,
                       const double *rate, int nr, double pre_invalid, double rate_invalid) {
  if (valid) {
    int t1 = (int)t1d;
    int t2 = (pad + P < t1 + nr - 1) ? pad + P : t1 + nr - 1;
    if (t1 > pad && t2 > t1) {
      for (int t = pad + 1; t <= t1; ++t) dst[t] = s2i[t - pad] * pre * rate[1];
      for (int t = t1; t <= t2; ++t) dst[t] = s2i[t - pad] * pre * rate[t - t1 + 1];
      for (int t = t2; t <= pad + P; ++t) dst[t] = s2i[t - pad] * pre * rate[nr];
    }
  } else {
    for (int t = pad + 1; t <= pad + P; ++t) dst[t] = s2i[t - pad] * pre_invalid * rate_invalid;
  }
}

/* calculateModel(params, period): R/models/model-age-groups2i.R:50-316 (params 1-based, 36) */
static int age2i_calculate_model(const epi_data *D, const double *params, int period, int m,
                                 age_state_t *st) {
  memset(st, 0, sizeof *st);
  const double y_N = D->y_N, o_N = D->o_N;
  const double y_died_rate = D->y_ifr[0], o_died_rate = D->o_ifr[0]; /* :11-12 */
  const double a1 = 1 / TINC1, a2 = 1 / TINC2, gamma = oracle_age_gamma();

  /* :52-93 */
  double betay[8], betao[8], betayo[8];
  betay[0] = params[1]; betao[0] = params[2]; betayo[0] = params[3];
  betay[1] = params[4]; betao[1] = params[5]; betayo[1] = params[6];
  betay[2] = params[7]; betao[2] = params[8]; betayo[2] = params[9];
  const double ycase_rate = params[10], ycase_latency = params[11], yhosp_rate = params[12],
               yhosp_latency = params[13], ydied_latency = params[14], ocase_rate = params[15],
               ocase_latency = params[16], ohosp_rate = params[17], ohosp_latency = params[18],
               odied_latency = params[19], t0_morts = params[20], t0o = params[21], CLsd = params[22],
               HLsd = params[23], DLsd = params[24], t3o = params[25], t4o = params[26], t6o = params[30];
  betay[3] = betay[2]; betao[3] = betao[2]; betayo[3] = betayo[2];
  betay[4] = params[27]; betao[4] = params[28]; betayo[4] = params[29];
  betay[5] = betay[4]; betao[5] = betao[4]; betayo[5] = betayo[4];
  betay[6] = params[31]; betao[6] = params[32]; betayo[6] = params[33];
  betay[7] = params[34] * betay[6]; betayo[7] = params[35] * betayo[6]; betao[7] = params[36] * betao[6];

  /* :96-101 */
  profile_t ycase, ocase, yhosp, ohosp, ydied, odied;
  if (calc_gamma_profile(ycase_latency, CLsd, &ycase) || calc_gamma_profile(ocase_latency, CLsd, &ocase) ||
      calc_gamma_profile(yhosp_latency, HLsd, &yhosp) || calc_gamma_profile(ohosp_latency, HLsd, &ohosp) ||
      calc_gamma_profile(ydied_latency, DLsd, &ydied) || calc_gamma_profile(odied_latency, DLsd, &odied))
    return -1;

  /* :103-104 */
  int padding = -ycase.kbegin;
  if (-ydied.kbegin > padding) padding = -ydied.kbegin;
  if (-ocase.kbegin > padding) padding = -ocase.kbegin;
  if (-odied.kbegin > padding) padding = -odied.kbegin;
  padding += 1;

  const int P = period, T = padding + period;
  st->pad = padding; st->T = T; st->P = P;
  st->yS = vec(T, y_N - INITIAL);   /* :108 */
  st->oS = vec(T, o_N);             /* :117 */
  st->ycasei = vec(T, 0); st->ocasei = vec(T, 0);
  st->yhospi = vec(T, 0); st->ohospi = vec(T, 0);
  st->ydeadi = vec(T, 0); st->odeadi = vec(T, 0);

  /* :133-143 */
  rhs_t rp;
  memset(&rp, 0, sizeof rp);
  rp.flavour = EPI_FLAVOUR_AGE2I;
  rp.Ny = y_N; rp.No = o_N; rp.a1 = a1; rp.a2 = a2; rp.gamma = gamma;
  rp.s.nbp = 8;
  static const int kind[8] = {1, 1, 1, 1, 1, 1, 0, 0}; /* model-agef.cpp:66-87 */
  for (int k = 0; k < 8; ++k) {
    rp.s.kind[k] = kind[k];
    rp.s.t[k] = 1E10;
    rp.by[k] = betay[k]; rp.bo[k] = betao[k]; rp.byo[k] = betayo[k];
  }
  const double Y[10] = {y_N - INITIAL, INITIAL, 0, 0, 0, o_N, 0, 0, 0, 0};

  /* pass 1 over the WHOLE period, :148-163 */
  double *out = (double *)malloc(sizeof(double) * 10 * (size_t)P);
  ode_rk4(&rp, 10, Y, padding + 1, P, m, out);
  for (int i = 1; i <= P; ++i) {
    st->yS[padding + i] = out[(i - 1) * 10 + 0];
    st->oS[padding + i] = out[(i - 1) * 10 + 5];
  }
  double *s2 = vec(P, 0), *s2b = vec(P, 0);
  convolute(st->yS, T, padding + 1, padding + P, &ydied, s2);
  convolute(st->oS, T, padding + 1, padding + P, &odied, s2b);
  double data_offset = EPI_INVALID_DATA_OFFSET; /* :165 */
  int lds1 = 0;
  for (int i = 1; i <= P; ++i) { /* which(state$died > t0_morts), :167 (NA on 1..padding) */
    double died = (y_N - s2[i]) * y_died_rate + (o_N - s2b[i]) * o_died_rate;
    if (died > t0_morts) { lds1 = padding + i; break; }
  }
  free(s2b);

  const int lo = D->lockdown_offset, ltp = D->lockdown_transition_period;
  if (lds1 > 0) { /* :168-197 */
    data_offset = lds1 - lo;
    double t2 = data_offset + lo + ltp;
    double t3 = t2 + t3o;
    double t4 = t3 + t4o;
    double t5 = data_offset + D->d5;
    double t6 = t5 + t6o;
    double t7 = data_offset + D->d7;
    rp.s.t[0] = data_offset + lo + t0o;
    rp.s.t[1] = data_offset + lo + fmax(t0o, 0);
    rp.s.t[2] = t2; rp.s.t[3] = t3; rp.s.t[4] = t4; rp.s.t[5] = t5; rp.s.t[6] = t6; rp.s.t[7] = t7;
    ode_rk4(&rp, 10, Y, padding + 1, P, m, out);
  }
  /* :199-210 */
  st->full = (double *)malloc(sizeof(double) * 10 * (size_t)P);
  memcpy(st->full, out, sizeof(double) * 10 * (size_t)P);
  st->yout = (double *)malloc(sizeof(double) * 4 * (size_t)P);
  yout_rows(&rp, 10, out, padding + 1, P, st->yout);
  for (int i = 1; i <= P; ++i) {
    st->yS[padding + i] = out[(i - 1) * 10 + 0];
    st->oS[padding + i] = out[(i - 1) * 10 + 5];
  }
  free(out);

  const int nr = D->n_rate;
  double *ycr = vec(nr, 0), *ocr = vec(nr, 0);
  for (int i = 1; i <= nr; ++i) {
    ycr[i] = pmin_(1, ycase_rate * D->y_ifr[i - 1]); /* :219-221 */
    ocr[i] = pmin_(1, ocase_rate * D->o_ifr[i - 1]); /* :267-269 */
  }
  const int valid = data_offset != EPI_INVALID_DATA_OFFSET;
  double *s2i = vec(P, 0);
  struct {
    const double *S; double Ng; const profile_t *prof; double t1; double pre; const double *rate;
    double pre_inv, rate_inv; double *dst;
  } ser[6] = {
      /* :212-226 */ {st->yS, y_N, &ycase, data_offset + nearbyint(-ydied_latency + ycase_latency), 1.0, ycr,
                      ycase_rate, y_died_rate, st->ycasei},
      /* :229-242 */ {st->yS, y_N, &yhosp, data_offset + nearbyint(-ydied_latency + yhosp_latency), yhosp_rate,
                      D->y_hr - 1, yhosp_rate, D->y_hr[0], st->yhospi},
      /* :245-257 */ {st->yS, y_N, &ydied, data_offset, 1.0, D->y_ifr - 1, 1.0, y_died_rate, st->ydeadi},
      /* :260-273 */ {st->oS, o_N, &ocase, data_offset + nearbyint(-ydied_latency + ocase_latency), 1.0, ocr,
                      ocase_rate, o_died_rate, st->ocasei},
      /* :276-289 */ {st->oS, o_N, &ohosp, data_offset + nearbyint(-ydied_latency + ohosp_latency), ohosp_rate,
                      D->o_hr - 1, ohosp_rate, D->o_hr[0], st->ohospi},
      /* :291-304 */ {st->oS, o_N, &odied, data_offset, 1.0, D->o_ifr - 1, 1.0, o_died_rate, st->odeadi},
  };
  for (int q = 0; q < 6; ++q) {
    convolute(ser[q].S, T, padding + 1, padding + P, ser[q].prof, s2);
    for (int i = 1; i <= P; ++i) s2[i] = ser[q].Ng - s2[i];
    s2i[1] = s2[1];
    for (int i = 2; i <= P; ++i) s2i[i] = s2[i] - s2[i - 1];
    rate_map2i(ser[q].dst, s2i, padding, P, valid, ser[q].t1, ser[q].pre, ser[q].rate, nr, ser[q].pre_inv,
               ser[q].rate_inv);
  }
  free(s2); free(s2i); free(ycr); free(ocr);
  st->offset = (int)data_offset;
  return 0;
}

/* calclogp(params): R/models/model-age-groups2i.R:362-438 (params 1-based) */
static double age2i_calclogp(const epi_data *D, const double *params) {
  const double gamma = oracle_age_gamma();
  const double betay0 = params[1], betao0 = params[2], betayo0 = params[3], betay1 = params[4],
               betao1 = params[5], betayo1 = params[6], betay2 = params[7], betao2 = params[8],
               betayo2 = params[9], ydied_latency = params[14], odied_latency = params[19],
               t0o = params[21], CLsd = params[22], HLsd = params[23], DLsd = params[24],
               t3o = params[25], t4o = params[26], betay4 = params[27], betao4 = params[28],
               betayo4 = params[29], t6o = params[30], betay6 = params[31], betao6 = params[32],
               betayo6 = params[33];
  const double betay3 = betay2, betao3 = betao2, betayo3 = betayo2;
  const double betay7 = params[34] * betay6, betayo7 = params[35] * betayo6, betao7 = params[36] * betao6;
  const double lo = D->lockdown_offset, ltp = D->lockdown_transition_period;
  double lp = 0;
  lp = lp + oracle_dnorm_log(t0o, 0, 10);
  lp = lp + oracle_dnorm_log(lo + ltp + t3o, D->d3, 10);
  lp = lp + oracle_dnorm_log(lo + ltp + t3o + t4o, D->d4, 10);
  lp = lp + oracle_dnorm_log(t6o, G_CONST, 3);
  lp = lp + oracle_dnorm_log(betay0 - betay1, 0, 2 * gamma);
  lp = lp + oracle_dnorm_log(betao0 - betao1, 0, 2 * gamma);
  lp = lp + oracle_dnorm_log(betayo0 - betayo1, 0, 2 * gamma);
  lp = lp + oracle_dnorm_log(betay1 - betay2, 0, 2 * gamma);
  lp = lp + oracle_dnorm_log(betao1 - betao2, 0, 2 * gamma);
  lp = lp + oracle_dnorm_log(betayo1 - betayo2, 0, 2 * gamma);
  lp = lp + oracle_dnorm_log(betay3 - betay4, 0, 0.5 * gamma);
  lp = lp + oracle_dnorm_log(betao3 - betao4, 0, 0.5 * gamma);
  lp = lp + oracle_dnorm_log(betayo3 - betayo4, 0, 0.5 * gamma);
  lp = lp + oracle_dnorm_log(CLsd, 5, 1);
  lp = lp + oracle_dnorm_log(HLsd, 5, 1);
  lp = lp + oracle_dnorm_log(DLsd, 5, 1);
  lp = lp + oracle_dnorm_log(ydied_latency, 21, 4);
  lp = lp + oracle_dnorm_log(odied_latency, 21, 4);
  lp = lp + oracle_dnorm_log(betay6, 0.6, 0.2);
  lp = lp + oracle_dnorm_log(betao6, 0.22, 0.2);
  lp = lp + oracle_dnorm_log(betayo6, 0.26, 0.2);
  lp = lp + oracle_dnorm_log(betay7, 0.85, 0.2);
  lp = lp + oracle_dnorm_log(betao7, 0.19, 0.05);
  lp = lp + oracle_dnorm_log(betayo7, 0.02, 0.02);
  return lp;
}

/* calclogl(params, x): R/models/model-age-groups2i.R:442-536.  terms[0..5] =
 * y.loglC, o.loglC, loglH, 0 (no ratio term in this generation), y.loglD, o.loglD */
static double age2i_calclogl(const epi_data *D, const double *params, int period, int m,
                             double *terms, int *offset_out, int *pad_out, int *status) {
  age_state_t st;
  if (age2i_calculate_model(D, params, period, m, &st)) {
    *status |= EPI_ST_UNSUPPORTED;
    for (int i = 0; i < 6; ++i) terms[i] = NA_REAL;
    *offset_out = 0; *pad_out = 0;
    return NA_REAL;
  }
  *offset_out = st.offset;
  *pad_out = st.pad;
  int offset = st.offset;
  if (offset == EPI_INVALID_DATA_OFFSET) { /* :445-446 */
    offset = 1;
    *status |= EPI_ST_INVALID_OFFSET;
  }
  const int T = st.T;
  int dstart, dend;
  const int r = D->d_reliable_cases, nc = D->n_dcasei;
  double one[2];

  /* cases, :449-478 */
  window(offset, nc, T, &dstart, &dend);
  const double *yx = D->y_dcasei - 1, *ox = D->o_dcasei - 1;
  one[1] = D->case_nbinom_size1;
  double yC = sum_dnbinom(yx, nc, 1, r - 1, st.ycasei, T, dstart, dstart + r, 0.1, one, 1);
  double oC = sum_dnbinom(ox, nc, 1, r - 1, st.ocasei, T, dstart, dstart + r, 0.1, one, 1);
  one[1] = D->case_nbinom_sizey2; /* case_nbinom_size2 */
  yC = yC + sum_dnbinom(yx, nc, r, nc, st.ycasei, T, dstart + r + 1, dend, 0.1, one, 1);
  one[1] = D->case_nbinom_sizeo2; /* case_nbinom_size2 */
  oC = oC + sum_dnbinom(ox, nc, r, nc, st.ocasei, T, dstart + r + 1, dend, 0.1, one, 1);

  /* hosp, :481-498 */
  const int nh = D->n_dhospi;
  window(offset, nh, T, &dstart, &dend);
  double *H = vec(T, 0);
  for (int t = 1; t <= T; ++t) H[t] = st.yhospi[t] + st.ohospi[t];
  one[1] = D->hosp_nbinom_size;
  double lH = sum_dnbinom(D->dhospi - 1, nh, 1, nh, H, T, dstart, dend, 0.001, one, 1);
  free(H);

  /* deaths, :501-522 */
  const int nm = D->n_dmorti;
  window(offset, nm, T, &dstart, &dend);
  one[1] = D->mort_nbinom_size;
  double yD = sum_dnbinom(D->y_dmorti - 1, nm, 1, nm, st.ydeadi, T, dstart, dend, 0.001, one, 1);
  double oD = sum_dnbinom(D->o_dmorti - 1, nm, 1, nm, st.odeadi, T, dstart, dend, 0.001, one, 1);

  terms[0] = yC; terms[1] = oC; terms[2] = lH; terms[3] = 0; terms[4] = yD; terms[5] = oD;
  double result = yC + oC + lH + yD + oD; /* :526 */
  if (isnan(result)) *status |= EPI_ST_NA;
  age_state_free(&st);
  return result;
}

/* df_params, R/models/model-age-groups2i.R:552-587 */
static void age2i_df_params(const epi_data *D, double *mn, double *mx, double *init) {
  const double g = oracle_age_gamma();
  const double lo = D->lockdown_offset, ltp = D->lockdown_transition_period;
  const double i_[36] = {3.6 * g, 3.6 * g, 3.6 * g, 2.0 * g, 2.0 * g, 2.0 * g, 0.8 * g, 0.8 * g, 0.8 * g,
                         100, 10, 1, 10, 21, 20, 10, 1, 10, 21,
                         D->total_deaths_at_lockdown, -1, 5, 5, 5,
                         D->d3 - lo - ltp, D->d4 - D->d3, 0.8 * g, 0.8 * g, 0.8 * g,
                         G_CONST, 0.8 * g, 0.8 * g, 0.8 * g, 1.2, 1.01, 1.01};
  const double mn_[36] = {2 * g, 2 * g, 2 * g, 1 * g, 1 * g, 1 * g, 0.05 * g, 0.01 * g, 0.01 * g,
                          3, 3, 0.5, 3, 10, 3, 3, 0.5, 3, 10,
                          0, -30, 2, 2, 2, 60,
                          10, 0.05 * g, 0.01 * g, 0.01 * g,
                          3, 0.05 * g, 0.01 * g, 0.01 * g, 1, 1, 1};
  const double mx_[36] = {8 * g, 8 * g, 8 * g, 5 * g, 5 * g, 5 * g, 2 * g, 2 * g, 2 * g,
                          5000, 25, 2, 25, 50, 100, 25, 2, 25, 50,
                          fmax(D->dmort_last / 10, D->total_deaths_at_lockdown * 10),
                          30, 9, 9, 9, 90,
                          50, 3 * g, 3 * g, 3 * g,
                          12, 1.5 * g, 1.5 * g, 1.5 * g, 2, 6, 6};
  for (int i = 0; i < 36; ++i) { mn[i] = mn_[i]; mx[i] = mx_[i]; init[i] = i_[i]; }
}

/* ------------------------------------------------------------ age-2x ------ */
/* R/models/model-age-groups2x.R + model-ager.cpp: the later generation of the two-group model (params 1-based, 52).
 * State series: yS/oS keep S (for calculateModel's fields), yI/oI hold the infection incidence y.i / o.i that
 * derivs() reports as output variables 5 and 6 — the series every observation model convolves here. */
typedef struct {
  age_state_t a;
  double *yi, *oi; /* length T, 0 before padding + 1 (:156-157) */
  int unsupported;  /* the symptomatic-testing window would start beyond the simulated period (vector growth in R) */
} age2x_state_t;

static void age2x_state_free(age2x_state_t *s) {
  free(s->yi); free(s->oi);
  age_state_free(&s->a);
  memset(s, 0, sizeof *s);
}

static void age2x_betas(const double *params, double *by, double *bo, double *byo) {
  /* :56-119 */
  by[0] = params[1]; bo[0] = params[2]; byo[0] = params[3];
  by[1] = params[4]; bo[1] = params[5]; byo[1] = params[6];
  by[2] = params[7]; bo[2] = params[8]; byo[2] = params[9];
  by[3] = params[25]; bo[3] = params[26]; byo[3] = params[27];
  by[4] = params[29]; bo[4] = params[30]; byo[4] = params[31];
  by[5] = params[32]; bo[5] = params[33]; byo[5] = params[34];
  by[6] = params[36]; bo[6] = params[37]; byo[6] = params[38];
  by[7] = params[40]; bo[7] = params[41]; byo[7] = params[42];
  by[8] = by[7]; bo[8] = bo[7]; byo[8] = byo[7];
  by[9] = params[43]; bo[9] = params[44]; byo[9] = params[45];
  by[10] = by[9]; bo[10] = bo[9]; byo[10] = byo[9];
  by[11] = params[50]; bo[11] = params[51]; byo[11] = params[52];
  by[12] = by[9]; bo[12] = bo[9]; byo[12] = byo[9];
}

/* R's cumsum: LDOUBLE accumulator, NA propagates from the first NA on */
static void cumsum_r(const double *x, int n, double *out /*1-based*/) {
  long double acc = 0;
  for (int i = 1; i <= n; ++i) {
    acc += x[i];
    out[i] = (double)acc;
  }
}

static int age2x_calculate_model(const epi_data *D, const double *params, int period, int m, age2x_state_t *sx) {
  memset(sx, 0, sizeof *sx);
  age_state_t *st = &sx->a;
  const double y_N = D->y_N, o_N = D->o_N;
  const double y_died_rate = D->y_ifr[0], o_died_rate = D->o_ifr[0]; /* :14-15 */
  const double a1 = 1 / TINC1, a2 = 1 / TINC2, gamma = oracle_age_gamma();
  const double eta = 1 / (0.75 * 356); /* :19 (sic: 356) */
  const double CLsd = 2.5;             /* :20 */

  double betay[14], betao[14], betayo[14];
  age2x_betas(params, betay, betao, betayo);
  const double ycase_rate = params[10], yhosp_rate = params[11], yhosp_latency = params[12],
               ydied_latency = params[13], ocase_rate = params[14], ohosp_rate = params[15],
               ohosp_latency = params[16], odied_latency = params[17], t0_morts = params[18],
               t0o = params[19], HLsd = params[20], DLsd = params[21], fyifr = params[22],
               fyhr = params[23], t3o = params[24], t4o = params[28], t6o = params[35], t7o = params[39],
               ifrred = params[46], ycase_latency = params[47], ocase_latency = params[48], t10o = params[49];

  /* :123-128 */
  profile_t ycase, ocase, yhosp, ohosp, ydied, odied;
  if (calc_gamma_profile(ycase_latency, CLsd, &ycase) || calc_gamma_profile(ocase_latency, CLsd, &ocase) ||
      calc_gamma_profile(yhosp_latency, HLsd, &yhosp) || calc_gamma_profile(ohosp_latency, HLsd, &ohosp) ||
      calc_gamma_profile(ydied_latency, DLsd, &ydied) || calc_gamma_profile(odied_latency, DLsd, &odied))
    return -1;
  /* :130-131 */
  int padding = -ycase.kbegin;
  if (-ydied.kbegin > padding) padding = -ydied.kbegin;
  if (-ocase.kbegin > padding) padding = -ocase.kbegin;
  if (-odied.kbegin > padding) padding = -odied.kbegin;
  padding += 1;

  const int P = period, T = padding + period;
  st->pad = padding; st->T = T; st->P = P;
  st->yS = vec(T, y_N - INITIAL); st->oS = vec(T, o_N);
  st->ycasei = vec(T, 0); st->ocasei = vec(T, 0);
  st->yhospi = vec(T, 0); st->ohospi = vec(T, 0);
  st->ydeadi = vec(T, 0); st->odeadi = vec(T, 0);
  sx->yi = vec(T, 0); sx->oi = vec(T, 0); /* :156-157 */

  /* :162-181 */
  rhs_t rp;
  memset(&rp, 0, sizeof rp);
  rp.flavour = EPI_FLAVOUR_AGE2X;
  rp.Ny = y_N; rp.No = o_N; rp.a1 = a1; rp.a2 = a2; rp.gamma = gamma; rp.eta = eta;
  rp.tuncertain = 1E10; rp.funcertain = 1;
  rp.s.nbp = 13;
  static const int kind[13] = {1, 1, 1, 1, 0, 0, 0, 0, 1, 0, 1, 1, 0}; /* model-ager.cpp:88-124 */
  for (int k = 0; k < 13; ++k) {
    rp.s.kind[k] = kind[k];
    rp.s.t[k] = 1E10;
    rp.by[k] = betay[k]; rp.bo[k] = betao[k]; rp.byo[k] = betayo[k];
  }
  const double Y[10] = {y_N - INITIAL, INITIAL, 0, 0, 0, o_N, 0, 0, 0, 0};

  /* pass 1, :183-203 */
  const int period1 = 60;
  const int rows_max = P > period1 ? P : period1;
  double *out = (double *)malloc(sizeof(double) * 10 * (size_t)rows_max);
  double *yo9 = (double *)malloc(sizeof(double) * 9 * (size_t)rows_max);
  int out_rows = period1;
  ode_rk4(&rp, 10, Y, padding + 1, period1, m, out);
  for (int i = 0; i < period1; ++i) derivs_ager(&rp, SEG_LITERAL, (double)(padding + 1 + i), out + (size_t)i * 10, 0, yo9 + (size_t)i * 9);
  const int T1 = (T > padding + period1) ? T : padding + period1; /* assignment may grow the vector */
  double *yi1 = vec(T1, 0), *oi1 = vec(T1, 0);
  for (int i = 1; i <= period1; ++i) {
    yi1[padding + i] = yo9[(size_t)(i - 1) * 9 + 4];
    oi1[padding + i] = yo9[(size_t)(i - 1) * 9 + 5];
  }
  double *s2 = vec(rows_max, 0), *s2b = vec(period1, 0), *cy = vec(period1, 0), *co = vec(period1, 0);
  convolute(yi1, T1, padding + 1, padding + period1, &ydied, s2);
  convolute(oi1, T1, padding + 1, padding + period1, &odied, s2b);
  for (int i = 1; i <= period1; ++i) { s2[i] = s2[i] * y_died_rate; s2b[i] = s2b[i] * o_died_rate; }
  cumsum_r(s2, period1, cy);  /* :194-195 */
  cumsum_r(s2b, period1, co); /* :197-198 */
  double data_offset = EPI_INVALID_DATA_OFFSET; /* :202 */
  int lds1 = 0;
  for (int i = 1; i <= period1; ++i) { /* which(state$died > t0_morts), :204 (NA before padding + 1) */
    const double died = cy[i] + co[i];
    if (died > t0_morts) { lds1 = padding + i; break; }
  }
  free(yi1); free(oi1); free(s2b); free(cy); free(co);

  const int lo = D->lockdown_offset, ltp = D->lockdown_transition_period;
  if (lds1 > 0) { /* :205-250 */
    data_offset = lds1 - lo;
    const double t2 = data_offset + lo + ltp;
    const double t3 = t2 + t3o;
    const double t4 = t3 + t4o;
    const double t5 = data_offset + D->d5;
    const double t6 = t5 + t6o;
    const double t7 = t6 + t7o;
    const double t8 = data_offset + D->d8;
    const double t9 = data_offset + D->d9;
    const double t10 = data_offset + t10o;
    const double t11 = t10 + 3;
    const double t12 = data_offset + D->d12;
    rp.tuncertain = data_offset + D->duncertain; /* :220 */
    rp.funcertain = 1;                           /* :221 */
    rp.s.t[0] = data_offset + lo + t0o;
    rp.s.t[1] = data_offset + lo + fmax(t0o, 0);
    rp.s.t[2] = t2; rp.s.t[3] = t3; rp.s.t[4] = t4; rp.s.t[5] = t5; rp.s.t[6] = t6; rp.s.t[7] = t7;
    rp.s.t[8] = t8; rp.s.t[9] = t9; rp.s.t[10] = t10; rp.s.t[11] = t11; rp.s.t[12] = t12;
    ode_rk4(&rp, 10, Y, padding + 1, P, m, out);
    out_rows = P;
    for (int i = 0; i < P; ++i) derivs_ager(&rp, SEG_LITERAL, (double)(padding + 1 + i), out + (size_t)i * 10, 0, yo9 + (size_t)i * 9);
  }
  /* :253-268 — when pass 2 was skipped the 60-row `out` is recycled into the P-length slices */
  st->full = (double *)malloc(sizeof(double) * 10 * (size_t)P);
  st->yout = (double *)malloc(sizeof(double) * 4 * (size_t)P);
  for (int i = 1; i <= P; ++i) {
    const size_t r = (size_t)((i - 1) % out_rows);
    const double *row = out + r * 10;
    st->yS[padding + i] = row[0];
    st->oS[padding + i] = row[5];
    sx->yi[padding + i] = yo9[r * 9 + 4];
    sx->oi[padding + i] = yo9[r * 9 + 5];
    memcpy(st->full + (size_t)(i - 1) * 10, row, sizeof(double) * 10);
    memcpy(st->yout + (size_t)(i - 1) * 4, yo9 + r * 9, sizeof(double) * 4);
  }
  free(out); free(yo9);

  /* :270-276 */
  const int nr = D->n_rate;
  double *gycr = vec(nr, 0), *gyifr = vec(nr, 0), *gyhr = vec(nr, 0), *gocr = vec(nr, 0),
         *goifr = vec(nr, 0), *gohr = vec(nr, 0);
  for (int i = 1; i <= nr; ++i) {
    const double yifr = D->y_ifr[i - 1], oifr = D->o_ifr[i - 1], yhr = D->y_hr[i - 1],
                 ohr = D->o_hr[i - 1], g = D->g[i - 1], gt = D->gtrimp[i - 1];
    gycr[i] = pmin_(1, ycase_rate * yifr * (1 + (fyifr - 1) * g));
    gyifr[i] = yifr * (1 + (fyifr - 1) * g) * (1 - ifrred * gt);
    gyhr[i] = yhosp_rate * yhr * (1 + (fyhr - 1) * g);
    gocr[i] = pmin_(1, ocase_rate * oifr);
    goifr[i] = oifr * (1 - ifrred * gt);
    gohr[i] = ohosp_rate * ohr;
  }

  const int valid = data_offset != EPI_INVALID_DATA_OFFSET;
  struct {
    const double *I; const profile_t *prof; double t1; const double *rate; double *dst; int cases;
  } ser[6] = {
      /* :281-297 */ {sx->yi, &ycase, data_offset + nearbyint(-ydied_latency + ycase_latency), gycr, st->ycasei, 1},
      /* :300-312 */ {sx->yi, &yhosp, data_offset + nearbyint(-ydied_latency + yhosp_latency), gyhr, st->yhospi, 0},
      /* :314-326 */ {sx->yi, &ydied, data_offset, gyifr, st->ydeadi, 0},
      /* :329-343 */ {sx->oi, &ocase, data_offset + nearbyint(-ydied_latency + ocase_latency), gocr, st->ocasei, 1},
      /* :346-357 */ {sx->oi, &ohosp, data_offset + nearbyint(-ydied_latency + ohosp_latency), gohr, st->ohospi, 0},
      /* :359-371 */ {sx->oi, &odied, data_offset, goifr, st->odeadi, 0},
  };
  /* :278-279 */
  const int t_symp = (int)data_offset + D->d_symp_cases, t_all = (int)data_offset + D->d_all_cases - 1;
  for (int q = 0; q < 6; ++q) {
    convolute(ser[q].I, T, padding + 1, padding + P, ser[q].prof, s2);
    rate_map(ser[q].dst, s2, padding, P, valid, ser[q].t1, ser[q].rate, nr);
    if (valid && ser[q].cases) {
      /* state$y.casei[t.symp:min(t.all, length)] *= symp.cases.factor (:293-294,340-341); R's a:b counts down when
       * a > b.  A window that starts beyond the vector would GROW it in R (and move calclogl's windows): flagged
       * unsupported — the symptomatic-testing window then begins after the simulated period */
      int a = t_symp, b = t_all < T ? t_all : T;
      if (a > T || a < 1 || b < 1) { sx->unsupported = 1; continue; }
      const int lo_ = a < b ? a : b, hi_ = a < b ? b : a;
      for (int t = lo_; t <= hi_; ++t) ser[q].dst[t] = ser[q].dst[t] * D->symp_cases_factor;
    }
  }
  free(s2);
  free(gycr); free(gyifr); free(gyhr); free(gocr); free(goifr); free(gohr);
  st->offset = (int)data_offset;
  return 0;
}

/* calclogp(params): R/models/model-age-groups2x.R:427-545 (params 1-based) */
static double age2x_calclogp(const epi_data *D, const double *params) {
  double by[14], bo[14], byo[14];
  age2x_betas(params, by, bo, byo);
  const double yhosp_rate = params[11], yhosp_latency = params[12], ydied_latency = params[13],
               ohosp_rate = params[15], ohosp_latency = params[16], odied_latency = params[17],
               t0o = params[19], HLsd = params[20], DLsd = params[21], fyifr = params[22],
               t3o = params[24], t4o = params[28], t6o = params[35], t7o = params[39], ifrred = params[46],
               ycase_latency = params[47], ocase_latency = params[48], t10o = params[49];
  const double lo = D->lockdown_offset, ltp = D->lockdown_transition_period;
  double lp = 0;
  lp = lp + oracle_dnorm_log(t0o, 0, 10);
  lp = lp + oracle_dnorm_log(lo + ltp + t3o, D->d3, 10);
  lp = lp + oracle_dnorm_log(lo + ltp + t3o + t4o, D->d4, 10);
  lp = lp + oracle_dnorm_log(D->d5 + t6o, D->d6, 5);
  lp = lp + oracle_dnorm_log(D->d5 + t6o + t7o, D->d7 - lo, 2);
  lp = lp + oracle_dnorm_log(t10o, D->d10, 3);
  const double SD = log(2), lSD = log(1.3), llSD = log(1.2);
  lp = lp + oracle_dlnorm_log(by[0] / by[1], 0, llSD);
  lp = lp + oracle_dlnorm_log(bo[0] / bo[1], 0, llSD);
  lp = lp + oracle_dlnorm_log(byo[0] / byo[1], 0, llSD);
  static const int pairs[5][2] = {{1, 2}, {2, 3}, {3, 4}, {5, 6}, {6, 7}};
  for (int q = 0; q < 5; ++q) {
    lp = lp + oracle_dlnorm_log(by[pairs[q][0]] / by[pairs[q][1]], 0, SD);
    lp = lp + oracle_dlnorm_log(bo[pairs[q][0]] / bo[pairs[q][1]], 0, SD);
    lp = lp + oracle_dlnorm_log(byo[pairs[q][0]] / byo[pairs[q][1]], 0, SD);
  }
  lp = lp + oracle_dlnorm_log(by[7] / by[9], log(1.3), lSD);
  lp = lp + oracle_dlnorm_log(bo[7] / bo[9], log(1.3), lSD);
  lp = lp + oracle_dlnorm_log(byo[7] / byo[9], log(1.3), lSD);
  lp = lp + oracle_dlnorm_log(by[9] / by[11], 0, lSD);
  lp = lp + oracle_dlnorm_log(bo[9] / bo[11], 0, lSD);
  lp = lp + oracle_dlnorm_log(byo[9] / byo[11], 0, lSD);
  lp = lp + oracle_dnorm_log(HLsd, 4, 1);
  lp = lp + oracle_dnorm_log(DLsd, 5, 1);
  lp = lp + oracle_dnorm_log(ycase_latency, 7, 3);
  lp = lp + oracle_dnorm_log(ocase_latency, 7, 3);
  lp = lp + oracle_dnorm_log(yhosp_latency, 13, 1);
  lp = lp + oracle_dnorm_log(ohosp_latency, 13, 1);
  lp = lp + oracle_dnorm_log(ydied_latency, 21, 0.5);
  lp = lp + oracle_dnorm_log(odied_latency, 21, 0.5);
  lp = lp + oracle_dlnorm_log(fyifr, log(0.6), log(1.5));
  lp = lp + oracle_dlnorm_log(ifrred, log(0.35), log(1.3));
  lp = lp + oracle_dlnorm_log(yhosp_rate, 0, lSD);
  lp = lp + oracle_dlnorm_log(ohosp_rate, 0, lSD);
  return lp;
}

/* calclogl(params, x): R/models/model-age-groups2x.R:547-681.  terms[0..5] = y.loglC, o.loglC, loglH, loglHRatio,
 * y.loglD, o.loglD */
static double age2x_calclogl(const epi_data *D, const double *params, int period, int m,
                             double *terms, int *offset_out, int *pad_out, int *status) {
  age2x_state_t sx;
  if (age2x_calculate_model(D, params, period, m, &sx)) {
    *status |= EPI_ST_UNSUPPORTED;
    for (int i = 0; i < 6; ++i) terms[i] = NA_REAL;
    *offset_out = 0; *pad_out = 0;
    return NA_REAL;
  }
  const age_state_t *st = &sx.a;
  *offset_out = st->offset;
  *pad_out = st->pad;
  if (sx.unsupported) {
    *status |= EPI_ST_UNSUPPORTED;
    for (int i = 0; i < 6; ++i) terms[i] = NA_REAL;
    age2x_state_free(&sx);
    return NA_REAL;
  }
  int offset = st->offset;
  if (offset == EPI_INVALID_DATA_OFFSET) { /* :552-553 */
    offset = 1;
    *status |= EPI_ST_INVALID_OFFSET;
  }
  const int T = st->T;
  int dstart, dend;
  const int r = D->d_reliable_cases, rh = D->d_reliable_hosp, nc = D->n_dcasei;
  double one[2];

  /* cases, :555-590 */
  window(offset, nc, T, &dstart, &dend);
  const double *yx = D->y_dcasei - 1, *ox = D->o_dcasei - 1;
  one[1] = D->case_nbinom_size1;
  double yC = sum_dnbinom(yx, nc, 1, r - 1, st->ycasei, T, dstart, dstart + r, 0.1, one, 1);
  double oC = sum_dnbinom(ox, nc, 1, r - 1, st->ocasei, T, dstart, dstart + r, 0.1, one, 1);
  one[1] = D->case_nbinom_sizey2;
  yC = yC + sum_dnbinom(yx, nc, r, nc, st->ycasei, T, dstart + r + 1, dend, 0.1, one, 1);
  one[1] = D->case_nbinom_sizeo2;
  oC = oC + sum_dnbinom(ox, nc, r, nc, st->ocasei, T, dstart + r + 1, dend, 0.1, one, 1);
  one[1] = D->case_nbinom_sizeo2 * D->hosp_last7;
  oC = oC + sum_dnbinom(ox, nc, nc - 7, nc, st->ocasei, T, dend - 7, dend, 0.1, one, 1); /* :586-588 */

  /* hosp, :592-619 */
  const int nh = D->n_dhospi, nhh = D->n_dhosp;
  window(offset, nh, T, &dstart, &dend);
  double *H = vec(T, 0);
  for (int t = 1; t <= T; ++t) H[t] = st->yhospi[t] + st->ohospi[t];
  const double *hx = D->dhospi - 1;
  one[1] = D->hosp_nbinom_size1;
  double lH = sum_dnbinom(hx, nh, 1, rh - 1, H, T, dstart, dstart + rh, 0.1, one, 1);
  one[1] = D->hosp_nbinom_size2;
  lH = lH + sum_dnbinom(hx, nh, rh, nhh, H, T, dstart + rh + 1, dend, 0.1, one, 1);
  one[1] = D->hosp_nbinom_size2 * D->hosp_last7;
  lH = lH + sum_dnbinom(hx, nh, nhh - 7, nhh, H, T, dend - 7, dend, 0.1, one, 1);
  free(H);

  /* :627-630 */
  long double yt = 0, ot = 0;
  for (int t = dstart + D->d_hosp_o1; t <= dstart + D->d_hosp_o2; ++t) {
    yt += at(st->yhospi, T, t);
    ot += at(st->ohospi, T, t);
  }
  const double ytotal = (double)yt, ototal = (double)ot;
  const double lHR = oracle_dlnorm_log(ytotal / (ytotal + ototal), log(56.7 / 100), log(1.05));

  /* deaths, :633-657 */
  const int nm = D->n_dmorti;
  window(offset, nm, T, &dstart, &dend);
  one[1] = D->mort_nbinom_size * 5;
  double yD = sum_dnbinom(D->y_dmorti - 1, nm, 1, nm, st->ydeadi, T, dstart, dend, 0.001, one, 1);
  double *sz = vec(nm, 0);
  for (int i = 1; i <= nm; ++i) sz[i] = D->o_wdmorti[i - 1] * D->mort_nbinom_size;
  double oD = sum_dnbinom(D->o_dmorti - 1, nm, 1, nm, st->odeadi, T, dstart, dend, 0.001, sz, nm);
  free(sz);
  one[1] = D->mort_nbinom_size * D->hosp_last7;
  oD = oD + sum_dnbinom(D->o_dmorti - 1, nm, nm - 7, nm, st->odeadi, T, dend - 7, dend, 0.001, one, 1); /* :653-655 */

  terms[0] = yC; terms[1] = oC; terms[2] = lH; terms[3] = lHR; terms[4] = yD; terms[5] = oD;
  double result = yC + oC + lH + lHR + yD + oD; /* :658 */
  if (isnan(result)) *status |= EPI_ST_NA;
  age2x_state_free(&sx);
  return result;
}

/* df_params, R/models/model-age-groups2x.R:683-751 */
static void age2x_df_params(const epi_data *D, double *mn, double *mx, double *init) {
  const double g = oracle_age_gamma();
  const double i_[52] = {3.1, 0.52, 0.50, 1.5, 0.50, 0.44, 0.8, 0.35, 0.09, 550, 0.75, 14, 21, 11, 1.5, 14, 21,
                         14, -8, 5.2, 5.8, 0.49, 0.30, 92, 1.0, 0.18, 0.02, 29, 1.4, 0.4, 0.04, 0.96, 0.27, 0.007,
                         28, 1.2, 0.9, 0.02, 32, 2.0, 0.6, 0.10, 1.3, 0.3, 0.05, 0.34, 6, 11, 245, 1.9, 0.3, 0.04};
  const double mn_[52] = {2 * g, 0.5 * g, 0.5 * g, 1 * g, 0.5 * g, 0.5 * g, 0.1 * g, 0.1 * g, 0.01 * g,
                          500, 0.5, 9, 14, 3, 0.5, 9, 14, 0, -20, 2, 2, 0.05, 0.05,
                          60, 0.5 * g, 0.01 * g, 0.002 * g, 10, 0.5 * g, 0.01 * g, 0.002 * g,
                          0.1 * g, 0.01 * g, 0.002 * g, 10, 0.2 * g, 0.01 * g, 0.002 * g,
                          10, 0.2 * g, 0.01 * g, 0.002 * g, 0.2 * g, 0.01 * g, 0.002 * g, 0.1, 4, 4,
                          (double)D->d10 - 10, 0.2 * g, 0.01 * g, 0.002 * g};
  const double tdl = D->total_deaths_at_lockdown * 10, dml = D->dmort_last / 10;
  const double mx_[52] = {8 * g, 5 * g, 5 * g, 5 * g, 3 * g, 3 * g, 2 * g, 2 * g, 2 * g,
                          5000, 2, 18, 25, 100, 6, 18, 25, dml > tdl ? dml : tdl, 10, 9, 9, 1, 1,
                          100, 3 * g, 1 * g, 0.5 * g, 50, 3 * g, 1 * g, 0.5 * g,
                          2 * g, 1 * g, 0.5 * g, 50, 4 * g, 3 * g, 0.5 * g,
                          50, 4 * g, 3 * g, 0.5 * g, 3 * g, 3 * g, 0.5 * g, 0.7, 14, 17,
                          (double)D->duncertain - 2, 3 * g, 3 * g, 0.5 * g};
  for (int i = 0; i < 52; ++i) { mn[i] = mn_[i]; mx[i] = mx_[i]; init[i] = i_[i]; }
}

/* ------------------------------------------------------- simple / Ne ------ */

typedef struct {
  int pad, T, P, offset;
  double *S, *hospi, *deadi;
  double *full; /* [P][4] */
  double *yout; /* [P][2] Re, Rt */
} simple_state_t;

static void simple_state_free(simple_state_t *s) {
  free(s->S); free(s->hospi); free(s->deadi); free(s->full); free(s->yout);
  memset(s, 0, sizeof *s);
}

/* calculateModel(params, period): R/models/model-simple-ode-drj-cmp.R:41-129
 * (params 1-based, MODEL space: params[5] is the hospitalisation rate). */
static int simple_calculate_model(const epi_data *D, int flavour, const double *params, int period,
                                  int m, simple_state_t *st) {
  memset(st, 0, sizeof *st);
  const double N = D->N, Tinc = 3, died_rate = 0.007; /* :8-9 */
  const double R0 = params[1], Rt = params[2], Tinf0 = params[3], Tinft = params[4],
               hosp_rate = params[5], hosp_latency = params[6], died_latency = params[7],
               mort_lockdown_threshold = params[8];
  const double beta0 = R0 / Tinf0, betat = Rt / Tinft, a = 1 / Tinc;
  profile_t hosp, died;
  if (calc_convolve_profile(-hosp_latency, 5, &hosp) || calc_convolve_profile(-died_latency, 5, &died))
    return -1;
  int padding = (-hosp.kbegin > -died.kbegin ? -hosp.kbegin : -died.kbegin) + 1; /* :59 */
  const int P = period, T = padding + P;
  st->pad = padding; st->T = T; st->P = P;
  st->S = vec(T, N - INITIAL);
  double *hospv = vec(T, 0), *diedv = vec(T, 0);

  rhs_t rp;
  memset(&rp, 0, sizeof rp);
  rp.flavour = flavour;
  rp.N = N; rp.a = a; rp.ldts = 1E10; rp.ldte = 1E10;
  rp.beta0 = beta0; rp.Tinf0 = Tinf0; rp.betat = betat; rp.Tinft = Tinft;
  rp.Ne = (flavour == EPI_FLAVOUR_NE) ? params[9] * N : N;
  const double Y[4] = {N - INITIAL, INITIAL, 0, 0};
  double *out = (double *)malloc(sizeof(double) * 4 * (size_t)P);
  ode_rk4(&rp, 4, Y, padding + 1, P, m, out); /* :84-86 */
  for (int i = 1; i <= P; ++i) st->S[padding + i] = out[(i - 1) * 4];
  double *s = vec(P, 0);
  convolute(st->S, T, padding + 1, padding + P, &died, s); /* :90-91 */
  double data_offset = EPI_INVALID_DATA_OFFSET;
  for (int i = 1; i <= P; ++i) {
    /* state$died is 0 (not NA) on 1..padding here (:70), and 0 > threshold is
     * false for threshold >= 0 */
    double d = (N - s[i]) * died_rate;
    if (d > mort_lockdown_threshold) { data_offset = padding + i - D->lockdown_offset; break; }
  }
  if (mort_lockdown_threshold < 0) data_offset = 1 - D->lockdown_offset; /* died[1] = 0 > threshold */
  if (data_offset != EPI_INVALID_DATA_OFFSET) { /* :96-106 */
    rp.ldts = data_offset + D->lockdown_offset;
    rp.ldte = data_offset + D->lockdown_offset + D->lockdown_transition_period;
    ode_rk4(&rp, 4, Y, padding + 1, P, m, out);
  }
  st->full = (double *)malloc(sizeof(double) * 4 * (size_t)P);
  memcpy(st->full, out, sizeof(double) * 4 * (size_t)P);
  st->yout = (double *)malloc(sizeof(double) * 2 * (size_t)P);
  yout_rows(&rp, 4, out, padding + 1, P, st->yout);
  for (int i = 1; i <= P; ++i) st->S[padding + i] = out[(i - 1) * 4];
  free(out);
  convolute(st->S, T, padding + 1, padding + P, &hosp, s); /* :115-116 */
  for (int i = 1; i <= P; ++i) hospv[padding + i] = (N - s[i]) * hosp_rate;
  convolute(st->S, T, padding + 1, padding + P, &died, s); /* :118-119 */
  for (int i = 1; i <= P; ++i) diedv[padding + i] = (N - s[i]) * died_rate;
  free(s);
  st->deadi = vec(T, 0);
  st->hospi = vec(T, 0);
  st->deadi[1] = diedv[1]; st->hospi[1] = hospv[1]; /* :121-122 */
  for (int t = 2; t <= T; ++t) {
    st->deadi[t] = diedv[t] - diedv[t - 1];
    st->hospi[t] = hospv[t] - hospv[t - 1];
  }
  free(hospv); free(diedv);
  st->offset = (int)data_offset;
  return 0;
}

/* calclogp(params): R/models/model-simple-ode-drj-cmp.R:156-179.  NB the driver
 * passes the UNTRANSFORMED sampler vector (R/MCMC/fitMCMC-drj.R:51); no prior
 * term reads params[5]. */
static double simple_calclogp(const double *params) {
  const double Tinc = 3;
  const double Tinf0 = params[3], Tinft = params[4], died_latency = params[7];
  const double G0 = Tinc + Tinf0 / 2, Gt = Tinc + Tinft / 2;
  double lp = 0;
  lp = lp + oracle_dnorm_log(died_latency, 21, 4);
  lp = lp + oracle_dnorm_log(G0, 5, 1);
  lp = lp + oracle_dnorm_log(G0 - Gt, 0, 1.5);
  return lp;
}

/* calclogl(params, x): R/models/model-simple-ode-drj-cmp.R:183-257; terms[0..2] =
 * loglLD, loglH, loglD */
static double simple_calclogl(const epi_data *D, int flavour, const double *params, int period, int m,
                              double *terms, int *offset_out, int *pad_out, int *status) {
  simple_state_t st;
  if (simple_calculate_model(D, flavour, params, period, m, &st)) {
    *status |= EPI_ST_UNSUPPORTED;
    for (int i = 0; i < 3; ++i) terms[i] = NA_REAL;
    *offset_out = 0; *pad_out = 0;
    return NA_REAL;
  }
  *offset_out = st.offset;
  *pad_out = st.pad;
  int offset = st.offset;
  if (offset == EPI_INVALID_DATA_OFFSET) {
    offset = 1;
    *status |= EPI_ST_INVALID_OFFSET;
  }
  const double mort_lockdown_threshold = params[8];
  double lLD = oracle_dnbinom_mu_log(D->total_deaths_at_lockdown, D->mort_nbinom_size,
                                     pmax_(0.1, mort_lockdown_threshold)); /* :203-204 */
  int dstart, dend;
  double one[2];
  window(offset, D->n_dhospi, st.T, &dstart, &dend);
  one[1] = D->hosp_nbinom_size;
  double lH = sum_dnbinom(D->dhospi - 1, D->n_dhospi, 1, D->n_dhospi, st.hospi, st.T, dstart, dend, 0.1, one, 1);
  window(offset, D->n_dmorti, st.T, &dstart, &dend);
  one[1] = D->mort_nbinom_size;
  double lD = sum_dnbinom(D->dmorti - 1, D->n_dmorti, 1, D->n_dmorti, st.deadi, st.T, dstart, dend, 0.1, one, 1);
  terms[0] = lLD; terms[1] = lH; terms[2] = lD;
  double result = lLD + lH + lD; /* :245 */
  if (isnan(result)) *status |= EPI_ST_NA;
  simple_state_free(&st);
  return result;
}

/* init/df_params: R/models/model-simple-ode-drj-cmp.R:263-270.  Ne flavour adds
 * Nef in (0.05, 1) (README.md:91-93), init 0.7. */
static void simple_df_params(const epi_data *D, int flavour, double *mn, double *mx, double *init) {
  const double tdl = D->total_deaths_at_lockdown;
  const double i_[9] = {2.9, 0.9, 3, 3, log(0.02), 10, 20, tdl, 0.7};
  const double mn_[9] = {0.1, 0.1, 0.1, 0.1, log(0.001), 5, 5, 0, 0.05};
  const double mx_[9] = {8, 8, 8, 8, log(0.5), 30, 40, fmax(D->dmort_last / 10, tdl * 10), 1.0};
  int d = flavour == EPI_FLAVOUR_NE ? 9 : 8;
  for (int i = 0; i < d; ++i) { mn[i] = mn_[i]; mx[i] = mx_[i]; init[i] = i_[i]; }
}

/* ------------------------------------------------------------- public ----- */

int oracle_n_params(int flavour) {
  return flavour == EPI_FLAVOUR_AGE2X ? 52 : flavour == EPI_FLAVOUR_AGE2Q ? 45 : flavour == EPI_FLAVOUR_AGE2I ? 36 : flavour == EPI_FLAVOUR_NE ? 9 : 8;
}

static int is_age(int flavour) { return flavour == EPI_FLAVOUR_AGE2Q || flavour == EPI_FLAVOUR_AGE2I || flavour == EPI_FLAVOUR_AGE2X; }

int oracle_n_series(int flavour) { return is_age(flavour) ? 8 : 3; }

int oracle_df_params(const epi_model_desc *desc, const epi_data *D, double *mn, double *mx, double *init) {
  if (desc->flavour == EPI_FLAVOUR_AGE2Q)
    age_df_params(D, mn, mx, init);
  else if (desc->flavour == EPI_FLAVOUR_AGE2I)
    age2i_df_params(D, mn, mx, init);
  else if (desc->flavour == EPI_FLAVOUR_AGE2X)
    age2x_df_params(D, mn, mx, init);
  else
    simple_df_params(D, desc->flavour, mn, mx, init);
  return 0;
}

/* transformParams: identity for age (model-age-groups2q.R:339-342); simple:
 * result[5] = exp(params[5]) (model-simple-ode-drj-cmp.R:136-142) */
static void transform_params(int flavour, const double *theta0 /*0-based*/, int d, double *params1 /*1-based*/) {
  for (int i = 0; i < d; ++i) params1[i + 1] = theta0[i];
  if (!is_age(flavour)) params1[5] = exp(theta0[4]);
}

/* One sampler-space theta (0-based, d doubles) -> logl, logp, detail.
 * Mirrors R/MCMC/fitMCMC-drj.R:19-21 (loglike) and :51 (logprior = calclogp on
 * the untransformed vector). */
void oracle_eval_one(const epi_model_desc *desc, const epi_data *D, const double *theta, double *logl,
                     double *logp, int32_t *status, int32_t *offset, int32_t *padding, double *terms8) {
  const int d = oracle_n_params(desc->flavour);
  double params[EPI_MAX_PARAMS + 1], raw[EPI_MAX_PARAMS + 1];
  double terms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int st = 0, off = 0, pad = 0;
  transform_params(desc->flavour, theta, d, params);
  for (int i = 0; i < d; ++i) raw[i + 1] = theta[i];
  double mn[EPI_MAX_PARAMS], mx[EPI_MAX_PARAMS], in[EPI_MAX_PARAMS];
  oracle_df_params(desc, D, mn, mx, in);
  for (int i = 0; i < d; ++i)
    if (!(theta[i] >= mn[i] && theta[i] <= mx[i])) st |= EPI_ST_OUT_OF_BOX;
  double ll, lp;
  if (desc->flavour == EPI_FLAVOUR_AGE2Q) {
    ll = age_calclogl(D, params, desc->period, desc->substeps, terms, &off, &pad, &st);
    lp = age_calclogp(D, raw);
  } else if (desc->flavour == EPI_FLAVOUR_AGE2I) {
    ll = age2i_calclogl(D, params, desc->period, desc->substeps, terms, &off, &pad, &st);
    lp = age2i_calclogp(D, raw);
  } else if (desc->flavour == EPI_FLAVOUR_AGE2X) {
    ll = age2x_calclogl(D, params, desc->period, desc->substeps, terms, &off, &pad, &st);
    lp = age2x_calclogp(D, raw);
  } else {
    ll = simple_calclogl(D, desc->flavour, params, desc->period, desc->substeps, terms, &off, &pad, &st);
    lp = simple_calclogp(raw);
  }
  if (logl) *logl = ll;
  if (logp) *logp = lp;
  if (status) *status = st;
  if (offset) *offset = off;
  if (padding) *padding = pad;
  if (terms8) memcpy(terms8, terms, sizeof terms);
}

/* Batched evaluation over n proposals (theta: n x d row-major), split over
 * n_threads POSIX threads (interleaved blocks of 4) — this is the cpu_baseline /
 * --impl reference arm of bench.py. */
typedef struct {
  const epi_model_desc *desc; const epi_data *D; const double *theta; int64_t n; int d;
  int tid, nt;
  double *logl, *logp; int32_t *status, *offset, *padding; double *terms8;
} job_t;

static void *eval_worker(void *arg) {
  job_t *j = (job_t *)arg;
  for (int64_t b = (int64_t)j->tid * 4; b < j->n; b += (int64_t)j->nt * 4) {
    for (int64_t i = b; i < b + 4 && i < j->n; ++i)
      oracle_eval_one(j->desc, j->D, j->theta + i * j->d, j->logl ? j->logl + i : 0,
                      j->logp ? j->logp + i : 0, j->status ? j->status + i : 0,
                      j->offset ? j->offset + i : 0, j->padding ? j->padding + i : 0,
                      j->terms8 ? j->terms8 + i * 8 : 0);
  }
  return 0;
}

int oracle_eval(const epi_model_desc *desc, const epi_data *D, const double *theta, int64_t n,
                int n_threads, double *logl, double *logp, int32_t *status, int32_t *offset,
                int32_t *padding, double *terms8) {
  const int d = oracle_n_params(desc->flavour);
  int nt = n_threads > 1 ? n_threads : 1;
  if (nt > 256) nt = 256;
  job_t jobs[256];
  pthread_t th[256];
  for (int t = 0; t < nt; ++t) {
    job_t j = {desc, D, theta, n, d, t, nt, logl, logp, status, offset, padding, terms8};
    jobs[t] = j;
  }
  if (nt == 1) {
    eval_worker(&jobs[0]);
    return 0;
  }
  for (int t = 0; t < nt; ++t) pthread_create(&th[t], 0, eval_worker, &jobs[t]);
  for (int t = 0; t < nt; ++t) pthread_join(th[t], 0);
  return 0;
}

/* calculateModel trajectories for one MODEL-space theta: out is
 * [n_series][period], days padding+1..padding+period; age series order
 * y.S,o.S,y.casei,o.casei,y.hospi,o.hospi,y.deadi,o.deadi; simple: S,hospi,deadi.
 * states (optional) receives the final ode() rows [period][neq]. */
int oracle_trajectory(const epi_model_desc *desc, const epi_data *D, const double *theta_model,
                      int period, double *out, double *states, double *ext, int32_t *offset, int32_t *padding) {
  const int d = oracle_n_params(desc->flavour);
  double params[EPI_MAX_PARAMS + 1];
  for (int i = 0; i < d; ++i) params[i + 1] = theta_model[i];
  if (is_age(desc->flavour)) {
    age_state_t st;
    age2x_state_t sx;
    memset(&sx, 0, sizeof sx);
    if (desc->flavour == EPI_FLAVOUR_AGE2X) {
      if (age2x_calculate_model(D, params, period, desc->substeps, &sx)) return -1;
      st = sx.a;
      memset(&sx.a, 0, sizeof sx.a); /* st owns the series now */
      age2x_state_free(&sx);
    } else if (desc->flavour == EPI_FLAVOUR_AGE2I ? age2i_calculate_model(D, params, period, desc->substeps, &st)
                                                  : age_calculate_model(D, params, period, desc->substeps, &st))
      return -1;
    const double *src[8] = {st.yS, st.oS, st.ycasei, st.ocasei, st.yhospi, st.ohospi, st.ydeadi, st.odeadi};
    for (int q = 0; q < 8; ++q)
      for (int i = 1; i <= period; ++i) out[(size_t)q * period + i - 1] = src[q][st.pad + i];
    if (states) memcpy(states, st.full, sizeof(double) * 10 * (size_t)period);
    if (ext) {
      /* the state fields calculateModel derives from the ode() matrix (model-age-groups2q.R:216-227;
       * 2i :199-210, where o.E = out[,8] and o.I = out[,9] + out[,10]): y.E,y.I,y.R,o.E,o.I,o.R,Re,Rt,y.Re,o.Re */
      const int two_i = desc->flavour == EPI_FLAVOUR_AGE2I;
      for (int i = 0; i < period; ++i) {
        const double *y = st.full + (size_t)i * 10, *yo = st.yout + (size_t)i * 4;
        ext[0 * (size_t)period + i] = y[1] + y[2];
        ext[1 * (size_t)period + i] = y[3];
        ext[2 * (size_t)period + i] = y[4];
        ext[3 * (size_t)period + i] = two_i ? y[6] : y[6] + y[7];
        ext[4 * (size_t)period + i] = two_i ? y[7] + y[8] : y[8];
        ext[5 * (size_t)period + i] = y[9];
        for (int q = 0; q < 4; ++q) ext[(6 + q) * (size_t)period + i] = yo[q];
      }
    }
    if (offset) *offset = st.offset;
    if (padding) *padding = st.pad;
    age_state_free(&st);
  } else {
    simple_state_t st;
    if (simple_calculate_model(D, desc->flavour, params, period, desc->substeps, &st)) return -1;
    const double *src[3] = {st.S, st.hospi, st.deadi};
    for (int q = 0; q < 3; ++q)
      for (int i = 1; i <= period; ++i) out[(size_t)q * period + i - 1] = src[q][st.pad + i];
    if (states) memcpy(states, st.full, sizeof(double) * 4 * (size_t)period);
    if (ext) { /* E, I, R, Re, Rt (model-simple-ode-drj-cmp.R:108-113) */
      for (int i = 0; i < period; ++i) {
        for (int q = 0; q < 3; ++q) ext[q * (size_t)period + i] = st.full[(size_t)i * 4 + 1 + q];
        ext[3 * (size_t)period + i] = st.yout[(size_t)i * 2];
        ext[4 * (size_t)period + i] = st.yout[(size_t)i * 2 + 1];
      }
    }
    if (offset) *offset = st.offset;
    if (padding) *padding = st.pad;
    simple_state_free(&st);
  }
  return 0;
}

/* latency profile exposed for unit tests: kind 0 = gamma (age), 1 = normal
 * (simple, pass the NEGATIVE latency as the model does).  values is 0-based,
 * returns n, and R's kbegin/kend. */
int oracle_profile(int kind, double mean, double sd, double *values, int *kbegin, int *kend) {
  profile_t p;
  int rc = kind == 0 ? calc_gamma_profile(mean, sd, &p) : calc_convolve_profile(mean, sd, &p);
  if (rc) return -1;
  for (int i = 1; i <= p.n; ++i) values[i - 1] = p.values[i];
  *kbegin = p.kbegin;
  *kend = p.kend;
  return p.n;
}
