// epi_sampler.cu — K2: the fused on-device Metropolis / DE-MC propose-accept
// kernel and the sampler entry points of include/epimcmc.h.
//
// The reference drives calclogl/calclogp one parameter vector at a time from
// drjacoby's host loop (R/MCMC/fitMCMC-drj.R:48-57).  Here the chain population
// lives on the device for the whole run: per iteration the host only enqueues
// the four K1 launches on the proposal buffer plus ONE k_accept_propose launch
// that (a) takes the Metropolis decision of iteration i for every chain, (b)
// commits accepted states, (c) optionally records the thinned draw and (d)
// writes the proposal of iteration i+1.  Random numbers are counter-based
// Philox4x32-10 keyed by (seed; chain id, iteration, use), so a run is
// reproducible for any sharding of chains over GPUs.
//
// HBM traffic of k_accept_propose per chain and iteration (d parameters):
// read theta (8d) + theta' (8d) + 4 log densities + status (36 B), write theta'
// (8d) always and theta (8d) + 2 log densities on accept; DE-MC adds two
// partner reads (16d).  See DESIGN.md §4 for the roofline.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/epimcmc.h"
#include "epi_handle.h"
#include "epi_launch.h"

namespace epi {

struct SamplerState {
  epi_sampler_cfg cfg{};
  int64_t ld = 0;
  double *cur = nullptr, *prop = nullptr;  // SoA d x ld
  double *logl = nullptr, *logp = nullptr, *logl_p = nullptr, *logp_p = nullptr;
  int *status_p = nullptr;
  double *lscale = nullptr;                  // per-chain log step multiplier (Robbins-Monro)
  double *psd = nullptr;                     // per-parameter sd of the population (proposal shape)
  double *cov = nullptr, *chol = nullptr;    // per rung: d x d population covariance and its Cholesky factor (lower, row-major)
  double *chol_tmp = nullptr;                // scratch of k_shape_factor
  int *shape_ok = nullptr, *shape_ready = nullptr;  // per-rung verdicts of the last factorisation; [1]: a shape is installed
  int n_rungs = 1, cpr = 0;                  // parallel tempering: rungs and chains per rung
  double *beta_rung = nullptr;               // device, n_rungs
  unsigned long long *swaps = nullptr;       // device: [0] proposed, [1] accepted
  unsigned long long *acc = nullptr;         // accepted moves, device scalar
  double *pop_own = nullptr;                 // own snapshot (1 rank)
  const double *pop = nullptr;               // population snapshot [rank][d][ld_rank] as gathered
  int pop_ranks = 0, pop_per_rank = 0;
  int64_t pop_ld = 0;
  // the snapshot the proposal kernel reads: one contiguous d-vector per member, [rank * per_rank + i][d].  A DE-MC
  // proposal gathers ~d/2 coordinates of two random members: in the gathered SoA block every one of them is its own
  // DRAM access (ncu, profiles/r2_ncu_k2.txt: 28 GB moved for 5.5 GB of algorithmic bytes at 2^22 chains, the
  // DRAM pipe 75 % busy), in a member-major row they share three 128-byte lines
  double *pop_aos = nullptr;
  int64_t pop_aos_cap = 0;
  double *draws = nullptr, *draws_lp = nullptr;  // [capacity][d][ld], [capacity][ld]
  int thin = 0, capacity = 0, kept = 0;
  int adapt = 0;
  double adapt_target = 0.0;  // <= 0: shape adaptation only (classical AM / DE-MC step sizes)
  int64_t iter = 0, evals = 0;
  // EPI_SAMPLER_DRJ: per-chain, per-parameter bandwidths on the phi scale (d x ld), the parameters with a
  // non-degenerate box (the sweep order), sub-iteration counter (the Philox "iteration" of a single-parameter
  // update) and the number of sweeps adaptation has seen (Robbins-Monro gain 1/sqrt(k))
  double *bw = nullptr;
  std::vector<int> free_params;
  int64_t sub = 0, adapt_sweeps = 0;
  bool replicates = false;
};

// ----------------------------------------------------------------- Philox4x32
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in (0,1) from 32 bits
__host__ __device__ inline double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

// two standard normals from two 32-bit words (Box-Muller, FP64)
__device__ inline void box_muller(uint32_t a, uint32_t b, double &z0, double &z1) {
  const double r = sqrt(-2.0 * log(u01(a)));
  double s, c;
  sincospi(2.0 * u01(b), &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// fold x into [lo, hi] by reflection (keeps a symmetric proposal symmetric)
__device__ inline double reflect(double x, double lo, double hi) {
  const double w = hi - lo;
  if (!(w > 0)) return lo;
  if (x >= lo && x <= hi) return x;  // the usual case: nothing to fold (and no fmod)
  double y = fmod(x - lo, 2.0 * w);
  if (y < 0) y += 2.0 * w;
  if (y > w) y = 2.0 * w - y;
  return lo + y;
}

enum { USE_ACCEPT = 0, USE_PARTNER = 1, USE_SWAP = 2, USE_NORMAL = 16, USE_CROSS = 64 };

// Accept/reject iteration `iter` (when do_accept) and write the proposal of iteration iter+1.
__global__ void __launch_bounds__(128)
k_accept_propose(int d, int64_t n, int64_t ld, int kind, uint64_t seed, int64_t chain_id0, int64_t iter, int do_accept,
                 double beta0, const double *__restrict__ beta_rung, int cpr, double scale, double de_eps, int flags, const double *__restrict__ box_min,
                 const double *__restrict__ box_max, double *__restrict__ cur, double *__restrict__ prop,
                 double *__restrict__ logl, double *__restrict__ logp, const double *__restrict__ logl_p,
                 const double *__restrict__ logp_p, const int *__restrict__ status_p, double *__restrict__ lscale,
                 int adapt, double adapt_target, unsigned long long *__restrict__ acc_total,
                 const double *__restrict__ pop, int pop_ranks, int pop_per_rank, int64_t pop_ld,
                 const double *__restrict__ psd_in, const double *__restrict__ chol_in, double *__restrict__ draw_out,
                 double *__restrict__ draw_lp, const int *__restrict__ shape_ready = nullptr) {
  // the proposal shape is measured and factorised on the device (k_shape_*): until the first population that
  // allowed it, *shape_ready is 0 and the fallbacks for "no shape yet" apply
  const bool shaped = !shape_ready || *shape_ready != 0;
  const double *__restrict__ psd = shaped ? psd_in : nullptr;
  const double *__restrict__ chol = shaped ? chol_in : nullptr;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const uint64_t cid = (uint64_t)(chain_id0 + i);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t r[4];
  bool accepted = false;
  const int rung = (beta_rung && live) ? (int)(i / cpr) : 0;
  const double beta = beta_rung ? beta_rung[rung] : beta0;
  if (live && do_accept) {
    philox4x32_10((uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)iter, USE_ACCEPT, k0, k1, r);
    const double ll0 = logl[i], lp0 = logp[i], ll1 = logl_p[i], lp1 = logp_p[i];
    const double cur_post = beta * ll0 + lp0, new_post = beta * ll1 + lp1;
    const bool new_ok = isfinite(new_post) && !(status_p[i] & (EPI_ST_NA | EPI_ST_UNSUPPORTED));
    const bool cur_ok = isfinite(cur_post);
    accepted = new_ok && (!cur_ok || log(u01(r[0])) < new_post - cur_post);
    if (accepted) {
      // (unrolled: with a run-time trip count every load -> store pair waited out its own L2 round trip; ncu,
      // prof_ap: 71 % of this kernel's stall samples were long-scoreboard at 3.5 warps per scheduler)
#pragma unroll 8
      for (int p = 0; p < d; ++p) cur[(int64_t)p * ld + i] = prop[(int64_t)p * ld + i];
      logl[i] = ll1;
      logp[i] = lp1;
    }
    if (adapt && adapt_target > 0.0) {  // optional Robbins-Monro on the per-chain log step size, gain 1/sqrt(iter+1)
      lscale[i] += ((accepted ? 1.0 : 0.0) - adapt_target) * rsqrt((double)(iter + 1));
    }
  }
  if (do_accept) {  // block-aggregated acceptance counter
    const unsigned b = __ballot_sync(0xffffffffu, accepted);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(acc_total, (unsigned long long)__popc(b));
  }
  if (!live) return;
  if (draw_out) {
#pragma unroll 8
    for (int p = 0; p < d; ++p) draw_out[(int64_t)p * ld + i] = cur[(int64_t)p * ld + i];
    draw_lp[i] = logl[i] + logp[i];
  }
  // ---- proposal for iteration iter+1
  const uint32_t it1 = (uint32_t)(iter + 1);
  const double step = scale * exp(lscale[i]);
  const double *z1 = nullptr, *z2 = nullptr;
  double gam = 0.0;
  philox4x32_10((uint32_t)cid, (uint32_t)(cid >> 32), it1, USE_PARTNER, k0, k1, r);
  // scale mixture: each proposal draws its own step factor u ~ U(0.05, 1), independent of the
  // state, so the kernel stays symmetric.  On this rugged target (ceil/floor kernel extents,
  // integer data_offset) a single "optimal" step is accepted ~2 % of the time; the mixture
  // keeps both the local moves and the occasional cell-crossing ones.
  const double umix = 0.05 + 0.95 * u01(r[2]);
  if (kind == EPI_SAMPLER_DEMC && pop) {
    // partners come from the chain's own temperature rung (cpr members per rank)
    const uint32_t per_rank = beta_rung ? (uint32_t)cpr : (uint32_t)pop_per_rank;
    const uint32_t first = beta_rung ? (uint32_t)rung * (uint32_t)cpr : 0u;
    const uint32_t tot = (uint32_t)pop_ranks * per_rank;
    uint32_t a = (uint32_t)(((uint64_t)r[0] * tot) >> 32);
    uint32_t b = (uint32_t)(((uint64_t)r[1] * (tot - 1)) >> 32);
    if (b >= a) b += 1;  // two distinct members of the snapshot
    // member-major snapshot (k_pop_to_aos): member = rank * pop_per_rank + chain within the rank
    z1 = pop + ((int64_t)(a / per_rank) * pop_per_rank + first + (a % per_rank)) * d;
    z2 = pop + ((int64_t)(b / per_rank) * pop_per_rank + first + (b % per_rank)) * d;
  }
  // Subspace (crossover) updates, as in DREAM (Vrugt et al. 2009): a DE-MC proposal moves only a random subset
  // of the coordinates, each chosen with probability CR in {0.1, 0.25, 0.5, 1} (one of the four per proposal,
  // independent of the state: the kernel stays symmetric), with gamma = 2.38 / sqrt(2 d') for the d' chosen ones.
  // On this target a handful of parameters (latencies, kernel sds, the death threshold) make logl jump at
  // ceil/floor and integer-offset boundaries; full-dimensional moves cross such a boundary almost always and
  // the integrated autocorrelation time was ~1,500 iterations, see profiles/r1_summary.md.
  unsigned long long sel = ~0ull;  // bit p: coordinate p moves
  int dsel = d;
  const bool hop = (it1 % 10u) == 0u;  // ter Braak (2006): every 10th proposal gamma = 1, all coordinates
  if (z1 && !hop && !(flags & EPI_SAMPLER_FLAG_FULLDIM)) {
    const uint32_t crsel = r[3] & 3u;
    const double CR = crsel == 0 ? 0.1 : crsel == 1 ? 0.25 : crsel == 2 ? 0.5 : 1.0;
    // The slot of CR = 1 (a quarter of the proposals) moves ONE coordinate instead, chosen uniformly and
    // independently of the state, with gamma = 2.38 / sqrt(2): the population's analogue of drjacoby's one-parameter
    // updates.  On the age posterior (ceil/floor kernel extents, integer data_offset) it raised the ESS growth per
    // evaluation 3.9x (profiles/r2_summary.md); full-dimensional moves remain as the gamma = 1 hop of every tenth
    // proposal.  EPI_SAMPLER_FLAG_NO_SINGLE restores CR = 1 in that slot.
    if (!(flags & EPI_SAMPLER_FLAG_NO_SINGLE) && crsel == 3u) {
      const int p = (int)(((uint64_t)(r[3] >> 2) * (uint32_t)d) >> 30);
      sel = 1ull << p;
      dsel = 1;
    } else if (CR < 1.0) {
      sel = 0ull;
      dsel = 0;
      uint32_t c[4];
      for (int p0 = 0; p0 < d; p0 += 4) {
        philox4x32_10((uint32_t)cid, (uint32_t)(cid >> 32), it1, USE_CROSS + (uint32_t)(p0 >> 2), k0, k1, c);
        for (int q = 0; q < 4 && p0 + q < d; ++q)
          if (u01(c[q]) < CR) { sel |= 1ull << (p0 + q); ++dsel; }
      }
      if (dsel == 0) {  // at least one coordinate
        const int p = (int)(((uint64_t)r[2] * (uint32_t)d) >> 32);
        sel = 1ull << p;
        dsel = 1;
      }
    }
  }
  if (z1) gam = hop ? 1.0 : umix * step * 2.38 / sqrt(2.0 * dsel);
  if (z1) {
    // DE-MC: x + gamma (z1 - z2) + e, e ~ N(0, (de_eps * population sd)^2), on the chosen coordinates; the others
    // keep their value bit for bit.  The normals of a group of four coordinates (one Philox call, as in the RWM
    // branch below) are only generated when one of the four moves: with CR = 0.1 most groups do not.
    for (int p0 = 0; p0 < d; p0 += 4) {
      // the loads of the group first (current values, and the two partners' values of the moving coordinates), so
      // that they are in flight together while the normals are generated
      const unsigned g = (unsigned)(sel >> p0) & 0xFu;
      double x[4], dz[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = (p0 + q < d) ? cur[(int64_t)(p0 + q) * ld + i] : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool mv = ((g >> q) & 1u) && p0 + q < d;
        const double a = mv ? z1[p0 + q] : 0.0, b = mv ? z2[p0 + q] : 0.0;
        dz[q] = a - b;
      }
      double z[4] = {0.0, 0.0, 0.0, 0.0};
      if (g) {
        philox4x32_10((uint32_t)cid, (uint32_t)(cid >> 32), it1, USE_NORMAL + (uint32_t)(p0 >> 2), k0, k1, r);
        box_muller(r[0], r[1], z[0], z[1]);
        box_muller(r[2], r[3], z[2], z[3]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int p = p0 + q;
        if (p >= d) break;
        double y = x[q];
        if ((g >> q) & 1u) {
          const double lo = box_min[p], hi = box_max[p];
          const double w = psd ? psd[p] : (hi - lo);
          y = reflect(x[q] + gam * dz[q] + de_eps * w * z[q], lo, hi);
        }
        prop[(int64_t)p * ld + i] = y;
      }
    }
    return;
  }
  double zn[EPI_MAX_PARAMS];
  for (int p0 = 0; p0 < d; p0 += 4) {
    philox4x32_10((uint32_t)cid, (uint32_t)(cid >> 32), it1, USE_NORMAL + (uint32_t)(p0 >> 2), k0, k1, r);
    box_muller(r[0], r[1], zn[p0], zn[p0 + 1]);
    box_muller(r[2], r[3], zn[p0 + 2], zn[p0 + 3]);
  }
  const double rw = umix * step * 2.38 * rsqrt((double)d);
  for (int p = 0; p < d; ++p) {
    const double lo = box_min[p], hi = box_max[p];
    const double x = cur[(int64_t)p * ld + i];
    double y;
    if (chol) {
      // RWM with the population covariance: x + step * 2.38/sqrt(d) * L z  (adaptive Metropolis shape,
      // measured during burn-in only)
      double a = 0.0;
      const double *L = chol + (size_t)rung * d * d;
      for (int q = 0; q <= p; ++q) a = fma(L[p * d + q], zn[q], a);
      y = x + rw * a;
    } else {
      y = x + step * 0.02 * rsqrt((double)d) * (hi - lo) * zn[p];  // before any adaptation: 2 % of the box / sqrt(d)
    }
    prop[(int64_t)p * ld + i] = reflect(y, lo, hi);
  }
}

// drjacoby's update, one parameter per launch (R/MCMC/fitMCMC-drj.R:48-57 drives drjacoby::run_mcmc; the
// algorithm restated here is the published one of the package, which is not part of the reference tree):
//   phi = log((theta - a) / (b - theta));  phi' ~ N(phi, bw);  theta' = (b e^phi' + a) / (1 + e^phi')
//   MH  = beta (logl' - logl) + (logp' - logp) + log((theta' - a)(b - theta')) - log((theta - a)(b - theta))
//   Robbins-Monro while adapting: log bw += gain ((accepted ? 1 : 0) - target), gain = 1 / sqrt(sweep)
// `prop` equals `cur` in every coordinate but the one being proposed, so a launch touches two coordinates:
// it decides the pending move of p_acc (restoring prop[p_acc] on reject) and writes the move of p_new.
// After a rung swap (full_sync) the whole proposal row is rewritten from the current state.
__global__ void __launch_bounds__(128)
k_drj_step(int d, int64_t n, int64_t ld, uint64_t seed, int64_t chain_id0, int64_t sub_acc, int p_acc, int p_new,
           int full_sync, const double *__restrict__ beta_rung, int cpr, const double *__restrict__ box_min,
           const double *__restrict__ box_max, double *__restrict__ cur, double *__restrict__ prop,
           double *__restrict__ logl, double *__restrict__ logp, const double *__restrict__ logl_p,
           const double *__restrict__ logp_p, const int *__restrict__ status_p, double *__restrict__ bw, int adapt,
           double gain, double target, unsigned long long *__restrict__ acc_total, double *__restrict__ draw_out,
           double *__restrict__ draw_lp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const uint64_t cid = (uint64_t)(chain_id0 + i);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t r[4];
  bool accepted = false;
  if (live && p_acc >= 0) {
    const double beta = beta_rung ? beta_rung[i / cpr] : 1.0;
    philox4x32_10((uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)sub_acc, USE_ACCEPT, k0, k1, r);
    const double a = box_min[p_acc], b = box_max[p_acc];
    const double x = cur[(int64_t)p_acc * ld + i], y = prop[(int64_t)p_acc * ld + i];
    const double ll0 = logl[i], lp0 = logp[i], ll1 = logl_p[i], lp1 = logp_p[i];
    const double adj = log((y - a) * (b - y)) - log((x - a) * (b - x));
    const double mh = beta * (ll1 - ll0) + (lp1 - lp0) + adj;
    const bool new_ok = isfinite(beta * ll1 + lp1) && !(status_p[i] & (EPI_ST_NA | EPI_ST_UNSUPPORTED)) && y > a && y < b;
    const bool cur_ok = isfinite(beta * ll0 + lp0);
    accepted = new_ok && (!cur_ok || log(u01(r[0])) < mh);
    if (accepted) {
      cur[(int64_t)p_acc * ld + i] = y;
      logl[i] = ll1;
      logp[i] = lp1;
    } else if (!full_sync) {
      prop[(int64_t)p_acc * ld + i] = x;
    }
    if (adapt) bw[(int64_t)p_acc * ld + i] *= exp(gain * ((accepted ? 1.0 : 0.0) - target));
  }
  if (p_acc >= 0) {
    const unsigned bal = __ballot_sync(0xffffffffu, accepted);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(acc_total, (unsigned long long)__popc(bal));
  }
  if (!live) return;
  if (draw_out) {
#pragma unroll 8
    for (int p = 0; p < d; ++p) draw_out[(int64_t)p * ld + i] = cur[(int64_t)p * ld + i];
    draw_lp[i] = logl[i] + logp[i];
  }
  if (full_sync) {
#pragma unroll 8
    for (int p = 0; p < d; ++p) prop[(int64_t)p * ld + i] = cur[(int64_t)p * ld + i];
  }
  if (p_new >= 0) {
    philox4x32_10((uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)(sub_acc + 1), USE_NORMAL, k0, k1, r);
    double z0, z1;
    box_muller(r[0], r[1], z0, z1);
    const double a = box_min[p_new], b = box_max[p_new], w = b - a;
    const double x = cur[(int64_t)p_new * ld + i];
    // a state exactly on the boundary has phi = +-inf: step from just inside (the Jacobian term of the true
    // state is +inf, so whatever is proposed from there is accepted and the chain leaves the boundary)
    const double xi = fmin(fmax(x, a + 1e-12 * w), b - 1e-12 * w);
    const double phi = log((xi - a) / (b - xi)) + bw[(int64_t)p_new * ld + i] * z0;
    prop[(int64_t)p_new * ld + i] = a + w / (1.0 + exp(-phi));
  }
}

// synthetic population for epi_k2_bench: uniform in the box, Philox keyed like the sampler
__global__ void k_fill_uniform(int d, int64_t n, int64_t ld, uint64_t seed, const double *__restrict__ box_min,
                               const double *__restrict__ box_max, double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r[4];
  for (int p0 = 0; p0 < d; p0 += 4) {
    philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0xFFFFFFFEu, (uint32_t)p0, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int q = 0; q < 4 && p0 + q < d; ++q)
      out[(int64_t)(p0 + q) * ld + i] = box_min[p0 + q] + (box_max[p0 + q] - box_min[p0 + q]) * u01(r[q]);
  }
}
// log density of the proposals of epi_k2_bench: a fixed, hashed fraction of the chains accepts every proposal
// (+1e6 on the first one, equal densities afterwards), the others none (-inf)
__global__ void k_fill_accept_pattern(int64_t n, uint64_t seed, double frac, double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0xFFFFFFFDu, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  out[i] = u01(r[0]) < frac ? 1e6 : -INFINITY;
}
__global__ void k_fill_const(int64_t n, double v, double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

// snapshot [rank][d][ld] -> member-major [rank * per_rank + i][d]: 32 members per block through a shared-memory tile
// (coalesced reads along the chains, coalesced writes along the parameters)
__global__ void __launch_bounds__(256) k_pop_to_aos(const double *__restrict__ pop, int per_rank, int64_t ld, int d,
                                                    double *__restrict__ out) {
  __shared__ double tile[EPI_MAX_PARAMS][33];
  const int rank = blockIdx.y, lane = threadIdx.x & 31, row = threadIdx.x >> 5;  // 8 rows of 32 lanes
  const int i0 = blockIdx.x * 32;
  const double *src = pop + (int64_t)rank * d * ld;
  for (int p = row; p < d; p += 8)
    if (i0 + lane < per_rank) tile[p][lane] = src[(int64_t)p * ld + i0 + lane];
  __syncthreads();
  const int nm = min(32, per_rank - i0);
  double *dst = out + ((int64_t)rank * per_rank + i0) * d;
  for (int k = threadIdx.x; k < nm * d; k += 256) dst[k] = tile[k % d][k / d];
}

// population covariance of the chains of this rank: one block per (p, q), q <= p
__global__ void __launch_bounds__(256) k_cov(const double *__restrict__ cur_all, int64_t n, int64_t ld, int d,
                                             double *__restrict__ cov_all) {
  // blockIdx.y = temperature rung: chains [rung * n, (rung + 1) * n)
  const double *cur = cur_all + (int64_t)blockIdx.y * n;
  double *cov = cov_all + (size_t)blockIdx.y * d * d;
  __shared__ double sa[256], sb[256], sab[256];
  int p = 0, rem = blockIdx.x;
  while (rem > p) { rem -= p + 1; ++p; }
  const int q = rem;
  const double *rp = cur + (int64_t)p * ld, *rq = cur + (int64_t)q * ld;
  const double cp = rp[0], cq = rq[0];  // shifts for a stable one-pass estimate
  double a = 0.0, b = 0.0, ab = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double u = rp[i] - cp, v = rq[i] - cq;
    a += u; b += v; ab = fma(u, v, ab);
  }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b; sab[threadIdx.x] = ab;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; sab[threadIdx.x] += sab[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double c = (sab[0] - sa[0] * sb[0] / (double)n) / (double)(n > 1 ? n - 1 : 1);
    cov[p * d + q] = c;
    cov[q * d + p] = c;
  }
}

// Parallel tempering: Metropolis swap of the states of chain j on rung r and rung r+1 for the rung
// pairs of the current parity (drjacoby's adjacent-rung swaps, R/MCMC/fitMCMC-drj.R:48-57).
__global__ void __launch_bounds__(128)
k_pt_swap(int d, int cpr, int n_rungs, int64_t ld, int parity, uint64_t seed, int64_t chain_id0, int64_t iter,
          const double *__restrict__ beta_rung, double *__restrict__ cur, double *__restrict__ logl,
          double *__restrict__ logp, unsigned long long *__restrict__ counters) {
  const int n_pairs = (n_rungs - parity) / 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool prop = false, acc = false;
  if (t < (int64_t)n_pairs * cpr) {
    const int r = parity + 2 * (int)(t / cpr), j = (int)(t % cpr);
    const int64_t i0 = (int64_t)r * cpr + j, i1 = i0 + cpr;
    const double l0 = logl[i0], l1 = logl[i1];
    if (isfinite(l0) && isfinite(l1)) {
      prop = true;
      uint32_t w[4];
      const uint64_t cid = (uint64_t)(chain_id0 + i0);
      philox4x32_10((uint32_t)cid, (uint32_t)(cid >> 32), (uint32_t)iter, USE_SWAP, (uint32_t)seed, (uint32_t)(seed >> 32), w);
      acc = log(u01(w[0])) < (beta_rung[r] - beta_rung[r + 1]) * (l1 - l0);
      if (acc) {
        for (int p = 0; p < d; ++p) {
          const double a = cur[(int64_t)p * ld + i0];
          cur[(int64_t)p * ld + i0] = cur[(int64_t)p * ld + i1];
          cur[(int64_t)p * ld + i1] = a;
        }
        logl[i0] = l1; logl[i1] = l0;
        const double p0 = logp[i0];
        logp[i0] = logp[i1]; logp[i1] = p0;
      }
    }
  }
  const unsigned bp = __ballot_sync(0xffffffffu, prop), ba = __ballot_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0) {
    if (bp) atomicAdd(counters, (unsigned long long)__popc(bp));
    if (ba) atomicAdd(counters + 1, (unsigned long long)__popc(ba));
  }
}

// raw Philox words for tests: out[4*j..] = philox(ctr = (j, 0, iter, use), key = seed)
__global__ void k_philox_probe(uint64_t seed, uint32_t iter, uint32_t use, int n, uint32_t *out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)j, 0u, iter, use, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  for (int q = 0; q < 4; ++q) out[4 * j + q] = r[q];
}

}  // namespace epi

using namespace epi;

#define CUDA_OK(h, call)                                                                              \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

void epi_sampler_free(epi_handle *h) {
  SamplerState *s = h->sampler;
  if (!s) return;
  cudaFree(s->cur); cudaFree(s->prop); cudaFree(s->logl); cudaFree(s->logp); cudaFree(s->logl_p); cudaFree(s->logp_p);
  cudaFree(s->status_p); cudaFree(s->lscale); cudaFree(s->psd); cudaFree(s->cov); cudaFree(s->chol); cudaFree(s->chol_tmp); cudaFree(s->shape_ok); cudaFree(s->shape_ready); cudaFree(s->beta_rung); cudaFree(s->swaps); cudaFree(s->acc); cudaFree(s->pop_own); cudaFree(s->pop_aos); cudaFree(s->draws); cudaFree(s->draws_lp); cudaFree(s->bw);
  delete s;
  h->sampler = nullptr;
}

// Proposal shape from the population covariance (per temperature rung / replicate), entirely on the device: round 1
// copied the covariances to the host, factorised there and copied back, behind two stream synchronisations per call —
// with 64 replicate populations ~1 ms of host work every ten burn-in iterations (the 8-GPU bench: burn-in 88 M evals/s
// against 105 M recorded).  One block per rung factorises into a scratch copy (column by column, the sums of a
// column's rows in parallel, every operation the explicitly rounded one of the host loop it replaces: same bits);
// the commit kernel installs ALL rungs or none (a degenerate population keeps the previous shape, as before).
__global__ void __launch_bounds__(64) k_shape_factor(const double *__restrict__ cov_all, int d, double *__restrict__ tmp_all,
                                                     int *__restrict__ ok_all) {
  const int r = blockIdx.x, t = threadIdx.x;
  const double *C = cov_all + (size_t)r * d * d;
  double *L = tmp_all + (size_t)r * d * d;
  __shared__ int s_ok;
  if (t == 0) {
    int ok = 1;
    for (int p = 0; p < d; ++p) {
      const double v = C[(size_t)p * d + p];
      if (!(v > 0) || !isfinite(v)) ok = 0;
    }
    s_ok = ok;
  }
  __syncthreads();
  if (!s_ok) { if (t == 0) ok_all[r] = 0; return; }
  double ridge = 1e-12;
  bool done = false;
  for (int attempt = 0; attempt < 8 && !done; ++attempt, ridge *= 100) {
    for (int i = t; i < d * d; i += blockDim.x) L[i] = 0.0;
    __syncthreads();
    if (t == 0) s_ok = 1;
    __syncthreads();
    for (int q = 0; q < d; ++q) {
      if (t == 0) {
        double a = __dadd_rn(C[(size_t)q * d + q], __dmul_rn(ridge, C[(size_t)q * d + q]));
        for (int k = 0; k < q; ++k) a = __dsub_rn(a, __dmul_rn(L[(size_t)q * d + k], L[(size_t)q * d + k]));
        if (!(a > 0)) s_ok = 0;
        else L[(size_t)q * d + q] = sqrt(a);
      }
      __syncthreads();
      if (!s_ok) break;
      const double lqq = L[(size_t)q * d + q];
      for (int p = q + 1 + t; p < d; p += blockDim.x) {
        double a = __dadd_rn(C[(size_t)p * d + q], 0.0);
        for (int k = 0; k < q; ++k) a = __dsub_rn(a, __dmul_rn(L[(size_t)p * d + k], L[(size_t)q * d + k]));
        L[(size_t)p * d + q] = a / lqq;
      }
      __syncthreads();
    }
    done = s_ok != 0;
    __syncthreads();
  }
  if (t == 0) ok_all[r] = done ? 1 : 0;
}

__global__ void __launch_bounds__(64) k_shape_commit(const double *__restrict__ cov_all, const double *__restrict__ tmp_all,
                                                     const int *__restrict__ ok_all, int d, int R, double *__restrict__ chol_all,
                                                     double *__restrict__ psd, int *__restrict__ ready) {
  for (int r = 0; r < R; ++r)
    if (!ok_all[r]) return;  // (uniform: every thread of every block reads the same flags)
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < d * d; i += blockDim.x) chol_all[(size_t)r * d * d + i] = tmp_all[(size_t)r * d * d + i];
  if (r == R - 1) {
    // jitter scale: the cold rung's spread
    for (int p = threadIdx.x; p < d; p += blockDim.x) psd[p] = sqrt(cov_all[((size_t)r * d + p) * d + p]);
    if (threadIdx.x == 0) *ready = 1;
  }
}

static int update_shape(epi_handle *h, SamplerState *s) {
  const int d = h->M.d, R = s->n_rungs;
  const int64_t n = s->cpr;
  k_cov<<<dim3(d * (d + 1) / 2, R), 256, 0, h->stream>>>(s->cur, n, s->ld, d, s->cov);
  k_shape_factor<<<R, 64, 0, h->stream>>>(s->cov, d, s->chol_tmp, s->shape_ok);
  k_shape_commit<<<R, 64, 0, h->stream>>>(s->cov, s->chol_tmp, s->shape_ok, d, R, s->chol, s->psd, s->shape_ready);
  h->launches += 3;
  CUDA_OK(h, cudaGetLastError());
  return EPI_OK;
}

// (re)build the member-major snapshot the proposal kernel reads from the gathered one (s->pop)
static int install_population(epi_handle *h, SamplerState *s) {
  const int d = h->M.d;
  const int64_t need = (int64_t)s->pop_ranks * s->pop_per_rank * d;
  if (need > s->pop_aos_cap) {
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    cudaFree(s->pop_aos);
    s->pop_aos = nullptr; s->pop_aos_cap = 0;
    CUDA_OK(h, cudaMalloc(&s->pop_aos, sizeof(double) * (size_t)need));
    s->pop_aos_cap = need;
  }
  k_pop_to_aos<<<dim3((unsigned)((s->pop_per_rank + 31) / 32), (unsigned)s->pop_ranks), 256, 0, h->stream>>>(
      s->pop, s->pop_per_rank, s->pop_ld, d, s->pop_aos);
  h->launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "k_pop_to_aos: %s", cudaGetErrorString(e));
  return EPI_OK;
}

// accept / propose for the chains [first, first + cnt) at iteration `iter` on `stream` (cnt < 0: all chains, the
// sampler's current iteration, the handle's stream).  The Philox counters carry the chain id, so a slice draws what
// the whole population would have drawn for its chains.
static int launch_ap(epi_handle *h, SamplerState *s, int do_accept, double *draw_out, double *draw_lp, int64_t first = 0,
                     int64_t cnt = -1, int64_t iter = -1, cudaStream_t stream = nullptr) {
  const int64_t n = cnt < 0 ? s->cfg.n_chains : cnt;
  if (iter < 0) iter = s->iter;
  if (!stream) stream = h->stream;
  k_accept_propose<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
      h->M.d, n, s->ld, s->cfg.kind, s->cfg.seed, s->cfg.chain_id0 + first, iter, do_accept, s->cfg.beta,
      s->n_rungs > 1 ? s->beta_rung : nullptr, s->cpr, s->cfg.scale,
      s->cfg.de_eps, s->cfg.reserved, h->M.box_min, h->M.box_max, s->cur + first, s->prop + first, s->logl + first, s->logp + first,
      s->logl_p + first, s->logp_p + first, s->status_p + first,
      s->lscale + first, s->adapt, s->adapt_target, s->acc, s->pop ? s->pop_aos : nullptr, s->pop_ranks, s->pop_per_rank, s->pop_ld,
      s->psd, s->cfg.kind == EPI_SAMPLER_RWM ? s->chol : nullptr,
      draw_out ? draw_out + first : nullptr, draw_lp ? draw_lp + first : nullptr, s->shape_ready);
  h->launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "k_accept_propose: %s", cudaGetErrorString(e));
  return EPI_OK;
}

// one k_drj_step launch for all chains on the handle's stream
static int launch_drj(epi_handle *h, SamplerState *s, int64_t sub_acc, int p_acc, int p_new, int full_sync, double *draw_out,
                      double *draw_lp) {
  const int64_t n = s->cfg.n_chains;
  const double gain = 1.0 / std::sqrt((double)(s->adapt_sweeps + 1));
  const double target = s->adapt_target > 0 ? s->adapt_target : 0.44;
  k_drj_step<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(
      h->M.d, n, s->ld, s->cfg.seed, s->cfg.chain_id0, sub_acc, p_acc, p_new, full_sync,
      s->n_rungs > 1 ? s->beta_rung : nullptr, s->cpr, h->M.box_min, h->M.box_max, s->cur, s->prop, s->logl, s->logp,
      s->logl_p, s->logp_p, s->status_p, s->bw, s->adapt, gain, target, s->acc, draw_out, draw_lp);
  h->launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "k_drj_step: %s", cudaGetErrorString(e));
  return EPI_OK;
}

// every failure path of init releases the half-built sampler: a later epi_sampler_* call then reports
// EPI_ERR_STATE instead of running on garbage
#define INIT_OK(h, call)                                                                              \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      epi_sampler_free(h);                                                                            \
      return epi_fail(h, EPI_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                      \
    }                                                                                                 \
  } while (0)

extern "C" int epi_sampler_init(epi_handle *h, const epi_sampler_cfg *cfg, const double *theta0) {
  if (!h || !cfg || cfg->n_chains < 1) return epi_fail(h, EPI_ERR_ARG, "bad sampler config");
  if (cfg->kind != EPI_SAMPLER_RWM && cfg->kind != EPI_SAMPLER_DEMC && cfg->kind != EPI_SAMPLER_DRJ)
    return epi_fail(h, EPI_ERR_ARG, "unknown sampler kind");
  // the whole configuration is checked before anything is allocated or installed
  const int n_rungs = cfg->n_rungs > 1 ? cfg->n_rungs : 1;
  if (cfg->n_chains % n_rungs) return epi_fail(h, EPI_ERR_ARG, "n_chains must be a multiple of n_rungs");
  if (cfg->kind == EPI_SAMPLER_DEMC && cfg->n_chains / n_rungs < 3)
    return epi_fail(h, EPI_ERR_ARG, "DE-MC needs at least 3 chains per rung");
  CUDA_OK(h, cudaSetDevice(h->device));
  epi_sampler_free(h);
  SamplerState *s = new SamplerState();
  h->sampler = s;
  s->cfg = *cfg;
  if (!(s->cfg.beta > 0)) s->cfg.beta = 1.0;
  s->replicates = (cfg->reserved & EPI_SAMPLER_FLAG_REPLICATES) != 0;
  const int d = h->M.d;
  const int64_t n = cfg->n_chains;
  s->ld = (n + kTile - 1) / kTile * kTile;
  s->n_rungs = n_rungs;
  s->cpr = cfg->n_chains / n_rungs;
  const size_t mat = sizeof(double) * (size_t)(d * s->ld), vec = sizeof(double) * (size_t)s->ld;
  INIT_OK(h, cudaMalloc(&s->cur, mat));
  INIT_OK(h, cudaMalloc(&s->prop, mat));
  INIT_OK(h, cudaMalloc(&s->logl, vec));
  INIT_OK(h, cudaMalloc(&s->logp, vec));
  INIT_OK(h, cudaMalloc(&s->logl_p, vec));
  INIT_OK(h, cudaMalloc(&s->logp_p, vec));
  INIT_OK(h, cudaMalloc(&s->status_p, sizeof(int) * (size_t)s->ld));
  INIT_OK(h, cudaMalloc(&s->lscale, vec));
  INIT_OK(h, cudaMalloc(&s->acc, sizeof(unsigned long long)));
  INIT_OK(h, cudaMalloc(&s->psd, sizeof(double) * (size_t)d));
  INIT_OK(h, cudaMalloc(&s->cov, sizeof(double) * (size_t)s->n_rungs * d * d));
  INIT_OK(h, cudaMalloc(&s->chol, sizeof(double) * (size_t)s->n_rungs * d * d));
  INIT_OK(h, cudaMalloc(&s->chol_tmp, sizeof(double) * (size_t)s->n_rungs * d * d));
  INIT_OK(h, cudaMalloc(&s->shape_ok, sizeof(int) * (size_t)s->n_rungs));
  INIT_OK(h, cudaMalloc(&s->shape_ready, sizeof(int)));
  INIT_OK(h, cudaMemsetAsync(s->shape_ready, 0, sizeof(int), h->stream));
  INIT_OK(h, cudaMalloc(&s->swaps, 2 * sizeof(unsigned long long)));
  INIT_OK(h, cudaMemsetAsync(s->swaps, 0, 2 * sizeof(unsigned long long), h->stream));
  {
    std::vector<double> br(s->n_rungs, 1.0);
    const double pw = cfg->gti_pow > 0 ? cfg->gti_pow : 1.0;
    for (int r = 0; r < s->n_rungs; ++r)
      br[r] = (s->n_rungs > 1 && !s->replicates) ? std::pow((double)r / (double)(s->n_rungs - 1), pw) : 1.0;
    INIT_OK(h, cudaMalloc(&s->beta_rung, sizeof(double) * (size_t)s->n_rungs));
    INIT_OK(h, cudaMemcpyAsync(s->beta_rung, br.data(), sizeof(double) * br.size(), cudaMemcpyHostToDevice, h->stream));
    INIT_OK(h, cudaStreamSynchronize(h->stream));
  }
  INIT_OK(h, cudaMemsetAsync(s->lscale, 0, vec, h->stream));
  INIT_OK(h, cudaMemsetAsync(s->acc, 0, sizeof(unsigned long long), h->stream));
  INIT_OK(h, cudaMemsetAsync(s->cur, 0, mat, h->stream));
  INIT_OK(h, cudaMemsetAsync(s->prop, 0, mat, h->stream));
  int rc = epi_ensure_workspace(h, h->ws, h->M, n, false);
  if (rc) { epi_sampler_free(h); return rc; }
  // initial state: theta0 (n x d AoS) or the model file's init with a small Philox jitter
  std::vector<double> host((size_t)d * s->ld, 0.0);
  for (int64_t i = 0; i < n; ++i) {
    for (int p = 0; p < d; ++p) {
      double v;
      if (theta0) {
        v = theta0[i * d + p];
      } else {
        uint32_t r[4];
        const uint64_t cid = (uint64_t)(cfg->chain_id0 + i);
        philox4x32_10((uint32_t)cid, (uint32_t)(cid >> 32), 0xFFFFFFFFu, (uint32_t)p, (uint32_t)cfg->seed,
                      (uint32_t)(cfg->seed >> 32), r);
        const double w = h->box_max[p] - h->box_min[p];
        v = h->init[p] + 0.02 * w * (u01(r[0]) - 0.5);
        v = std::fmin(std::fmax(v, h->box_min[p]), h->box_max[p]);
      }
      host[(size_t)p * s->ld + i] = v;
    }
  }
  INIT_OK(h, cudaMemcpyAsync(s->cur, host.data(), mat, cudaMemcpyHostToDevice, h->stream));
  INIT_OK(h, cudaStreamSynchronize(h->stream));
  // log posterior of the initial state
  rc = epi_launch_eval_ws(h, h->ws, h->M, s->cur, s->ld, n, 0, s->logl, s->logp, s->status_p, nullptr, nullptr, h->stream);
  if (rc) { epi_sampler_free(h); return rc; }
  s->evals += n;
  if (cfg->kind == EPI_SAMPLER_DEMC) {  // default population: a snapshot of this rank's own chains
    INIT_OK(h, cudaMalloc(&s->pop_own, mat));
    INIT_OK(h, cudaMemcpyAsync(s->pop_own, s->cur, mat, cudaMemcpyDeviceToDevice, h->stream));
    s->pop = s->pop_own; s->pop_ranks = 1; s->pop_per_rank = (int)n; s->pop_ld = s->ld;
    if ((rc = install_population(h, s))) { epi_sampler_free(h); return rc; }
  }
  s->iter = 0;
  if (cfg->kind == EPI_SAMPLER_DRJ) {
    for (int p = 0; p < d; ++p)
      if (h->box_max[p] > h->box_min[p]) s->free_params.push_back(p);
    if (s->free_params.empty()) { epi_sampler_free(h); return epi_fail(h, EPI_ERR_ARG, "no free parameter"); }
    INIT_OK(h, cudaMalloc(&s->bw, mat));
    k_fill_const<<<(unsigned)(((int64_t)d * s->ld + 255) / 256), 256, 0, h->stream>>>((int64_t)d * s->ld, 1.0, s->bw);  // drjacoby starts every bandwidth at 1
    h->launches += 1;
    rc = launch_drj(h, s, /*sub_acc=*/0, /*p_acc=*/-1, s->free_params[0], /*full_sync=*/1, nullptr, nullptr);
  } else {
    rc = launch_ap(h, s, /*do_accept=*/0, nullptr, nullptr);  // proposal of iteration 1
  }
  if (rc) { epi_sampler_free(h); return rc; }
  INIT_OK(h, cudaStreamSynchronize(h->stream));
  return EPI_OK;
}

// EPI_SAMPLER_DRJ: n_iter sweeps over the free parameters, d' evaluations each
static int run_drj(epi_handle *h, SamplerState *s, int32_t n_iter) {
  const int64_t n = s->cfg.n_chains;
  const int nf = (int)s->free_params.size();
  const bool coupled = s->n_rungs > 1 && !s->replicates;
  for (int it = 0; it < n_iter; ++it) {
    for (int j = 0; j < nf; ++j) {
      int rc = epi_launch_eval_ws(h, h->ws, h->M, s->prop, s->ld, n, 0, s->logl_p, s->logp_p, s->status_p, nullptr, nullptr,
                                  h->stream);
      if (rc) return rc;
      s->evals += n;
      s->sub += 1;
      if (j + 1 < nf) {
        if ((rc = launch_drj(h, s, s->sub, s->free_params[j], s->free_params[j + 1], 0, nullptr, nullptr))) return rc;
        continue;
      }
      // end of the sweep: record, couple adjacent rungs, start the next sweep
      s->iter += 1;
      double *dout = nullptr, *dlp = nullptr;
      if (s->draws && s->thin > 0 && (s->iter % s->thin) == 0 && s->kept < s->capacity) {
        dout = s->draws + (size_t)s->kept * h->M.d * s->ld;
        dlp = s->draws_lp + (size_t)s->kept * s->ld;
        s->kept += 1;
      }
      if (!coupled) {
        if ((rc = launch_drj(h, s, s->sub, s->free_params[j], s->free_params[0], 0, dout, dlp))) return rc;
      } else {
        if ((rc = launch_drj(h, s, s->sub, s->free_params[j], -1, 0, nullptr, nullptr))) return rc;
        const int parity = (int)(s->iter & 1);
        const int64_t pairs = (int64_t)((s->n_rungs - parity) / 2) * s->cpr;
        if (pairs > 0) {
          k_pt_swap<<<(unsigned)((pairs + 127) / 128), 128, 0, h->stream>>>(h->M.d, s->cpr, s->n_rungs, s->ld, parity,
                                                                           s->cfg.seed, s->cfg.chain_id0, s->sub, s->beta_rung,
                                                                           s->cur, s->logl, s->logp, s->swaps);
          h->launches += 1;
        }
        if ((rc = launch_drj(h, s, s->sub, -1, s->free_params[0], 1, dout, dlp))) return rc;
      }
      if (s->adapt) s->adapt_sweeps += 1;
    }
  }
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return EPI_OK;
}

extern "C" int epi_sampler_run(epi_handle *h, int32_t n_iter) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  SamplerState *s = h->sampler;
  CUDA_OK(h, cudaSetDevice(h->device));
  const int64_t n = s->cfg.n_chains;
  if (s->cfg.kind == EPI_SAMPLER_DRJ) return run_drj(h, s, n_iter);
  if (s->adapt && s->cpr >= 2 * h->M.d) {
    // burn-in only: re-measure the population covariance that shapes the proposals.  (It is
    // frozen when adaptation is switched off, so the sampling phase is a fixed-kernel Markov chain.)
    int rc = update_shape(h, s);
    if (rc) return rc;
  }
  // A small population without tempering runs as `ways` slices, each a free-running pipeline on its own stream for
  // the whole call: nothing couples two chains inside a run (DE-MC partners come from the fixed snapshot, the
  // proposal shape is fixed), the ODE passes of <= 16 k chains are latency bound (0.43 ms whatever the batch), so
  // the slices drift apart and the score kernel of one fills the issue slots the ODE warps of the others leave
  // idle.  Same draws as one slice: the Philox counters carry (chain id, iteration).  (Slices launched together
  // and joined after every evaluation stay in phase and gain nothing: profiles/r2_summary.md.)
  int ways = h->split_ways;
  if (ways == 0) ways = n <= 3072 ? 1 : n <= 24576 ? 4 : n <= 40960 ? 2 : 1;
  if (s->n_rungs > 1 || h->timing) ways = 1;
  ways = (int)std::min<int64_t>(ways, std::max<int64_t>(1, n / 512));
  if (ways > 1) {
    const int64_t slice = (((n + ways - 1) / ways + 127) / 128) * 128;
    cudaEventRecord(h->ev_fork, h->stream);
    int used = 0;
    for (int k = 1; k < ways && (int64_t)k * slice < n; ++k) { cudaStreamWaitEvent(h->split_stream[k - 1], h->ev_fork, 0); used = k; }
    for (int it = 0; it < n_iter; ++it) {
      double *dout = nullptr, *dlp = nullptr;
      const int64_t iter = s->iter + 1;
      if (s->draws && s->thin > 0 && (iter % s->thin) == 0 && s->kept < s->capacity) {
        dout = s->draws + (size_t)s->kept * h->M.d * s->ld;
        dlp = s->draws_lp + (size_t)s->kept * s->ld;
        s->kept += 1;
      }
      for (int k = 0; k <= used; ++k) {
        const int64_t first = k * slice, cnt = std::min(slice, n - first);
        cudaStream_t sk = k == 0 ? h->stream : h->split_stream[k - 1];
        int rc = epi_launch_eval_range(h, h->ws, h->M, s->prop, s->ld, first, cnt, s->logl_p, s->logp_p, s->status_p, sk);
        if (rc) return rc;
        if ((rc = launch_ap(h, s, 1, dout, dlp, first, cnt, iter, sk))) return rc;
      }
      s->evals += n;
      s->iter = iter;
    }
    for (int k = 1; k <= used; ++k) {
      cudaEventRecord(h->ev_join[k - 1], h->split_stream[k - 1]);
      cudaStreamWaitEvent(h->stream, h->ev_join[k - 1], 0);
    }
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return EPI_OK;
  }
  for (int it = 0; it < n_iter; ++it) {
    int rc = epi_launch_eval_ws(h, h->ws, h->M, s->prop, s->ld, n, 0, s->logl_p, s->logp_p, s->status_p, nullptr, nullptr,
                                h->stream);
    if (rc) return rc;
    s->evals += n;
    s->iter += 1;
    double *dout = nullptr, *dlp = nullptr;
    if (s->draws && s->thin > 0 && (s->iter % s->thin) == 0 && s->kept < s->capacity) {
      dout = s->draws + (size_t)s->kept * h->M.d * s->ld;
      dlp = s->draws_lp + (size_t)s->kept * s->ld;
      s->kept += 1;
    }
    if ((rc = launch_ap(h, s, 1, dout, dlp))) return rc;
    if (s->n_rungs > 1 && !s->replicates) {
      const int parity = (int)(s->iter & 1);
      const int64_t pairs = (int64_t)((s->n_rungs - parity) / 2) * s->cpr;
      if (pairs > 0) {
        k_pt_swap<<<(unsigned)((pairs + 127) / 128), 128, 0, h->stream>>>(h->M.d, s->cpr, s->n_rungs, s->ld, parity,
                                                                         s->cfg.seed, s->cfg.chain_id0, s->iter, s->beta_rung,
                                                                         s->cur, s->logl, s->logp, s->swaps);
        h->launches += 1;
        // the proposal of the next iteration was drawn around the pre-swap state: redraw it
        if ((rc = launch_ap(h, s, 0, nullptr, nullptr))) return rc;
      }
    }
  }
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return EPI_OK;
}

extern "C" int epi_sampler_set_scale(epi_handle *h, double scale) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  if (!(scale > 0) || !std::isfinite(scale)) return epi_fail(h, EPI_ERR_ARG, "scale must be positive and finite");
  SamplerState *s = h->sampler;
  CUDA_OK(h, cudaSetDevice(h->device));
  if (s->cfg.kind == EPI_SAMPLER_DRJ) return epi_fail(h, EPI_ERR_STATE, "the drjacoby-style sampler has per-parameter bandwidths, no global scale");
  s->cfg.scale = scale;
  // the pending proposal (iteration iter + 1) was drawn with the old scale: redraw it (same Philox counters)
  int rc = launch_ap(h, s, 0, nullptr, nullptr);
  if (rc) return rc;
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return EPI_OK;
}

extern "C" int epi_sampler_pt_stats(epi_handle *h, int64_t *proposed, int64_t *accepted) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  unsigned long long c[2] = {0, 0};
  CUDA_OK(h, cudaSetDevice(h->device));
  CUDA_OK(h, cudaMemcpyAsync(c, h->sampler->swaps, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  if (proposed) *proposed = (int64_t)c[0];
  if (accepted) *accepted = (int64_t)c[1];
  return EPI_OK;
}

extern "C" int epi_sampler_adapt(epi_handle *h, int enable, double target_accept) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  h->sampler->adapt = enable != 0;
  h->sampler->adapt_target = (target_accept > 0 && target_accept < 1) ? target_accept : 0.0;
  return EPI_OK;
}

extern "C" int epi_sampler_state_device(epi_handle *h, double **d_theta_soa, double **d_logl, double **d_logp, int64_t *ld) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  if (d_theta_soa) *d_theta_soa = h->sampler->cur;
  if (d_logl) *d_logl = h->sampler->logl;
  if (d_logp) *d_logp = h->sampler->logp;
  if (ld) *ld = h->sampler->ld;
  return EPI_OK;
}

extern "C" int epi_sampler_set_population(epi_handle *h, const double *d_pop, int32_t n_ranks, int32_t chains_per_rank,
                                          int64_t ld_rank) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  SamplerState *s = h->sampler;
  if (!d_pop) {  // refresh the built-in single-rank snapshot from the current state
    if (!s->pop_own) return epi_fail(h, EPI_ERR_STATE, "no own population (sampler kind is not DE-MC)");
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaMemcpyAsync(s->pop_own, s->cur, sizeof(double) * (size_t)(h->M.d * s->ld), cudaMemcpyDeviceToDevice, h->stream));
    s->pop = s->pop_own; s->pop_ranks = 1; s->pop_per_rank = s->cfg.n_chains; s->pop_ld = s->ld;
    return install_population(h, s);
  }
  if (n_ranks < 1 || chains_per_rank < 1 || (int64_t)n_ranks * chains_per_rank < 3 || ld_rank < chains_per_rank)
    return epi_fail(h, EPI_ERR_ARG, "bad population shape");
  s->pop = d_pop; s->pop_ranks = n_ranks; s->pop_per_rank = chains_per_rank; s->pop_ld = ld_rank;
  CUDA_OK(h, cudaSetDevice(h->device));
  return install_population(h, s);
}

extern "C" int epi_sampler_get(epi_handle *h, double *theta, double *logl, double *logp) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  SamplerState *s = h->sampler;
  CUDA_OK(h, cudaSetDevice(h->device));
  const int d = h->M.d;
  const int64_t n = s->cfg.n_chains;
  if (theta) {
    std::vector<double> host((size_t)d * s->ld);
    CUDA_OK(h, cudaMemcpyAsync(host.data(), s->cur, sizeof(double) * host.size(), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < n; ++i)
      for (int p = 0; p < d; ++p) theta[i * d + p] = host[(size_t)p * s->ld + i];
  }
  if (logl) CUDA_OK(h, cudaMemcpyAsync(logl, s->logl, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  if (logp) CUDA_OK(h, cudaMemcpyAsync(logp, s->logp, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return EPI_OK;
}

extern "C" int epi_sampler_stats(epi_handle *h, int64_t *n_iter, int64_t *n_accept, int64_t *n_evals) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  SamplerState *s = h->sampler;
  CUDA_OK(h, cudaSetDevice(h->device));
  unsigned long long acc = 0;
  CUDA_OK(h, cudaMemcpyAsync(&acc, s->acc, sizeof acc, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  if (n_iter) *n_iter = s->iter;
  if (n_accept) *n_accept = (int64_t)acc;
  if (n_evals) *n_evals = s->evals;
  return EPI_OK;
}

extern "C" int epi_sampler_record(epi_handle *h, int32_t thin, int32_t capacity) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  SamplerState *s = h->sampler;
  CUDA_OK(h, cudaSetDevice(h->device));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  cudaFree(s->draws); cudaFree(s->draws_lp);
  s->draws = s->draws_lp = nullptr;
  s->thin = thin; s->capacity = capacity; s->kept = 0;
  if (thin > 0 && capacity > 0) {
    CUDA_OK(h, cudaMalloc(&s->draws, sizeof(double) * (size_t)capacity * h->M.d * s->ld));
    CUDA_OK(h, cudaMalloc(&s->draws_lp, sizeof(double) * (size_t)capacity * s->ld));
  }
  return EPI_OK;
}

extern "C" int epi_sampler_draws(epi_handle *h, double *draws, double *logpost, int32_t *n_kept) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  SamplerState *s = h->sampler;
  CUDA_OK(h, cudaSetDevice(h->device));
  const int d = h->M.d;
  const int64_t n = s->cfg.n_chains;
  if (n_kept) *n_kept = s->kept;
  if (draws && s->kept > 0) {  // device [k][p][ld] -> host [k][chain][p]
    std::vector<double> host((size_t)d * s->ld);
    for (int k = 0; k < s->kept; ++k) {
      CUDA_OK(h, cudaMemcpyAsync(host.data(), s->draws + (size_t)k * d * s->ld, sizeof(double) * host.size(),
                                 cudaMemcpyDeviceToHost, h->stream));
      CUDA_OK(h, cudaStreamSynchronize(h->stream));
      for (int64_t i = 0; i < n; ++i)
        for (int p = 0; p < d; ++p) draws[((size_t)k * n + i) * d + p] = host[(size_t)p * s->ld + i];
    }
  }
  if (logpost && s->kept > 0) {
    CUDA_OK(h, cudaMemcpy2DAsync(logpost, sizeof(double) * (size_t)n, s->draws_lp, sizeof(double) * (size_t)s->ld,
                                 sizeof(double) * (size_t)n, s->kept, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
  }
  return EPI_OK;
}

extern "C" int epi_sampler_draws_subset(epi_handle *h, int32_t first, int32_t n_sub, double *draws, double *logpost,
                                        int32_t *n_kept) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  SamplerState *s = h->sampler;
  if (first < 0 || n_sub < 1 || first + n_sub > s->cfg.n_chains) return epi_fail(h, EPI_ERR_ARG, "chain range out of bounds");
  CUDA_OK(h, cudaSetDevice(h->device));
  if (n_kept) *n_kept = s->kept;
  if (s->kept == 0) return EPI_OK;
  const int d = h->M.d;
  if (draws)  // device [k][p][ld] -> host [k][p][n_sub]: one strided copy
    CUDA_OK(h, cudaMemcpy2DAsync(draws, sizeof(double) * (size_t)n_sub, s->draws + first, sizeof(double) * (size_t)s->ld,
                                 sizeof(double) * (size_t)n_sub, (size_t)s->kept * d, cudaMemcpyDeviceToHost, h->stream));
  if (logpost)
    CUDA_OK(h, cudaMemcpy2DAsync(logpost, sizeof(double) * (size_t)n_sub, s->draws_lp + first, sizeof(double) * (size_t)s->ld,
                                 sizeof(double) * (size_t)n_sub, (size_t)s->kept, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return EPI_OK;
}

extern "C" int epi_sampler_draws_device(epi_handle *h, double **d_draws, double **d_logpost, int32_t *n_kept, int64_t *ld) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  SamplerState *s = h->sampler;
  if (d_draws) *d_draws = s->draws;
  if (d_logpost) *d_logpost = s->draws_lp;
  if (n_kept) *n_kept = s->kept;
  if (ld) *ld = s->ld;
  return EPI_OK;
}

extern "C" int epi_sampler_bandwidths(epi_handle *h, double *bw) {
  if (!h || !h->sampler) return epi_fail(h, EPI_ERR_STATE, "sampler not initialised");
  SamplerState *s = h->sampler;
  if (!s->bw || !bw) return epi_fail(h, EPI_ERR_STATE, "no bandwidths (sampler kind is not EPI_SAMPLER_DRJ)");
  CUDA_OK(h, cudaSetDevice(h->device));
  const int d = h->M.d;
  const int64_t n = s->cfg.n_chains;
  std::vector<double> host((size_t)d * s->ld);
  CUDA_OK(h, cudaMemcpyAsync(host.data(), s->bw, sizeof(double) * host.size(), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  for (int64_t i = 0; i < n; ++i)
    for (int p = 0; p < d; ++p) bw[i * d + p] = host[(size_t)p * s->ld + i];
  return EPI_OK;
}

// K2 alone on a synthetic population (see include/epimcmc.h)
extern "C" int epi_k2_bench(epi_handle *h, int32_t kind, int64_t n_chains, int32_t iters, double *us_per_iter,
                            double *bytes_per_iter, double *accept_rate) {
  if (!h || n_chains < 128 || iters < 1) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  if (kind != EPI_SAMPLER_DEMC && kind != EPI_SAMPLER_DRJ) return epi_fail(h, EPI_ERR_ARG, "epi_k2_bench: DE-MC or DRJ");
  CUDA_OK(h, cudaSetDevice(h->device));
  const int d = h->M.d;
  const int64_t n = n_chains, ld = (n + kTile - 1) / kTile * kTile;
  const size_t mat = sizeof(double) * (size_t)d * ld, vec = sizeof(double) * (size_t)ld;
  double *cur = nullptr, *prop = nullptr, *pop = nullptr, *v[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *psd = nullptr, *bw = nullptr;
  int *st = nullptr;
  unsigned long long *acc = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = EPI_OK;
  auto cleanup = [&]() {
    cudaFree(cur); cudaFree(prop); cudaFree(pop); for (double *q : v) cudaFree(q); cudaFree(psd); cudaFree(bw); cudaFree(st); cudaFree(acc);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define K2_OK(call)                                                                                      \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) { cleanup(); return epi_fail(h, EPI_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } \
  } while (0)
  K2_OK(cudaMalloc(&cur, mat));
  K2_OK(cudaMalloc(&prop, mat));
  if (kind == EPI_SAMPLER_DEMC) K2_OK(cudaMalloc(&pop, mat));
  if (kind == EPI_SAMPLER_DRJ) K2_OK(cudaMalloc(&bw, mat));
  for (double *&q : v) K2_OK(cudaMalloc(&q, vec));  // logl, logp, logl', logp', lscale
  K2_OK(cudaMalloc(&st, sizeof(int) * (size_t)ld));
  K2_OK(cudaMalloc(&acc, sizeof(unsigned long long)));
  K2_OK(cudaMalloc(&psd, sizeof(double) * d));
  K2_OK(cudaMemsetAsync(st, 0, sizeof(int) * (size_t)ld, h->stream));
  K2_OK(cudaMemsetAsync(acc, 0, sizeof(unsigned long long), h->stream));
  const unsigned gv = (unsigned)((ld + 255) / 256);
  k_fill_uniform<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(d, n, ld, 99, h->M.box_min, h->M.box_max, cur);
  K2_OK(cudaMemcpyAsync(prop, cur, mat, cudaMemcpyDeviceToDevice, h->stream));
  if (pop) k_pop_to_aos<<<dim3((unsigned)((n + 31) / 32), 1), 256, 0, h->stream>>>(cur, (int)n, ld, d, pop);  // member-major snapshot
  if (bw) k_fill_const<<<(unsigned)(((int64_t)d * ld + 255) / 256), 256, 0, h->stream>>>((int64_t)d * ld, 0.05, bw);
  k_fill_const<<<gv, 256, 0, h->stream>>>(ld, 0.0, v[0]);
  k_fill_const<<<gv, 256, 0, h->stream>>>(ld, 0.0, v[1]);
  k_fill_accept_pattern<<<gv, 256, 0, h->stream>>>(ld, 7, 0.25, v[2]);  // a quarter of the chains accepts, every iteration
  k_fill_const<<<gv, 256, 0, h->stream>>>(ld, 0.0, v[3]);
  k_fill_const<<<gv, 256, 0, h->stream>>>(ld, 0.0, v[4]);
  {
    std::vector<double> sd(d);
    for (int p = 0; p < d; ++p) sd[p] = 0.1 * (h->box_max[p] - h->box_min[p]);
    K2_OK(cudaMemcpyAsync(psd, sd.data(), sizeof(double) * d, cudaMemcpyHostToDevice, h->stream));
    K2_OK(cudaStreamSynchronize(h->stream));
  }
  std::vector<int> freep;
  for (int p = 0; p < d; ++p)
    if (h->box_max[p] > h->box_min[p]) freep.push_back(p);
  const int nf = (int)freep.size();
  auto one = [&](int64_t it, int do_accept) {
    if (kind == EPI_SAMPLER_DEMC)
      k_accept_propose<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(
          d, n, ld, EPI_SAMPLER_DEMC, 42, 0, it, do_accept, 1.0, nullptr, (int)n, 1.0, 1e-4, 0, h->M.box_min, h->M.box_max, cur,
          prop, v[0], v[1], v[2], v[3], st, v[4], 0, 0.0, acc, pop, 1, (int)n, ld, psd, nullptr, nullptr, nullptr);
    else
      k_drj_step<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(
          d, n, ld, 42, 0, it, do_accept ? freep[(it + nf - 1) % nf] : -1, freep[it % nf], do_accept ? 0 : 1, nullptr, (int)n,
          h->M.box_min, h->M.box_max, cur, prop, v[0], v[1], v[2], v[3], st, bw, 0, 0.0, 0.44, acc, nullptr, nullptr);
    h->launches += 1;
  };
  one(0, 0);
  for (int it = 1; it <= 3; ++it) one(it, 1);  // warm-up
  K2_OK(cudaMemsetAsync(acc, 0, sizeof(unsigned long long), h->stream));
  K2_OK(cudaEventCreate(&e0));
  K2_OK(cudaEventCreate(&e1));
  K2_OK(cudaEventRecord(e0, h->stream));
  for (int it = 0; it < iters; ++it) one(4 + it, 1);
  K2_OK(cudaEventRecord(e1, h->stream));
  K2_OK(cudaEventSynchronize(e1));
  float ms = 0.f;
  K2_OK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long a = 0;
  K2_OK(cudaMemcpy(&a, acc, sizeof a, cudaMemcpyDeviceToHost));
  const double ar = (double)a / ((double)n * iters);
  if (us_per_iter) *us_per_iter = 1e3 * ms / iters;
  if (accept_rate) *accept_rate = ar;
  if (bytes_per_iter) {
    if (kind == EPI_SAMPLER_DEMC) {
      // moving coordinates per proposal: every tenth proposal all d; otherwise CR in {0.1, 0.25, 0.5} or a single
      // coordinate, with equal odds
      const double dmove = 0.1 * d + 0.9 * 0.25 * ((0.1 + 0.25 + 0.5) * d + 1.0);
      *bytes_per_iter = (double)n * (8.0 * d /*theta read*/ + 8.0 * d /*proposal written*/ + 16.0 * dmove /*two partners*/ +
                                     ar * (16.0 * d + 16.0) /*accepted rows copied*/ + 36.0 /*log densities, status*/ + 8.0 /*step size*/);
    } else {
      *bytes_per_iter = (double)n * (16.0 /*theta_p, theta'_p*/ + 36.0 + 8.0 /*bandwidth*/ + ar * 24.0 + (1 - ar) * 8.0 + 16.0 /*next coordinate*/ + 8.0);
    }
  }
#undef K2_OK
  cleanup();
  return rc;
}

// test hook: raw Philox4x32-10 words from the device
extern "C" int epi_philox_probe(epi_handle *h, uint64_t seed, uint32_t iter, uint32_t use, int32_t n, uint32_t *out) {
  if (!h || !out || n < 1) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  CUDA_OK(h, cudaSetDevice(h->device));
  uint32_t *d = nullptr;
  CUDA_OK(h, cudaMalloc(&d, sizeof(uint32_t) * 4 * (size_t)n));
  k_philox_probe<<<(n + 127) / 128, 128, 0, h->stream>>>(seed, iter, use, n, d);
  h->launches += 1;
  CUDA_OK(h, cudaMemcpyAsync(out, d, sizeof(uint32_t) * 4 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  cudaFree(d);
  return EPI_OK;
}
