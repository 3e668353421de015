// epi_kernels.cu — the batched log-posterior evaluation kernels (K1), sm_100a.
//
// One evaluation of calclogl(transformParams(theta)) + calclogp(theta)
// (R/MCMC/fitMCMC-drj.R:19-21,51) is split into four launches with the thread
// mapping that suits each phase:
//
//   k_ode<FL,false>  thread per chain   pass-1 ODE, beta0 forever        (model-age-groups2q.R:160-165)
//   k_mid<FL>        warp per chain     latency profiles (pgamma/pnorm), died-threshold search,
//                                       data_offset + padding            (:106-114,167-191)
//   k_ode<FL,true>   thread per chain   pass-2 ODE with the beta(t) schedule (:193-213)
//   k_score<FL>      warp per chain     6 convolutions, rate map, NB scoring, ratio term, prior
//                                       (:229-331, 488-602, 380-486)
//
// The S(t) histories travel between the phases through HBM in a tiled layout
// [tile = chain/32][day][group][lane = chain%32]: the thread-per-chain writer
// stores 256 B per (day, group) fully coalesced, the warp-per-chain reader
// gathers its chain's days once into shared memory.
#include "epi_device.cuh"
#include "epi_launch.h"

namespace epi {

// S histories in HBM: hist[(chain * NG + group) * pitch + day], pitch = days rounded up to even so
// that every row is 16-byte aligned.  The thread-per-chain ODE writer scatters 8 B per (day, group)
// (L2 merges the four days of a 32-byte sector before write-back); the warp-per-chain readers then
// pull a whole row into shared memory with ONE 1-D bulk copy on the TMA unit (cp.async.bulk, SASS
// UBLKCP) instead of ~600 strided 8-byte loads per chain (profiles/r1_v5: history staging was 8 %
// of the score kernel's stall samples).
__host__ __device__ constexpr int hist_pitch(int ndays) { return (ndays + 1) & ~1; }

// The PASS-2 history (hist2, read by k_score / k_series) is stored in four PHASES per (chain, group):
//   day d  ->  row (d & 3) of length Q = hist_pitch2(ndays) / 4, element d >> 2.
// k_score's convolution gives every lane four CONSECUTIVE days and slides a register window of S over the delay
// (one new S value per delay and lane instead of one per day and delay): the lanes of a warp then read S at a
// stride of four days, which in a day-major row would be an eight-way shared-memory bank conflict and in the phase
// rows is one conflict-free load (round 1, profiles/r1_v15_ncu_full.txt: 192 M shared wavefronts per launch, 4 LDS
// per 6 DFMA, the FP64 pipe 30 % busy).  The thread-per-chain writer scatters 8 bytes per day either way.
__host__ __device__ constexpr int hist_pitch2(int ndays) { return (ndays + 7) & ~7; }
// index of day d in a history row: phase layout for pass 2 (Q = hist_pitch2 / 4), day-major for pass 1
template <bool PASS2>
__device__ __forceinline__ int hist_index(int d, int Q) { return PASS2 ? (d & 3) * Q + (d >> 2) : d; }

// Pass-2 restart.  Both passes use the same fixed-step scheme and pass 2 has beta0 until its first
// breakpoint, so its first D = floor(min breakpoint) - t0 days are BIT-identical to pass 1
// (SURVEY.md §8d allows the restart provided the reduced step count is reported: bench.py does).
// Pass 1 therefore checkpoints the full state at every day boundary — tiled so that a warp's store
// of one state is one contiguous 256 B line, ckpt[((tile*cdays + d)*NEQ + i)*32 + lane]; a
// chain-major layout made pass 1 LSU bound, 0.24 -> 1.14 ms — and pass 2 resumes from day D
// instead of re-integrating it.
__host__ __device__ constexpr int ckpt_days(int P1) { return P1 < 64 ? P1 : 64; }
constexpr int kPass1Chunk = 128;  // first-round days of a whole-period pass 1 (>= the checkpoint window)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(mbar)),
      "r"(parity)
      : "memory");
}

// internal status bit: pass 1 was integrated for its first chunk of days only and the threshold was not
// crossed there — the chain's pass 1 + threshold search are redone over the whole pass-1 period by the retry
// launches that follow in the same batch (never visible outside the library)
constexpr int kStRetry = 1 << 30;

template <int FL>
struct Traits {
  static constexpr bool k2i = (FL == EPI_FLAVOUR_AGE2I);  // model-age-groups2i.R + model-agef.cpp
  // model-age-groups2x.R + model-ager.cpp: waning immunity (R flows back into S, so R is ALWAYS integrated), the
  // younger group's effective-susceptible cap, 13 breakpoints, and the histories between the kernels hold the
  // infection INCIDENCE (derivs() output variables y.i / o.i) instead of S: the observation series are plain
  // convolutions of it, no N - x and no diff
  static constexpr bool k2x = (FL == EPI_FLAVOUR_AGE2X);
  static constexpr bool kAge = (FL == EPI_FLAVOUR_AGE2Q) || k2i || k2x;
  static constexpr int NEQ = kAge ? 10 : 4;
  static constexpr int NG = kAge ? 2 : 1;   // S series per chain
  static constexpr int NK = kAge ? 6 : 2;   // latency profiles per chain
  static constexpr int NBP = k2i ? 8 : k2x ? 13 : (kAge ? 10 : 2);  // schedule breakpoints
  // prefill slots kept in front of a staged S series == largest padding supported == rows of the
  // absolute-delay profile table: the 2i box lets the death latency reach 50 + 3*9 = 77 days
  // (model-age-groups2i.R:580-582), every other flavour stays below 64
  static constexpr int kPad = k2i ? EPI_MAX_PADDING : kPadMax;
};

// ------------------------------------------------------------------ theta map
// beta triple (y, o, yo) of schedule index k for the age flavours, from the host-built schedule table
// (R/models/model-age-groups2q.R:52-99, model-age-groups2i.R:52-93, model-age-groups2x.R:56-119).
template <int FL>
__device__ __forceinline__ void age_beta(const DevModel &M, const double *__restrict__ th, int64_t ld, int k, double &by,
                                         double &bo, double &byo) {
  static_assert(Traits<FL>::kAge, "age flavours only");
  auto T = [&](int i) { return th[(int64_t)i * ld]; };
  const SchedBeta b = M.sched->beta[k];
  by = T(b.i[0]); bo = T(b.i[1]); byo = T(b.i[2]);
  if (b.m[0] >= 0) by = T(b.m[0]) * by;
  if (b.m[1] >= 0) bo = T(b.m[1]) * bo;
  if (b.m[2] >= 0) byo = T(b.m[2]) * byo;
}

// segment kinds: 1 = linear towards k+1, 0 = hold.  model-agen.cpp:78-99: k = 0,1,2,3,8 linear;
// model-agef.cpp:66-87: k = 0..5 linear, 6 and 7 hold; model-ager.cpp:88-124: k = 0,1,2,3,8,10,11 linear
__device__ __forceinline__ bool age_linear(const DevModel &M, int k) { return (M.sched->linear >> k) & 1u; }

// ---------------------------------------------------------------------- K_ode
// Fixed-step classical RK4, m sub-steps per day, each sub-step cut again at the
// schedule breakpoints that fall strictly inside it, the beta(t) segment frozen
// per piece — the scheme of oracle/epi_oracle.c:ode_rk4.
//
// ncu (profiles/r1_v1_*): with the bounds guard of model-agen.cpp:109-124 written
// as 20 FP64 compares + "dy = 0" the compiler if-converts it into ~20 DSETP +
// ~40 SEL/CS2R/IMAD.MOV per stage: the FP64 pipe spends 1/3 of its cycles on
// compares and only 34 % of the issued instructions are FP64 math.  The guard
// almost never fires, so each stage only computes a CONSERVATIVE "maybe out of
// bounds" flag on the integer pipe (which is otherwise idle); when no stage of a
// step raises it the unguarded arithmetic is exactly what the guarded code would
// have produced, otherwise the step is redone by the exact guarded slow path.

// hh / 6 (the RK4 weight of a piece a breakpoint cut), correctly rounded without the division subroutine:
//   z = RN(1/6) = (1/6)(1 - 2^-54);  q = RN(x z) is RN(x/6) or its lower neighbour;  r = x - 6 q is exact (a small
//   multiple of 2 ulp(q));  q + r z = x/6 - (r/6) 2^-54 lies within 2^-54 ulp of x/6, and x/6 is never closer than
//   ulp/6 to a rounding boundary (3 does not divide a power of two)  =>  RN(q + r z) == RN(x / 6) for every normal x.
// (Markstein's division by a constant; tests/test_oracle.py checks it against the division on 10^6 values.)  The
// division's slow-path branch kept the compiler from overlapping it with the stages of the piece it belongs to.
__device__ __forceinline__ double div6(double x) {
  constexpr double z = 1.0 / 6.0;
  const double q = __dmul_rn(x, z);
  const double r = fma(-6.0, q, x);
  return fma(r, z, q);
}

// current schedule segment: value at tk and slope per curve (slope 0 = hold)
struct Seg {
  double tk, b0, b1, b2, s0, s1, s2;  // age: betay/betao/betayo; simple: beta, Tinf, 1/Tinf cache
  bool fast_ok;                       // beta(t) >= 0 on the whole segment (premise of the fast guard)
  bool hold;                          // every slope is 0: the curves are constants (fma(dt, 0, b) == b exactly)
};

// y > N or y < 0 or NaN  =>  bit 63 of (bits(N) - bits(y)) | bits(y) is set (conservative: it is
// also set for -0.0, which the slow path then treats exactly)
// Fast-path out-of-bounds test on the integer pipe, per group of states (S first).  It may only
// err on the safe side (a raised flag sends the step to the exact guarded path):
//   * "x < 0": OR of the high words, sign bit (also raised by -0.0; the exact path sorts that out);
//   * "x > N" for the non-S states: max of their high words >= high word of N, i.e. x within
//     2^-20 of N or above (never true in practice: E, I, R stay far below N);
//   * "S > N" needs no test at all: S starts at N-1 (older group: exactly N, and "S > N" is
//     strict) and every stage derivative of S is -(beta >= 0)*(I >= 0)/N*(S >= 0) <= 0 whenever the
//     guard let that stage through, so fma(c, k, S) <= S by monotone rounding.  The premise
//     beta(t) >= 0 is checked once per schedule segment (Seg::fast_ok).
// ncu/bench (profiles/r1_summary.md): the previous 64-bit form cost 30 ALU-pipe instructions per
// stage and more than half of the ODE kernel time (2.07 ms -> 0.99 ms with the test removed).
// (As unsigned integers the high words of negative doubles are >= 0x80000000 > hi(N), so one
// unsigned max over the non-S states covers both "x < 0" and "x > N" for them; only S needs its
// own sign test.)
__device__ __forceinline__ unsigned oob_group5(double S, double a, double b, double c, double d, unsigned nhi_m1) {
  const unsigned mx = max(max((unsigned)__double2hiint(a), (unsigned)__double2hiint(b)),
                          max((unsigned)__double2hiint(c), (unsigned)__double2hiint(d)));
  return (nhi_m1 - mx) | (unsigned)__double2hiint(S);  // bit 31 set <=> maybe out of bounds
}
__device__ __forceinline__ unsigned oob_group4(double S, double a, double b, double c, unsigned nhi_m1) {
  const unsigned mx = max(max((unsigned)__double2hiint(a), (unsigned)__double2hiint(b)), (unsigned)__double2hiint(c));
  return (nhi_m1 - mx) | (unsigned)__double2hiint(S);
}

// The removed compartment R is not integrated on the hot path.  R only ever enters the reference through
// the bounds guard (R < 0 || R > N, model-agen.cpp:109-124): it never feeds back into S, E, I.
//   * R >= 0 always: R starts at 0 and every stage derivative the guard lets through is gamma * (I >= 0), so
//     fma(c, k, R) >= R by monotone rounding.
//   * R <= N whenever S >= 1 at that stage input and the other states pass the guard: the stage inputs conserve
//     the population, S + E(1,2) + I + R = N up to the accumulated rounding (< 1e-5 persons after 10^3 steps of
//     <= 5 half-ulp(N) errors each), so R <= N - 1 + 1e-5.
// The fast path therefore tests "S < 1" instead of "S < 0" (same cost class: integer ops on the high word) and
// sends such steps to the exact path; the exact path, which cannot evaluate the R test without R, raises
// `need_r` when a stage input has 0 <= S < 1 with the guard otherwise silent — the chain is then recomputed
// from scratch by the R-tracking instance of the kernel (launched right behind, exits at once otherwise).
// In other words: results are bit-identical to integrating R, and R is integrated whenever it could matter.
// bit 31 of the result set <=> S < 1.0 (incl. negative, -0.0, NaN)
__device__ __forceinline__ unsigned below_one(double S) {
  const unsigned v = (unsigned)__double2hiint(S) - 0x3ff00000u;  // S in [1, +inf) <=> v in [0, 0x40000000]
  return v | (0x40000000u - v);
}
__device__ __forceinline__ unsigned oob_group4_nor(double S, double a, double b, double c, unsigned nhi_m1) {
  const unsigned mx = max(max((unsigned)__double2hiint(a), (unsigned)__double2hiint(b)), (unsigned)__double2hiint(c));
  return (nhi_m1 - mx) | below_one(S);
}
__device__ __forceinline__ unsigned oob_group3_nor(double S, double a, double b, unsigned nhi_m1) {
  const unsigned mx = max((unsigned)__double2hiint(a), (unsigned)__double2hiint(b));
  return (nhi_m1 - mx) | below_one(S);
}

// TRACKR = false: y[4], y[9] (age) / y[3] (simple) are not integrated (see below_one); `flag` bit 0 of the
// GUARDED instance reports a stage where the missing R test could have mattered.
// HOLD = true: the segment's curves are constants (pass 1 and the step segments of the schedule, i.e. most of the
// integration): no t - tk and no fma per curve and stage; same bits, since fma(dt, 0, b) == b.
template <int FL, bool GUARDED, bool TRACKR, bool HOLD = false>
__device__ __forceinline__ void rhs(const DevModel &M, const Seg &sg, double t, const double *y, double *dy,
                                    double inv0, double inv1, unsigned nb0, unsigned nb1, unsigned &flag) {
  const double dt = HOLD ? 0.0 : t - sg.tk;
  if constexpr (Traits<FL>::kAge) {
    const double Sy = y[0], E1y = y[1], E2y = y[2], Iy = y[3], Ry = TRACKR ? y[4] : 0.0;
    const double So = y[5], E1o = y[6], E2o = y[7], Io = y[8], Ro = TRACKR ? y[9] : 0.0;
    if constexpr (GUARDED) {
      const double Ny = M.Ny, No = M.No;
      // bounds guard, model-agen.cpp:109-124: freeze the state
      if (Sy < 0 || E1y < 0 || E2y < 0 || Iy < 0 || Ry < 0 || Sy > Ny || E1y > Ny || E2y > Ny || Iy > Ny ||
          Ry > Ny || So < 0 || E1o < 0 || E2o < 0 || Io < 0 || Ro < 0 || So > No || E1o > No || E2o > No ||
          Io > No || Ro > No) {
#pragma unroll
        for (int i = 0; i < 10; ++i) dy[i] = 0.0;
        return;
      }
      if constexpr (!TRACKR) flag |= (Sy < 1.0 || So < 1.0) ? 1u : 0u;  // the R test is undecidable here
    } else if constexpr (TRACKR) {
      flag |= oob_group5(Sy, E1y, E2y, Iy, Ry, nb0) | oob_group5(So, E1o, E2o, Io, Ro, nb1);
      // with waning immunity S is no longer monotone: "S > N" gets its own (exact) test
      if constexpr (Traits<FL>::k2x) flag |= (Sy > M.Ny || So > M.No) ? 0x80000000u : 0u;
    } else {
      flag |= oob_group4_nor(Sy, E1y, E2y, Iy, nb0) | oob_group4_nor(So, E1o, E2o, Io, nb1);
    }
    // Every multiply-add below is an EXPLICIT fma and every other operation an explicit rounding intrinsic, so
    // that all instances of this function (fast / exact guard, with / without R) round identically: left to the
    // compiler, the contraction of a * b - c differed between instances and the R-tracking fallback was not
    // bit-identical to the hot path.  (a2 = 1/Tinc2 = 1, model-age-groups2q.R:10,14.)
    const double betay = HOLD ? sg.b0 : fma(dt, sg.s0, sg.b0);
    const double betao = HOLD ? sg.b1 : fma(dt, sg.s1, sg.b1);
    const double betayo = HOLD ? sg.b2 : fma(dt, sg.s2, sg.b2);
    // (betay, betao, betayo are beta_y / N_y, beta_o / N_o, beta_yo / N_y here, see select_segment)
    if constexpr (Traits<FL>::k2x) {
      // model-ager.cpp:176-200 (uncertain() multiplies by funcertain = 1, model-age-groups2x.R:221: exact identity)
      static_assert(TRACKR, "the waning-immunity flavour always integrates R");
      const double Syeff = fmax(0.0, __dsub_rn(__dmul_rn(M.Ny, 0.8), __dsub_rn(M.Ny, Sy)));
      const double yinf = __dmul_rn(__dmul_rn(betay, Iy), Syeff);
      const double oinf = __dmul_rn(fma(betao, Io, __dmul_rn(betayo, Iy)), So);
      const double yrev = __dmul_rn(M.eta, Ry), orev = __dmul_rn(M.eta, Ro);
      dy[0] = __dsub_rn(yrev, yinf);
      dy[1] = fma(-M.a1, E1y, yinf);
      dy[2] = fma(M.a1, E1y, -E2y);
      dy[3] = fma(-M.gamma, Iy, E2y);
      dy[4] = fma(M.gamma, Iy, -yrev);
      dy[5] = __dsub_rn(orev, oinf);
      dy[6] = fma(-M.a1, E1o, oinf);
      dy[7] = fma(M.a1, E1o, -E2o);
      dy[8] = fma(-M.gamma, Io, E2o);
      dy[9] = fma(M.gamma, Io, -orev);
      return;
    }
    const double yinf = __dmul_rn(__dmul_rn(betay, Iy), Sy);
    const double oinf = __dmul_rn(fma(betao, Io, __dmul_rn(betayo, Iy)), So);
    dy[0] = -yinf;
    dy[1] = fma(-M.a1, E1y, yinf);      // got_infected - a1 * E1
    dy[2] = fma(M.a1, E1y, -E2y);       // a1 * E1 - E2
    dy[3] = fma(-M.gamma, Iy, E2y);     // E2 - gamma * I
    dy[5] = -oinf;
    dy[6] = fma(-M.a1, E1o, oinf);
    dy[7] = fma(M.a1, E1o, -E2o);
    dy[8] = fma(-M.gamma, Io, E2o);
    if constexpr (TRACKR) { dy[4] = __dmul_rn(M.gamma, Iy); dy[9] = __dmul_rn(M.gamma, Io); }
  } else {
    const double S = y[0], E = y[1], I = y[2], R = TRACKR ? y[3] : 0.0;
    const double N = M.N;
    if constexpr (GUARDED) {
      if (S < 0 || E < 0 || I < 0 || R < 0 || S > N || E > N || I > N || R > N) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dy[i] = 0.0;
        return;
      }
      if constexpr (!TRACKR) flag |= (S < 1.0) ? 1u : 0u;
    } else if constexpr (TRACKR) {
      flag |= oob_group4(S, E, I, R, nb0);
    } else {
      flag |= oob_group3_nor(S, E, I, nb0);
    }
    const double beta = HOLD ? sg.b0 : fma(dt, sg.s0, sg.b0);
    const double rT = (HOLD || sg.s1 == 0.0) ? sg.b2 : __ddiv_rn(1.0, fma(dt, sg.s1, sg.b1));  // b2 caches 1/Tinf on hold segments
    double inf;
    if constexpr (FL == EPI_FLAVOUR_NE) {
      const double Seff = fmax(0.0, __dsub_rn(inv1, __dsub_rn(N, S)));             // inv1 carries Ne = Nef*N
      inf = __dmul_rn(__dmul_rn(beta, I), Seff);  // beta is beta / Ne here, see select_segment
    } else {
      inf = __dmul_rn(__dmul_rn(beta, I), S);     // beta is beta / N here
    }
    const double ir = __dmul_rn(I, rT);
    dy[0] = -inf;
    dy[1] = fma(-M.a_inc, E, inf);  // got_infected - a * E
    dy[2] = fma(M.a_inc, E, -ir);   // a * E - I / Tinf
    if constexpr (TRACKR) dy[3] = ir;
  }
}

// index i is a removed compartment (not integrated when TRACKR is false)
template <int FL>
__device__ __forceinline__ constexpr bool is_r_state(int i) { return Traits<FL>::kAge ? (i == 4 || i == 9) : (i == 3); }

template <int FL, bool GUARDED, bool TRACKR, bool HOLD = false>
__device__ __forceinline__ bool rk4_try(const DevModel &M, const Seg &sg, const double *y, double *ynew, double pa,
                                        double pb, double hh, double hh2, double h6, double inv0, double inv1,
                                        unsigned nb0, unsigned nb1, bool &need_r) {
  constexpr int NEQ = Traits<FL>::NEQ;
  const double tm = pa + hh2;
  double k[NEQ], acc[NEQ], yt[NEQ];
  unsigned flag = 0;
  rhs<FL, GUARDED, TRACKR, HOLD>(M, sg, pa, y, k, inv0, inv1, nb0, nb1, flag);
#pragma unroll
  for (int i = 0; i < NEQ; ++i) {
    if (!TRACKR && is_r_state<FL>(i)) continue;
    acc[i] = k[i]; yt[i] = fma(hh2, k[i], y[i]);
  }
  rhs<FL, GUARDED, TRACKR, HOLD>(M, sg, tm, yt, k, inv0, inv1, nb0, nb1, flag);
#pragma unroll
  for (int i = 0; i < NEQ; ++i) {
    if (!TRACKR && is_r_state<FL>(i)) continue;
    acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh2, k[i], y[i]);
  }
  rhs<FL, GUARDED, TRACKR, HOLD>(M, sg, tm, yt, k, inv0, inv1, nb0, nb1, flag);
#pragma unroll
  for (int i = 0; i < NEQ; ++i) {
    if (!TRACKR && is_r_state<FL>(i)) continue;
    acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh, k[i], y[i]);
  }
  rhs<FL, GUARDED, TRACKR, HOLD>(M, sg, pb, yt, k, inv0, inv1, nb0, nb1, flag);
#pragma unroll
  for (int i = 0; i < NEQ; ++i) {
    if (!TRACKR && is_r_state<FL>(i)) { ynew[i] = 0.0; continue; }
    ynew[i] = fma(h6, __dadd_rn(acc[i], k[i]), y[i]);
  }
  if constexpr (GUARDED) {
    if (!TRACKR && (flag & 1u)) need_r = true;
    return true;
  }
  return (int)flag >= 0;  // no stage raised the maybe-out-of-bounds bit
}

// one RK4 step from y to yn (distinct arrays).  hh = pb - pa, hh2 = hh / 2, h6 = hh / 6: constants of the kernel for a
// regular sub-step, computed by the caller for a piece that a breakpoint cut
template <int FL, bool TRACKR>
__device__ __forceinline__ void rk4_step(const DevModel &M, const Seg &sg, const double *y, double *yn, double pa,
                                         double pb, double hh, double hh2, double h6, double inv0, double inv1,
                                         unsigned nb0, unsigned nb1, bool &need_r) {
  bool ok = false;
  // the hold instance only saves the three beta(t) fmas (fma(dt, 0, b) == b exactly), so a warp whose lanes sit in
  // segments of both kinds — chains of a sampler cross their breakpoints on different days — takes the general
  // instance together instead of running the two one after the other
  if (sg.fast_ok) {
    const bool hold = __all_sync(__activemask(), sg.hold);
    ok = hold ? rk4_try<FL, false, TRACKR, true>(M, sg, y, yn, pa, pb, hh, hh2, h6, inv0, inv1, nb0, nb1, need_r)
                 : rk4_try<FL, false, TRACKR, false>(M, sg, y, yn, pa, pb, hh, hh2, h6, inv0, inv1, nb0, nb1, need_r);
  }
  if (!ok) {
    // rare: redo the step with the reference's guard evaluated exactly in FP64
    rk4_try<FL, true, TRACKR>(M, sg, y, yn, pa, pb, hh, hh2, h6, inv0, inv1, 0u, 0u, need_r);
  }
}

template <int FL, bool TRACKR>
__device__ __forceinline__ void rk4_piece(const DevModel &M, const Seg &sg, double *y, double pa, double pb,
                                          double hh, double hh2, double h6, double inv0, double inv1, unsigned nb0,
                                          unsigned nb1, bool &need_r) {
  constexpr int NEQ = Traits<FL>::NEQ;
  double yn[NEQ];
  rk4_step<FL, TRACKR>(M, sg, y, yn, pa, pb, hh, hh2, h6, inv0, inv1, nb0, nb1, need_r);
#pragma unroll
  for (int i = 0; i < NEQ; ++i) y[i] = yn[i];
}

// Breakpoints of pass 2 from theta + data_offset (model-age-groups2q.R:184-197;
// simple: model-simple-ode-drj-cmp.R:98-99).
template <int FL>
__device__ __forceinline__ int breakpoints(const DevModel &M, const double *__restrict__ th, int64_t ld,
                                           double data_offset, double *bp) {
  if constexpr (Traits<FL>::kAge) {
    // the schedule table of the model file (model-age-groups2q.R:184-197, 2i :169-183, 2x :206-219)
    const int n = M.sched->n;
    for (int k = 0; k < n; ++k) {
      const SchedBp e = M.sched->bp[k];
      double t = data_offset + e.c;
      if (e.ia >= 0) {
        const double a = th[(int64_t)e.ia * ld];
        t = t + (e.clamp ? fmax(a, 0.0) : a);
      }
      if (e.ib >= 0) t = t + th[(int64_t)e.ib * ld];
      bp[k] = t + e.post;
    }
    return n;
  } else {
    bp[0] = data_offset + M.lockdown_offset;
    bp[1] = data_offset + M.lockdown_offset + M.ltp;
    return 2;
  }
}

// (Re)select the schedule segment active on a piece that starts at pa: the
// highest k with t_k <= pa (== the highest k with tm > t_k for a piece without
// interior breakpoints), and the next cut.
template <int FL>
// The curves are stored already divided by the population they multiply (beta_y / N_y, beta_o / N_o,
// beta_yo / N_y; simple: beta / N or beta / Ne), so that the RHS needs no I / N products.
__device__ __noinline__ Seg select_segment(const DevModel &M, const double *__restrict__ th, int64_t ld,
                                           const double *bp, int nbp, double pa, double *tcut, double inv0,
                                           double inv1, int *kcur, bool sorted) {
  int k = -1;
  double tc = INFINITY;
  if (sorted && *kcur >= -1) {
    // breakpoints in index order are in time order (the usual case, checked once per chain) and the segment
    // before this cut is known: the scan below reduces to stepping over the breakpoints just passed
    k = *kcur;
    while (k + 1 < nbp && bp[k + 1] <= pa) ++k;
    if (k + 1 < nbp) tc = bp[k + 1];
  } else {
    for (int i = 0; i < nbp; ++i) {
      if (bp[i] <= pa) k = i;
      if (bp[i] > pa && bp[i] < tc) tc = bp[i];
    }
  }
  *kcur = k;
  *tcut = tc;
  Seg sg;
  if constexpr (Traits<FL>::kAge) {
    double by, bo, byo;
    age_beta<FL>(M, th, ld, k < 0 ? 0 : k, by, bo, byo);
    sg.b0 = by; sg.b1 = bo; sg.b2 = byo;
    sg.s0 = sg.s1 = sg.s2 = 0.0;
    sg.tk = 0.0;
    if (k >= 0 && age_linear(M, k)) {
      double ny, no, nyo;
      age_beta<FL>(M, th, ld, k + 1, ny, no, nyo);
      const double len = bp[k + 1] - bp[k];
      sg.tk = bp[k];
      sg.s0 = (ny - by) / len; sg.s1 = (no - bo) / len; sg.s2 = (nyo - byo) / len;
      sg.fast_ok = by >= 0 && bo >= 0 && byo >= 0 && ny >= 0 && no >= 0 && nyo >= 0 && len > 0;
    } else {
      sg.fast_ok = by >= 0 && bo >= 0 && byo >= 0;
    }
    sg.b0 = __dmul_rn(sg.b0, inv0); sg.s0 = __dmul_rn(sg.s0, inv0);
    sg.b1 = __dmul_rn(sg.b1, inv1); sg.s1 = __dmul_rn(sg.s1, inv1);
    sg.b2 = __dmul_rn(sg.b2, inv0); sg.s2 = __dmul_rn(sg.s2, inv0);
    sg.hold = sg.s0 == 0.0 && sg.s1 == 0.0 && sg.s2 == 0.0;
  } else {
    auto T = [&](int i) { return th[(int64_t)i * ld]; };
    const double Tinf0 = T(2), Tinft = T(3);
    const double beta0 = T(0) / Tinf0, betat = T(1) / Tinft;
    sg.tk = 0.0; sg.s0 = 0.0; sg.s1 = 0.0; sg.s2 = 0.0;
    if (k < 0) { sg.b0 = beta0; sg.b1 = Tinf0; sg.b2 = 1.0 / Tinf0; }
    else if (k == 1) { sg.b0 = betat; sg.b1 = Tinft; sg.b2 = 1.0 / Tinft; }
    else {
      const double len = bp[1] - bp[0];
      sg.tk = bp[0]; sg.b0 = beta0; sg.b1 = Tinf0; sg.b2 = 0.0;
      sg.s0 = (betat - beta0) / len; sg.s1 = (Tinft - Tinf0) / len;
      if (sg.s1 == 0.0) sg.b2 = 1.0 / Tinf0;
    }
    sg.fast_ok = beta0 >= 0 && betat >= 0 && bp[1] > bp[0];
    sg.b0 = __dmul_rn(sg.b0, inv0); sg.s0 = __dmul_rn(sg.s0, inv0);  // beta / N (Ne flavour: beta / Ne)
    sg.hold = sg.s0 == 0.0 && sg.s1 == 0.0;  // (then b2 = 1 / Tinf is set)
  }
  return sg;
}

// Last simulation day any data window of calclogl reads (window clamps of model-age-groups2q.R:497-512
// for every scored series): the scoring path needs S(t) only up to there, so pass 2 stops at that day
// when it feeds k_score (never when it feeds calculateModel trajectories).
__device__ __forceinline__ int last_window_day(const DevModel &M, int off_raw, int pad) {
  const int off = off_raw != EPI_INVALID_DATA_OFFSET ? off_raw : 1, T = pad + M.P;
  int thi = 1;
  for (int s = 0; s < M.n_series; ++s) {
    const int n = M.n_obs[s];
    int ds = off, de = off + n - 1;
    if (de > T) { de = T; ds = de - n + 1; }
    if (ds < 1) ds = 1;
    thi = max(thi, ds + n - 1);
  }
  return thi;
}

// ---- the state fields calculateModel derives from the ode() matrix besides S (epi_trajectories_ext)
// age (model-age-groups2q.R:216-227): y.E = E1y + E2y, y.I, y.R, o.E, o.I, o.R, Re, Rt, y.Re, o.Re;
// the 2i file (:199-210) assigns o.E = out[,8] (E1o only) and o.I = out[,9] + out[,10] (E2o + Io):
// reproduced as written.  simple (model-simple-ode-drj-cmp.R:108-113): E, I, R, Re, Rt.
// Re..o.Re are derivs()' output variables (model-agen.cpp:159-163): deSolve evaluates them with one
// derivs() call at the output time, i.e. with the LITERAL `t > tk` schedule; when the bounds guard
// returns early the reference leaves them stale — NaN here.
template <int FL>
__device__ __noinline__ void write_ext(const DevModel &M, const double *__restrict__ th, int64_t ld, const double *bp,
                                       int nbp, double t, const double *y, double inv1, double *__restrict__ eo, int P,
                                       int d) {
  if constexpr (Traits<FL>::kAge) {
    const double Sy = y[0], E1y = y[1], E2y = y[2], Iy = y[3], Ry = y[4];
    const double So = y[5], E1o = y[6], E2o = y[7], Io = y[8], Ro = y[9];
    const double Ny = M.Ny, No = M.No;
    eo[0 * P + d] = E1y + E2y; eo[1 * P + d] = Iy; eo[2 * P + d] = Ry;
    eo[3 * P + d] = Traits<FL>::k2i ? E1o : E1o + E2o;
    eo[4 * P + d] = Traits<FL>::k2i ? E2o + Io : Io;
    eo[5 * P + d] = Ro;
    double o0 = NAN, o1 = NAN, o2 = NAN, o3 = NAN;
    if (!(Sy < 0 || E1y < 0 || E2y < 0 || Iy < 0 || Ry < 0 || Sy > Ny || E1y > Ny || E2y > Ny || Iy > Ny || Ry > Ny ||
          So < 0 || E1o < 0 || E2o < 0 || Io < 0 || Ro < 0 || So > No || E1o > No || E2o > No || Io > No || Ro > No)) {
      int k = -1;
      for (int i = nbp - 1; i >= 0; --i)
        if (t > bp[i]) { k = i; break; }
      double by, bo, byo;
      age_beta<FL>(M, th, ld, k < 0 ? 0 : k, by, bo, byo);
      if (k >= 0 && age_linear(M, k)) {
        double ny, no, nyo;
        age_beta<FL>(M, th, ld, k + 1, ny, no, nyo);
        const double f = (t - bp[k]) / (bp[k + 1] - bp[k]);
        by = by + f * (ny - by); bo = bo + f * (no - bo); byo = byo + f * (nyo - byo);
      }
      // (model-ager.cpp:176-178: the younger group's infections run on the effective susceptibles)
      const double Syx = Traits<FL>::k2x ? fmax(0.0, Ny * 0.8 - (Ny - Sy)) : Sy;
      const double yinf = (by * Iy) / Ny * Syx;
      const double oinf = ((bo * Io) / No + (byo * Iy) / Ny) * So;
      const double ig = 1 / M.gamma;
      o0 = ig * (yinf + oinf) / (Iy + Io);
      o1 = o0 * (Ny + No) / (Sy + So);
      o2 = ig * ((by * Iy) / Ny * Sy + (byo * Iy) / Ny * So) / Iy;
      o3 = ig * (bo * Io) / No * So / Io;
    }
    eo[6 * P + d] = o0; eo[7 * P + d] = o1; eo[8 * P + d] = o2; eo[9 * P + d] = o3;
  } else {
    const double S = y[0], E = y[1], I = y[2], R = y[3], N = M.N;
    eo[0 * P + d] = E; eo[1 * P + d] = I; eo[2 * P + d] = R;
    double re = NAN, rt = NAN;
    if (!(S < 0 || E < 0 || I < 0 || R < 0 || S > N || E > N || I > N || R > N)) {
      auto T = [&](int i) { return th[(int64_t)i * ld]; };
      const double Tinf0 = T(2), Tinft = T(3);
      const double beta0 = T(0) / Tinf0, betat = T(1) / Tinft;
      double beta = beta0, Tinf = Tinf0;
      if (nbp == 2 && t > bp[1]) { beta = betat; Tinf = Tinft; }
      else if (nbp == 2 && t > bp[0]) {
        const double f = (t - bp[0]) / (bp[1] - bp[0]);
        beta = beta0 + f * (betat - beta0); Tinf = Tinf0 + f * (Tinft - Tinf0);
      }
      if constexpr (FL == EPI_FLAVOUR_NE) re = Tinf * beta * fmax(0.0, inv1 - (N - S)) / inv1;  // inv1 = Ne
      else re = Tinf * beta * S / N;
      rt = beta * Tinf;
    }
    eo[3 * P + d] = re; eo[4 * P + d] = rt;
  }
}

// What the history rows between the kernels hold at simulation day t: S of each group — or, for the age-2x
// flavour, the infection incidences y.i / o.i, i.e. derivs()' output variables 5 and 6 (model-ager.cpp:206-207) as
// deSolve evaluates them for the output matrix: one derivs() call at the output time with the LITERAL `t > tk`
// schedule, in the reference's own operation order (explicit roundings: bit-identical to the oracle's restatement
// for the same state).  When the bounds guard returns early the reference leaves them stale: NaN here.
template <int FL>
// kknown >= -1: the caller knows the schedule index of the literal rule (at the end of a day it is the segment of the
// day's last piece: pieces are cut at the breakpoints, so no breakpoint lies in (start of the piece, t), and one AT t
// belongs to neither rule) — the scan below reads bp[] from local memory, thirteen dependent L1 misses per day and
// chain (age-2x pass 2 at 8,192 chains: 7,800 cycles per day against 1,100-1,500 for age-2q).
__device__ __noinline__ void incidence_out(const DevModel &M, const double *__restrict__ th, int64_t ld, const double *bp,
                                           int nbp, double t, const double *y, double &yi, double &oi, int kknown = -2) {
  const double Sy = y[0], E1y = y[1], E2y = y[2], Iy = y[3], Ry = y[4];
  const double So = y[5], E1o = y[6], E2o = y[7], Io = y[8], Ro = y[9];
  const double Ny = M.Ny, No = M.No;
  yi = NAN; oi = NAN;
  if (Sy < 0 || E1y < 0 || E2y < 0 || Iy < 0 || Ry < 0 || Sy > Ny || E1y > Ny || E2y > Ny || Iy > Ny || Ry > Ny ||
      So < 0 || E1o < 0 || E2o < 0 || Io < 0 || Ro < 0 || So > No || E1o > No || E2o > No || Io > No || Ro > No)
    return;
  int k = kknown;
  if (kknown < -1) {
    k = -1;
    for (int i = nbp - 1; i >= 0; --i)
      if (t > bp[i]) { k = i; break; }
  }
  double by, bo, byo;
  age_beta<FL>(M, th, ld, k < 0 ? 0 : k, by, bo, byo);
  if (k >= 0 && age_linear(M, k)) {
    double ny, no, nyo;
    age_beta<FL>(M, th, ld, k + 1, ny, no, nyo);
    const double f = __ddiv_rn(__dsub_rn(t, bp[k]), __dsub_rn(bp[k + 1], bp[k]));
    by = __dadd_rn(by, __dmul_rn(f, __dsub_rn(ny, by)));
    bo = __dadd_rn(bo, __dmul_rn(f, __dsub_rn(no, bo)));
    byo = __dadd_rn(byo, __dmul_rn(f, __dsub_rn(nyo, byo)));
  }
  const double Syeff = fmax(0.0, __dsub_rn(__dmul_rn(Ny, 0.8), __dsub_rn(Ny, Sy)));
  yi = __dmul_rn(__ddiv_rn(__dmul_rn(by, Iy), Ny), Syeff);
  oi = __dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(bo, Io), No), __ddiv_rn(__dmul_rn(byo, Iy), Ny)), So);
}

// TRACKR = false (the hot instance): the removed compartment is not integrated (see below_one); a chain where
// that could have changed a guard decision sets need_r[chain] and is recomputed by the TRACKR = true instance,
// which is launched right behind and only works on those chains.  FULL (ext trajectories) always tracks R.
template <int FL, bool PASS2, bool FULL = false, bool TRACKR = false>
__global__ void __maxnreg__(128)
k_ode(const DevModel M, const double *__restrict__ theta, int64_t ld, int64_t n,
      const int *__restrict__ offset, const int *__restrict__ padv, const double *__restrict__ hist1,
      double *__restrict__ hist_out, double *__restrict__ ckpt, double *__restrict__ series = nullptr,
      int ns_total = 0, int window_only = 0, int ndays_limit = 0, const int *__restrict__ retry_status = nullptr,
      int *__restrict__ need_r_flag = nullptr, int force_need_r = 0) {
  constexpr int NEQ = Traits<FL>::NEQ, NG = Traits<FL>::NG;
  constexpr bool TR = TRACKR || FULL || Traits<FL>::k2x;  // integrate the removed compartment
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if constexpr (TRACKR && !FULL) {
    if (!need_r_flag[i]) return;  // fallback instance: only the chains the hot instance flagged
  }
  bool need_r = false;
  const double *th = theta + i;
  const int ndays = PASS2 ? M.P : M.P1;
  const int pitch = PASS2 ? hist_pitch2(ndays) : hist_pitch(ndays);
  const int Q = pitch >> 2;  // phase-row length of the pass-2 history (hist_index)
  double *out = hist_out + (i * NG) * (int64_t)pitch;

  double y[NEQ];
  double inv0, inv1;
  unsigned nb0, nb1 = 0;  // high words of the group sizes, minus one
  if constexpr (Traits<FL>::kAge) {
    y[0] = M.Ny - 1.0; y[1] = 1.0; y[2] = y[3] = y[4] = 0.0;
    y[5] = M.No; y[6] = y[7] = y[8] = y[9] = 0.0;
    inv0 = 1.0 / M.Ny; inv1 = 1.0 / M.No;
    nb0 = (unsigned)__double2hiint(M.Ny) - 1u;
    nb1 = (unsigned)__double2hiint(M.No) - 1u;
  } else {
    y[0] = M.N - 1.0; y[1] = 1.0; y[2] = y[3] = 0.0;
    if constexpr (FL == EPI_FLAVOUR_NE) { inv1 = th[8 * ld] * M.N; inv0 = 1.0 / inv1; }
    else { inv0 = 1.0 / M.N; inv1 = 0.0; }
    nb0 = (unsigned)__double2hiint(M.N) - 1u;
  }

  double bp[kMaxBp];
  int nbp = 0;
  double tcut = INFINITY;
  int t0 = 1;
  int nint = ndays;  // days integrated here
  if constexpr (!PASS2) {
    // pass 1 in two rounds when it covers the whole period (simple / age-2i): the first round integrates only
    // the first ndays_limit days, where the threshold is usually crossed; chains that did not cross there are
    // flagged by k_mid and integrated over the whole pass-1 period by the retry round
    if (retry_status && !(retry_status[i] & kStRetry)) return;
    if (ndays_limit > 0) nint = min(nint, ndays_limit);
  }
  // FULL: the ext series of this chain, [series][P] behind the basic ones (8 age / 3 simple)
  double *eo = nullptr;
  if constexpr (FULL) eo = series + ((int64_t)i * ns_total + (Traits<FL>::kAge ? 8 : 3)) * (int64_t)M.P;
  if constexpr (PASS2) {
    const int off = offset[i];
    if (off == EPI_INVALID_DATA_OFFSET) {
      // pass 2 skipped: the pass-1 rows are recycled into the P-length slices
      // (model-age-groups2q.R:216-223)
      const int pitch1 = hist_pitch(M.P1);
      const double *in = hist1 + (i * NG) * (int64_t)pitch1;
      for (int d = 0; d < M.P; ++d)
#pragma unroll
        for (int g = 0; g < NG; ++g) out[g * pitch + hist_index<PASS2>(d, Q)] = in[g * pitch1 + d % M.P1];
      if constexpr (!FULL) {
        if constexpr (!TR) { if (need_r_flag) need_r_flag[i] = 0; }
        return;
      }
      // FULL: the other columns of the recycled matrix are not kept by pass 1: integrate its days
      // again (beta0 forever, bit-identical) and recycle the rows below
      t0 = padv[i] + 1;
      nint = min(M.P1, M.P);
    } else {
      t0 = padv[i] + 1;
      nbp = breakpoints<FL>(M, th, ld, (double)off, bp);
      if (!FULL && window_only) nint = min(nint, max(last_window_day(M, off, padv[i]) - padv[i], 1));
    }
  }
  const int cdays = ckpt_days(M.P1);
  // element (d, q) of this chain: ck[(d * NEQ + q) * kTile]
  double *ck = ckpt ? ckpt + ((i / kTile) * (int64_t)cdays * NEQ) * kTile + (i % kTile) : nullptr;
  int d0 = 0;  // last day already known
  if constexpr (PASS2 && !TRACKR) {  // (the R-tracking fallback integrates from day 0: the checkpoints hold no R)
    if (ck) {
      double minbp = INFINITY;
      for (int k = 0; k < nbp; ++k) minbp = fmin(minbp, bp[k]);
      const double dd = floor(minbp) - (double)t0;
      d0 = dd > 0 ? (int)fmin(dd, (double)min(cdays - 1, ndays - 1)) : 0;
      if (d0 > 0) {
        const int pitch1 = hist_pitch(M.P1);
        const double *in = hist1 + (i * NG) * (int64_t)pitch1;
        for (int d = 0; d <= d0; ++d)
#pragma unroll
          for (int g = 0; g < NG; ++g) out[g * pitch + hist_index<PASS2>(d, Q)] = in[g * pitch1 + d];
#pragma unroll
        for (int q = 0; q < NEQ; ++q)
          if (TR || !is_r_state<FL>(q)) y[q] = ck[((int64_t)d0 * NEQ + q) * kTile];
      }
    }
  }
  bool bp_sorted = true;  // (NaN breakpoints compare false: the full scan handles them as before)
  for (int k = 1; k < nbp; ++k) bp_sorted = bp_sorted && bp[k - 1] <= bp[k];
  int kseg = -2;          // segment index not known yet
  Seg sg = select_segment<FL>(M, th, ld, bp, nbp, (double)(t0 + d0), &tcut, inv0, inv1, &kseg, bp_sorted);

  const int m = M.m;
  const double h = 1.0 / (double)m;
  if (d0 == 0) {
    if constexpr (Traits<FL>::k2x) {
      double yi, oi;
      incidence_out<FL>(M, th, ld, bp, nbp, (double)t0, y, yi, oi);
      out[0] = yi; out[pitch] = oi;
      if constexpr (FULL) { eo[-8 * (int64_t)M.P] = y[0]; eo[-7 * (int64_t)M.P] = y[5]; }  // series 0, 1: y.S, o.S
    } else {
      out[0] = y[0];
      if constexpr (NG == 2) out[pitch] = y[5];
    }
    if constexpr (FULL) write_ext<FL>(M, th, ld, bp, nbp, (double)t0, y, inv1, eo, M.P, 0);
  }
  if constexpr (!PASS2) {
    if (ck) {
#pragma unroll
      for (int q = 0; q < NEQ; ++q)
        if (TR || !is_r_state<FL>(q)) ck[q * kTile] = y[q];
    }
  }
  // a regular sub-step has hh = h exactly when 1/m is a power of two (m = 1, 2, 4, 8, ...: tday + s*h is then exact
  // too): its hh, hh/2 and hh/6 are constants of the kernel and the sub-step boundaries follow by addition.  A day
  // takes that route when no breakpoint falls inside it; otherwise (and for other m) every piece computes them.
  // Both routes give a regular day the same bits, so the choice is made per WARP: a day that a breakpoint cuts in
  // one lane takes every converged lane through the general loop (4 or 5 pieces) instead of serialising the two
  // routes (4 + 5) — chains of a running sampler have their breakpoints on different days (tests/_samp_overhead.py:
  // pass 2 on a DE-MC population 1.29 ms against 0.91 ms on the bench's narrow theta cloud before this vote).
  const bool pow2 = (m & (m - 1)) == 0;
  const double hh2reg = 0.5 * h, h6reg = h / 6.0;
  for (int d = d0 + 1; d < nint; ++d) {
    const double tday = (double)(t0 + d - 1);
    const bool regular = pow2 && __all_sync(__activemask(), tcut >= tday + 1.0);
    double pa = tday;
    int s = 0;
    if (regular && (m & 1) == 0) {
      // sub-steps in pairs, ping-pong between two state arrays: no copy of the new state back into the old one
      double z[NEQ];
      for (; s < m; s += 2) {
        const double pm = pa + h, pb = pm + h;
        rk4_step<FL, TR>(M, sg, y, z, pa, pm, h, hh2reg, h6reg, inv0, inv1, nb0, nb1, need_r);
        rk4_step<FL, TR>(M, sg, z, y, pm, pb, h, hh2reg, h6reg, inv0, inv1, nb0, nb1, need_r);
        pa = pb;
      }
    }
    while (s < m) {
      double pb, hh, hh2, h6;
      bool step_done = true;
      if (regular) {
        pb = pa + h; hh = h; hh2 = hh2reg; h6 = h6reg;
      } else {
        const double b = (s == m - 1) ? tday + 1.0 : tday + (double)(s + 1) * h;
        if constexpr (PASS2) {
          if (pa >= tcut) sg = select_segment<FL>(M, th, ld, bp, nbp, pa, &tcut, inv0, inv1, &kseg, bp_sorted);
        }
        pb = b;
        if constexpr (PASS2) {
          if (tcut < b) { pb = tcut; step_done = false; }  // a breakpoint strictly inside (pa, b): cut there
        }
        hh = pb - pa; hh2 = 0.5 * hh; h6 = div6(hh);
      }
      rk4_piece<FL, TR>(M, sg, y, pa, pb, hh, hh2, h6, inv0, inv1, nb0, nb1, need_r);
      pa = pb;
      if (step_done) ++s;
    }
    if constexpr (Traits<FL>::k2x) {
      double yi, oi;
      incidence_out<FL>(M, th, ld, bp, nbp, (double)(t0 + d), y, yi, oi, bp_sorted ? kseg : -2);
      out[hist_index<PASS2>(d, Q)] = yi; out[pitch + hist_index<PASS2>(d, Q)] = oi;
      if constexpr (FULL) { eo[-8 * (int64_t)M.P + d] = y[0]; eo[-7 * (int64_t)M.P + d] = y[5]; }
    } else {
      out[hist_index<PASS2>(d, Q)] = y[0];
      if constexpr (NG == 2) out[pitch + hist_index<PASS2>(d, Q)] = y[5];
    }
    if constexpr (FULL) write_ext<FL>(M, th, ld, bp, nbp, (double)(t0 + d), y, inv1, eo, M.P, d);
    if constexpr (!PASS2) {
      if (ck && d < cdays) {
#pragma unroll
        for (int q = 0; q < NEQ; ++q)
          if (TR || !is_r_state<FL>(q)) ck[((int64_t)d * NEQ + q) * kTile] = y[q];
      }
    }
  }
  if constexpr (!TR) {
    if (need_r_flag) need_r_flag[i] = (need_r || force_need_r) ? 1 : 0;
  }
  if constexpr (FULL) {
    // invalid data_offset with a pass 1 shorter than the period: R recycles the short matrix (:216-227)
    constexpr int NE = Traits<FL>::kAge ? 10 : 5;
    for (int d = nint; d < ndays; ++d) {
#pragma unroll
      for (int q = 0; q < NE; ++q) eo[q * M.P + d] = eo[q * M.P + d % nint];
      if constexpr (Traits<FL>::k2x) {
        eo[-8 * (int64_t)M.P + d] = eo[-8 * (int64_t)M.P + d % nint];
        eo[-7 * (int64_t)M.P + d] = eo[-7 * (int64_t)M.P + d % nint];
      }
    }
  }
}

// ------------------------------------------------- K_ode, two lanes per chain
// One lane per age group: the even lane integrates (Sy, E1y, E2y, Iy), the odd lane (So, E1o, E2o, Io); the only
// coupling is I_y, one shuffle per RK4 stage (model-agen.cpp:136-144), which the younger group's lane does not
// even wait for.  Like the four-lane kernel below it is WARP-SYNCHRONOUS (all 16 chains of a warp execute every
// piece together, full-warp shuffles and votes, `commit` freezes the chains that have nothing to do): round 1's
// two-lane kernel shuffled over per-pair masks behind early returns and never beat the other mappings.  Every
// operation is the explicitly rounded one of rhs<> (the younger group's beta_yo = 0 term adds an exact +0): same bits
// as the thread-per-chain kernel.  R is not integrated; chains whose guard decision could have depended on it are
// flagged for the R-tracking instance of k_ode.
struct Seg2 {
  double tk, bA, sA, bB, sB;  // y lane: A = betay, B = 0;  o lane: A = betao, B = betayo (already divided by N)
  bool fast_ok;               // all three beta curves >= 0 on the segment (premise of the fast guard)
  bool hold;                  // all three slopes of the CHAIN are zero
};

constexpr unsigned kFull = 0xffffffffu;

template <int FL>
__device__ __noinline__ Seg2 select_segment_age2(const DevModel &M, const double *__restrict__ th, int64_t ld, const double *bp, int nbp,
                                                 double pa, int is_o, double *tcut, double invy, double invo, int *kcur,
                                                 bool sorted) {
  int k = -1;
  double tc = INFINITY;
  if (sorted && *kcur >= -1) {
    k = *kcur;
    while (k + 1 < nbp && bp[k + 1] <= pa) ++k;
    if (k + 1 < nbp) tc = bp[k + 1];
  } else {
    for (int i = 0; i < nbp; ++i) {
      if (bp[i] <= pa) k = i;
      if (bp[i] > pa && bp[i] < tc) tc = bp[i];
    }
  }
  *kcur = k;
  *tcut = tc;
  double by, bo, byo;
  age_beta<FL>(M, th, ld, k < 0 ? 0 : k, by, bo, byo);
  Seg2 sg;
  sg.tk = 0.0; sg.sA = 0.0; sg.sB = 0.0;
  sg.bA = is_o ? bo : by;
  sg.bB = is_o ? byo : 0.0;
  if (k >= 0 && age_linear(M, k)) {
    double ny, no, nyo;
    age_beta<FL>(M, th, ld, k + 1, ny, no, nyo);
    const double len = bp[k + 1] - bp[k];
    sg.tk = bp[k];
    sg.sA = is_o ? (no - bo) / len : (ny - by) / len;
    sg.sB = is_o ? (nyo - byo) / len : 0.0;
    sg.fast_ok = by >= 0 && bo >= 0 && byo >= 0 && ny >= 0 && no >= 0 && nyo >= 0 && len > 0;
    // hold <=> all three slopes of the chain are zero (the same test as select_segment, on the same roundings)
    sg.hold = __dmul_rn((ny - by) / len, invy) == 0.0 && __dmul_rn((no - bo) / len, invo) == 0.0 &&
              __dmul_rn((nyo - byo) / len, invy) == 0.0;
  } else {
    sg.fast_ok = by >= 0 && bo >= 0 && byo >= 0;
    sg.hold = true;
  }
  // the same normalisation (and the same roundings) as select_segment
  const double invA = is_o ? invo : invy;
  sg.bA = __dmul_rn(sg.bA, invA); sg.sA = __dmul_rn(sg.sA, invA);
  sg.bB = __dmul_rn(sg.bB, invy); sg.sB = __dmul_rn(sg.sB, invy);
  return sg;
}

// The schedule of one chain as a TABLE in shared memory (pass 2): entry e = k + 1 holds the lane's Seg2 of schedule
// index k = -1 .. nbp - 1 (beta pair and slopes already divided by N, flags), computed once in the kernel prologue
// with the operations of select_segment_age2 — same bits.  Cycle counters in the kernel (profiles/r2_summary.md,
// 8,192 chains): a warp re-selected ~60 times per launch at ~1,500 cycles each (two dependent global loads through
// the schedule entry and theta, three divisions) = 14 % of pass 2; from the table it is a scan of bp[] and five LDS.
// Layout [entry][field][thread]: consecutive lanes -> consecutive banks.
constexpr int kSegFields = 6;  // by', sy', bo', so', byo', syo'  (beta and slope of the three curves, divided by N)
// Shared memory of a block, per PAIR of lanes (np = blockDim.x / 2 pairs; a per-lane copy was 66 KB per block and
// three blocks per SM): the breakpoints [NBP][np] (the re-selection scans them: in local memory every probe was an
// L1 miss, ~800 cycles per call), the entries [NBP + 1][field][np], then the flags [NBP + 1][np] as ints.
template <int FL>
__host__ __device__ constexpr size_t seg_table_bytes(int threads) {
  return (size_t)(Traits<FL>::NBP + (Traits<FL>::NBP + 1) * kSegFields) * (threads / 2) * sizeof(double) +
         (size_t)(Traits<FL>::NBP + 1) * (threads / 2) * sizeof(int);
}

struct SegTab {
  double *bps;  // + pair index
  double *tab;
  int *flags;
  int np;
};

template <int FL>
__device__ __forceinline__ SegTab seg_tab_of(double *smem, int nthreads, int tid) {
  SegTab T;
  T.np = nthreads >> 1;
  T.bps = smem + (tid >> 1);
  T.tab = smem + Traits<FL>::NBP * T.np + (tid >> 1);
  T.flags = reinterpret_cast<int *>(smem + (Traits<FL>::NBP + (Traits<FL>::NBP + 1) * kSegFields) * T.np) + (tid >> 1);
  return T;
}

// Warp-synchronous (every lane of the warp calls it, uniform trip counts).  The younger group's lane writes the
// breakpoints, its own curve and the flags, the older group's lane its two curves.
template <int FL>
__device__ __forceinline__ void fill_segment_table(const DevModel &M, const double *__restrict__ th, int64_t ld, const double *bp,
                                                   int nbp, int is_o, double invy, double invo, const SegTab &T) {
  constexpr int NBP = Traits<FL>::NBP;
  const int np = T.np;
  // raw beta triples first, all loads of the schedule issued back to back (they are independent; entry by entry the
  // prologue paid two dependent global round trips per entry); parked in the beta fields of entry k + 1
#pragma unroll
  for (int k = 0; k < NBP; ++k) {
    double by = 0.0, bo = 0.0, byo = 0.0;
    if (k < nbp) age_beta<FL>(M, th, ld, k, by, bo, byo);
    if (!is_o) {
      double *t = T.tab + (size_t)(k + 1) * kSegFields * np;
      t[0] = by; t[2 * np] = bo; t[4 * np] = byo;
      T.bps[k * np] = k < nbp ? bp[k] : INFINITY;
    }
  }
  __syncwarp();
  // entry k + 1 from the triples k and k + 1, in ascending order (entry k + 1 overwrites triple k only after both
  // lanes hold it in registers; triple k + 1 is still raw).  Entry 0 (k = -1: before the first breakpoint) is beta_0 held.
  double by = T.tab[(size_t)kSegFields * np], bo = T.tab[(size_t)(kSegFields + 2) * np], byo = T.tab[(size_t)(kSegFields + 4) * np];
  __syncwarp();
#pragma unroll 1
  for (int e = 0; e <= NBP; ++e) {
    const int k = e - 1;
    if (e <= nbp) {
      double ny = 0.0, no = 0.0, nyo = 0.0;
      if (k + 1 < nbp && k + 1 >= 1) {
        const double *t = T.tab + (size_t)(k + 2) * kSegFields * np;
        ny = t[0]; no = t[2 * np]; nyo = t[4 * np];
      }
      double sy = 0.0, so = 0.0, syo = 0.0;
      bool fast_ok, hold, lin = false;
      if (k >= 0 && age_linear(M, k)) {
        const double len = T.bps[(k + 1) * np] - T.bps[k * np];
        lin = true;
        sy = (ny - by) / len; so = (no - bo) / len; syo = (nyo - byo) / len;
        fast_ok = by >= 0 && bo >= 0 && byo >= 0 && ny >= 0 && no >= 0 && nyo >= 0 && len > 0;
        hold = __dmul_rn(sy, invy) == 0.0 && __dmul_rn(so, invo) == 0.0 && __dmul_rn(syo, invy) == 0.0;
      } else {
        fast_ok = by >= 0 && bo >= 0 && byo >= 0;
        hold = true;
      }
      double *t = T.tab + (size_t)e * kSegFields * np;
      if (!is_o) {
        t[0] = __dmul_rn(by, invy);
        t[np] = __dmul_rn(sy, invy);
        T.flags[e * np] = (fast_ok ? 1 : 0) | (hold ? 2 : 0) | (lin ? 4 : 0);
      } else {
        t[2 * np] = __dmul_rn(bo, invo);
        t[3 * np] = __dmul_rn(so, invo);
        t[4 * np] = __dmul_rn(byo, invy);
        t[5 * np] = __dmul_rn(syo, invy);
      }
      if (k >= 0) { by = ny; bo = no; byo = nyo; }  // (k = -1 and k = 0 both start from beta_0)
    }
    __syncwarp();
  }
}

// the segment active on a piece that starts at pa (the rule of select_segment), read from the table
__device__ __forceinline__ Seg2 select_segment_tab(const SegTab &T, int is_o, int nbp, double pa, double *tcut, int *kcur,
                                                   bool sorted) {
  const int np = T.np;
  int k = -1;
  double tc = INFINITY;
  if (sorted && *kcur >= -1) {
    k = *kcur;
    while (k + 1 < nbp && T.bps[(k + 1) * np] <= pa) ++k;
    if (k + 1 < nbp) tc = T.bps[(k + 1) * np];
  } else {
    for (int i = 0; i < nbp; ++i) {
      const double b = T.bps[i * np];
      if (b <= pa) k = i;
      if (b > pa && b < tc) tc = b;
    }
  }
  *kcur = k;
  *tcut = tc;
  const double *t = T.tab + (size_t)(k + 1) * kSegFields * np;
  Seg2 sg;
  if (is_o) { sg.bA = t[2 * np]; sg.sA = t[3 * np]; sg.bB = t[4 * np]; sg.sB = t[5 * np]; }
  else { sg.bA = t[0]; sg.sA = t[np]; sg.bB = 0.0; sg.sB = 0.0; }
  const int fl = T.flags[(k + 1) * np];
  sg.fast_ok = fl & 1; sg.hold = fl & 2;
  sg.tk = (fl & 4) ? T.bps[k * np] : 0.0;
  return sg;
}

struct Lane2 {
  double a1, gamma, Ngrp;
  double eta, cap;  // age-2x: waning rate; 0.8 * N_y (the younger group's effective-susceptible cap, model-ager.cpp:176-178)
  bool is_y;
  int shift;     // first lane of this chain's pair == the lane that holds I_y
  unsigned nb;   // hi(N) - 1
};

// NS = 4: (S, E1, E2, I) per lane, R not integrated (age-2q / 2i).  NS = 5: age-2x (model-ager.cpp:176-200) — R is a
// state (it flows back into S at rate eta), the younger group is infected through its effective susceptibles, and
// "S > N" gets its own exact test (S is no longer monotone); the operations are those of rhs<> for that flavour.
template <bool GUARDED, bool HOLD, int NS = 4>
__device__ __forceinline__ void rhs2w(const Lane2 &L, const Seg2 &sg, double t, const double *y, double *dy, unsigned &flag) {
  constexpr bool X2 = NS == 5;
  const double S = y[0], E1 = y[1], E2 = y[2], I = y[3], R = X2 ? y[NS - 1] : 0.0;
  const double iy = __shfl_sync(kFull, I, L.shift);  // I of the younger group
  bool bad = false;
  if constexpr (GUARDED) {
    // bounds guard over the chain's states, model-agen.cpp:109-124 (NS = 4: R is not integrated here, see below_one)
    const int bad_own = (S < 0 || E1 < 0 || E2 < 0 || I < 0 || R < 0 || S > L.Ngrp || E1 > L.Ngrp || E2 > L.Ngrp ||
                         I > L.Ngrp || R > L.Ngrp);
    bad = ((__ballot_sync(kFull, bad_own) >> L.shift) & 3u) != 0u;
    if constexpr (!X2) flag |= (!bad && S < 1.0) ? 1u : 0u;  // the R test is undecidable here
  } else if constexpr (X2) {
    flag |= oob_group5(S, E1, E2, I, R, L.nb) | ((S > L.Ngrp) ? 0x80000000u : 0u);
  } else {
    flag |= oob_group4_nor(S, E1, E2, I, L.nb);  // bit 31 set <=> maybe out of bounds
  }
  const double dt = HOLD ? 0.0 : t - sg.tk;
  const double bA = HOLD ? sg.bA : fma(dt, sg.sA, sg.bA), bB = HOLD ? sg.bB : fma(dt, sg.sB, sg.bB);
  double Sx = S;
  if constexpr (X2) Sx = L.is_y ? fmax(0.0, __dsub_rn(L.cap, __dsub_rn(L.Ngrp, S))) : S;
  const double inf = __dmul_rn(fma(bA, I, __dmul_rn(bB, iy)), Sx);  // y: (betay/Ny*Iy)*Sy; o: (betao/No*Io + betayo/Ny*Iy)*So
  if constexpr (X2) {
    const double rev = __dmul_rn(L.eta, R);
    dy[0] = __dsub_rn(rev, inf);
    dy[4] = fma(L.gamma, I, -rev);
  } else {
    dy[0] = -inf;
  }
  dy[1] = fma(-L.a1, E1, inf);
  dy[2] = fma(L.a1, E1, -E2);
  dy[3] = fma(-L.gamma, I, E2);
  if constexpr (GUARDED) {
    if (bad) {  // the guard freezes the whole state
#pragma unroll
      for (int i = 0; i < NS; ++i) dy[i] = 0.0;
    }
  }
}

// returns (fast instance) whether no lane of the chain raised the maybe-out-of-bounds bit
template <bool GUARDED, bool HOLD, int NS = 4>
__device__ __forceinline__ bool rk4_try2w(const Lane2 &L, const Seg2 &sg, const double *y, double *yn, double pa, double pb,
                                          double hh, double hh2, double h6, bool &need_r) {
  const double tm = pa + hh2;
  double k[NS], acc[NS], yt[NS];
  unsigned flag = 0;
  rhs2w<GUARDED, HOLD, NS>(L, sg, pa, y, k, flag);
#pragma unroll
  for (int i = 0; i < NS; ++i) { acc[i] = k[i]; yt[i] = fma(hh2, k[i], y[i]); }
  rhs2w<GUARDED, HOLD, NS>(L, sg, tm, yt, k, flag);
#pragma unroll
  for (int i = 0; i < NS; ++i) { acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh2, k[i], y[i]); }
  rhs2w<GUARDED, HOLD, NS>(L, sg, tm, yt, k, flag);
#pragma unroll
  for (int i = 0; i < NS; ++i) { acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh, k[i], y[i]); }
  rhs2w<GUARDED, HOLD, NS>(L, sg, pb, yt, k, flag);
#pragma unroll
  for (int i = 0; i < NS; ++i) yn[i] = fma(h6, __dadd_rn(acc[i], k[i]), y[i]);
  if constexpr (GUARDED) {
    need_r = (flag & 1u) != 0u;
    return true;
  }
  return ((__ballot_sync(kFull, (int)flag < 0) >> L.shift) & 3u) == 0u;
}

// the fast instance of a step without its verdict: returns the accumulated maybe-out-of-bounds flag (bit 31) for the
// caller to vote on — k_ode_age2 votes once per DAY on its regular days
template <bool HOLD, int NS = 4>
__device__ __forceinline__ unsigned rk4_raw2w(const Lane2 &L, const Seg2 &sg, const double *y, double *yn, double pa, double pb,
                                             double hh, double hh2, double h6) {
  const double tm = pa + hh2;
  double k[NS], acc[NS], yt[NS];
  unsigned flag = 0;
  rhs2w<false, HOLD, NS>(L, sg, pa, y, k, flag);
#pragma unroll
  for (int i = 0; i < NS; ++i) { acc[i] = k[i]; yt[i] = fma(hh2, k[i], y[i]); }
  rhs2w<false, HOLD, NS>(L, sg, tm, yt, k, flag);
#pragma unroll
  for (int i = 0; i < NS; ++i) { acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh2, k[i], y[i]); }
  rhs2w<false, HOLD, NS>(L, sg, tm, yt, k, flag);
#pragma unroll
  for (int i = 0; i < NS; ++i) { acc[i] = fma(2.0, k[i], acc[i]); yt[i] = fma(hh, k[i], y[i]); }
  rhs2w<false, HOLD, NS>(L, sg, pb, yt, k, flag);
#pragma unroll
  for (int i = 0; i < NS; ++i) yn[i] = fma(h6, __dadd_rn(acc[i], k[i]), y[i]);
  return flag;
}

// a whole regular day (m sub-steps, m a power of two) of fast steps from y into w, ping-pong over z; returns the
// accumulated flag
template <bool HOLD, int NS = 4>
__device__ __forceinline__ unsigned rk4_day2w(const Lane2 &L, const Seg2 &sg, const double *y, double *w, double tday, double h,
                                             double hh2, double h6, int m) {
  double z[NS];
  double pa = tday, pm, pb;
  unsigned flag;
  int s;
  if (m & 1) {  // (m == 1)
    pb = pa + h;
    flag = rk4_raw2w<HOLD, NS>(L, sg, y, w, pa, pb, h, hh2, h6);
    s = 1;
  } else {
    pm = pa + h; pb = pm + h;
    flag = rk4_raw2w<HOLD, NS>(L, sg, y, z, pa, pm, h, hh2, h6);
    flag |= rk4_raw2w<HOLD, NS>(L, sg, z, w, pm, pb, h, hh2, h6);
    s = 2;
  }
  for (; s < m; s += 2) {
    pa = pb; pm = pa + h; pb = pm + h;
    flag |= rk4_raw2w<HOLD, NS>(L, sg, w, z, pa, pm, h, hh2, h6);
    flag |= rk4_raw2w<HOLD, NS>(L, sg, z, w, pm, pb, h, hh2, h6);
  }
  return flag;
}

// one RK4 piece for all sixteen chains of the warp; a chain takes the result only when `commit`
template <int NS = 4>
__device__ __forceinline__ void rk4_step2w(const Lane2 &L, const Seg2 &sg, double *y, bool commit, double pa, double pb,
                                           double hh, double hh2, double h6, bool &need_r) {
  double yn[NS];
  bool nr = false;
  const bool hold = __all_sync(kFull, sg.hold);
  bool ok = hold ? rk4_try2w<false, true, NS>(L, sg, y, yn, pa, pb, hh, hh2, h6, nr)
                 : rk4_try2w<false, false, NS>(L, sg, y, yn, pa, pb, hh, hh2, h6, nr);
  ok = ok && sg.fast_ok;
  if (__any_sync(kFull, commit && !ok)) {
    // rare: the reference's guard evaluated exactly in FP64, by the whole warp; taken by the chains that need it
    double gy[NS];
    bool gnr = false;
    rk4_try2w<true, false, NS>(L, sg, y, gy, pa, pb, hh, hh2, h6, gnr);
    if (!ok) {
#pragma unroll
      for (int i = 0; i < NS; ++i) yn[i] = gy[i];
      nr = gnr;
    }
  }
  if (commit) {
#pragma unroll
    for (int i = 0; i < NS; ++i) y[i] = yn[i];
    need_r = need_r || nr;
  }
}

// the same piece when EVERY chain of the warp takes the result (the usual case on a regular day): from yin into a
// distinct array yout, no per-lane selects and no copy back (the caller ping-pongs between two state arrays)
template <int NS = 4>
__device__ __forceinline__ void rk4_step2w_all(const Lane2 &L, const Seg2 &sg, const double *yin, double *yout, double pa,
                                               double pb, double hh, double hh2, double h6, bool &need_r) {
  bool nr = false;
  const bool hold = __all_sync(kFull, sg.hold);
  bool ok = hold ? rk4_try2w<false, true, NS>(L, sg, yin, yout, pa, pb, hh, hh2, h6, nr)
                 : rk4_try2w<false, false, NS>(L, sg, yin, yout, pa, pb, hh, hh2, h6, nr);
  ok = ok && sg.fast_ok;
  if (__any_sync(kFull, !ok)) {
    double gy[NS];
    bool gnr = false;
    rk4_try2w<true, false, NS>(L, sg, yin, gy, pa, pb, hh, hh2, h6, gnr);
    if (!ok) {
#pragma unroll
      for (int i = 0; i < NS; ++i) yout[i] = gy[i];
      nr = gnr;
    }
  }
  need_r = need_r || nr;
}

// age-2x: what the history rows hold at simulation day t — this lane's infection incidence (derivs()' output variable
// y.i resp. o.i, model-ager.cpp:206-207) as deSolve evaluates it for the output matrix: the LITERAL `t > tk` schedule,
// the reference's own operation order (the roundings of incidence_out(), bit for bit); NaN when the bounds guard over
// the chain's ten states returns early.  Warp-synchronous (ballot + shuffle): every lane calls it.
template <int FL>
__device__ __noinline__ double incidence_out2(const DevModel &M, const double *__restrict__ th, int64_t ld, const double *bps,
                                              int bstride, int nbp, double t, const Lane2 &L, const double *y, int is_o) {
  const double S = y[0], E1 = y[1], E2 = y[2], I = y[3], R = y[4];
  const int bad_own = (S < 0 || E1 < 0 || E2 < 0 || I < 0 || R < 0 || S > L.Ngrp || E1 > L.Ngrp || E2 > L.Ngrp ||
                       I > L.Ngrp || R > L.Ngrp);
  const bool bad = ((__ballot_sync(kFull, bad_own) >> L.shift) & 3u) != 0u;
  const double Iy = __shfl_sync(kFull, I, L.shift);
  int k = -1;
  for (int i = nbp - 1; i >= 0; --i)
    if (t > bps[i * bstride]) { k = i; break; }
  double by, bo, byo;
  age_beta<FL>(M, th, ld, k < 0 ? 0 : k, by, bo, byo);
  if (k >= 0 && age_linear(M, k)) {
    double ny, no, nyo;
    age_beta<FL>(M, th, ld, k + 1, ny, no, nyo);
    const double f = __ddiv_rn(__dsub_rn(t, bps[k * bstride]), __dsub_rn(bps[(k + 1) * bstride], bps[k * bstride]));
    by = __dadd_rn(by, __dmul_rn(f, __dsub_rn(ny, by)));
    bo = __dadd_rn(bo, __dmul_rn(f, __dsub_rn(no, bo)));
    byo = __dadd_rn(byo, __dmul_rn(f, __dsub_rn(nyo, byo)));
  }
  if (bad) return NAN;
  if (!is_o) {
    const double Syeff = fmax(0.0, __dsub_rn(__dmul_rn(M.Ny, 0.8), __dsub_rn(M.Ny, S)));
    return __dmul_rn(__ddiv_rn(__dmul_rn(by, I), M.Ny), Syeff);
  }
  return __dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(bo, I), M.No), __ddiv_rn(__dmul_rn(byo, Iy), M.Ny)), S);
}

template <int FL, bool PASS2, int MINB = 4>
__global__ void __launch_bounds__(128, MINB)
k_ode_age2(const DevModel M, const double *__restrict__ theta, int64_t ld, int64_t n, const int *__restrict__ offset,
           const int *__restrict__ padv, const double *__restrict__ hist1, double *__restrict__ hist_out,
           double *__restrict__ ckpt, int window_only = 0, int ndays_limit = 0,
           const int *__restrict__ retry_status = nullptr, int *__restrict__ need_r_flag = nullptr,
           int force_need_r = 0) {
  constexpr bool X2 = Traits<FL>::k2x;
  constexpr int NS = X2 ? 5 : 4;  // states per lane: (S, E1, E2, I) and, age-2x, R
  bool need_r = false;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool exists = (tid >> 1) < n;
  const int64_t i = exists ? (tid >> 1) : n - 1;  // lanes behind the batch shadow its last chain and never store
  const int lane = threadIdx.x & 31;
  const int g = lane & 1;  // 0 = younger group, 1 = older group
  Lane2 L;
  L.shift = lane & ~1;
  L.a1 = M.a1; L.gamma = M.gamma;
  L.Ngrp = g ? M.No : M.Ny;
  L.nb = (unsigned)__double2hiint(L.Ngrp) - 1u;
  L.is_y = g == 0;
  L.eta = X2 ? M.eta : 0.0;
  L.cap = __dmul_rn(M.Ny, 0.8);
  const double *th = theta + i;
  const int ndays = PASS2 ? M.P : M.P1;
  int nint = ndays;  // days integrated here
  bool alive = exists;
  if constexpr (!PASS2) {  // two-round pass 1, see k_ode
    if (retry_status && !(retry_status[i] & kStRetry)) alive = false;
    if (ndays_limit > 0) nint = min(nint, ndays_limit);
  }
  const int pitch = PASS2 ? hist_pitch2(ndays) : hist_pitch(ndays), Q = pitch >> 2;
  double *out = hist_out + (i * 2 + g) * (int64_t)pitch;
  // model-age-groups2q.R:157-158
  double y[NS];
  y[0] = g ? M.No : M.Ny - 1.0; y[1] = g ? 0.0 : 1.0;
#pragma unroll
  for (int j = 2; j < NS; ++j) y[j] = 0.0;

  double bp[kMaxBp];
  int nbp = 0;
  double tcut = INFINITY;
  int t0 = 1;
  if constexpr (PASS2) {
    const int off = offset[i];
    if (off == EPI_INVALID_DATA_OFFSET) {
      // pass 2 skipped: the pass-1 rows are recycled into the P-length slices (:216-223)
      if (alive) {
        const double *in = hist1 + (i * 2 + g) * (int64_t)hist_pitch(M.P1);
        for (int d = 0; d < M.P; ++d) out[hist_index<PASS2>(d, Q)] = in[d % M.P1];
      }
      if (alive && g == 0 && need_r_flag) need_r_flag[i] = 0;
      alive = false;
    } else {
      t0 = padv[i] + 1;
      nbp = breakpoints<FL>(M, th, ld, (double)off, bp);
      if (window_only) nint = min(nint, max(last_window_day(M, off, padv[i]) - padv[i], 1));
    }
  }
  const bool flags_r = alive;  // this chain reports need_r at the end
  const int cdays = ckpt_days(M.P1);
  // states of this lane in the checkpoint tile: q = g * 5 + j, j = 0..3
  double *ck = ckpt ? ckpt + ((i / kTile) * (int64_t)cdays * 10 + g * 5) * kTile + (i % kTile) : nullptr;
  int d0 = 0;
  if constexpr (PASS2) {
    if (ck && alive) {
      double minbp = INFINITY;
      for (int k = 0; k < nbp; ++k) minbp = fmin(minbp, bp[k]);
      const double dd = floor(minbp) - (double)t0;
      d0 = dd > 0 ? (int)fmin(dd, (double)min(cdays - 1, ndays - 1)) : 0;
      if (d0 > 0) {
        const double *in = hist1 + (i * 2 + g) * (int64_t)hist_pitch(M.P1);
        for (int d = 0; d <= d0; ++d) out[hist_index<PASS2>(d, Q)] = in[d];
#pragma unroll
        for (int j = 0; j < NS; ++j) y[j] = ck[((int64_t)d0 * 10 + j) * kTile];
      }
    }
  }
  bool bp_sorted = true;
  for (int k = 1; k < nbp; ++k) bp_sorted = bp_sorted && bp[k - 1] <= bp[k];
  int kseg = -2;
  const double invy = 1.0 / M.Ny, invo = 1.0 / M.No;
  Seg2 sg;
  [[maybe_unused]] SegTab T{};
  if constexpr (PASS2) {
    extern __shared__ double seg_smem[];
    T = seg_tab_of<FL>(seg_smem, blockDim.x, threadIdx.x);
    fill_segment_table<FL>(M, th, ld, bp, nbp, g, invy, invo, T);
    sg = select_segment_tab(T, g, nbp, (double)(t0 + d0), &tcut, &kseg, bp_sorted);
  } else {
    sg = select_segment_age2<FL>(M, th, ld, bp, nbp, (double)(t0 + d0), g, &tcut, invy, invo, &kseg, bp_sorted);
  }

  const int m = M.m;
  const double h = 1.0 / (double)m;
  // what a history row holds: S of this lane's group — age-2x: its incidence (incidence_out2, every lane calls it)
  [[maybe_unused]] const double *bps_inc = bp;
  [[maybe_unused]] int bstride_inc = 1;
  if constexpr (PASS2) { bps_inc = T.bps; bstride_inc = T.np; }
  auto row_value = [&](int d) -> double {
    if constexpr (X2) return incidence_out2<FL>(M, th, ld, bps_inc, bstride_inc, nbp, (double)(t0 + d), L, y, g);
    return y[0];
  };
  {
    const double v0 = row_value(0);
    if (alive && d0 == 0) out[0] = v0;
  }
  if constexpr (!PASS2) {
    if (ck && alive) {
#pragma unroll
      for (int j = 0; j < NS; ++j) ck[j * kTile] = y[j];
    }
  }
  // day range of the warp: every chain walks it, live only inside its own (d0, nint)
  const int d_first = __reduce_min_sync(kFull, alive ? d0 + 1 : 0x7fffffff);
  const int d_end = __reduce_max_sync(kFull, alive ? nint : 0);
  const bool pow2 = (m & (m - 1)) == 0;
  const double hh2reg = 0.5 * h, h6reg = h / 6.0;
  // end-of-day outputs of this lane: its S into the history row, and (pass 1) the checkpoint of its four states
  auto day_out = [&](int d, bool store) {
    const double v = row_value(d);
    if (!store) return;
    out[hist_index<PASS2>(d, Q)] = v;
    if constexpr (!PASS2) {
      if (ck && d < cdays) {
#pragma unroll
        for (int j = 0; j < NS; ++j) ck[((int64_t)d * 10 + j) * kTile] = y[j];
      }
    }
  };
  for (int d = d_first; d < d_end; ++d) {
    bool live = alive && d > d0 && d < nint;
    double tday = (double)(t0 + d - 1);
    // RUNS of usual days — every chain of the warp live, no breakpoint inside, beta >= 0 — cost two votes per run and
    // ONE per day: the run length is the warp minimum of the days each chain has left before its next breakpoint /
    // its last day; the fast steps of a day run back to back (ping-pong between state arrays) and only accumulate
    // their maybe-out-of-bounds flag, voted on at the end of the day; a day some chain raised it in (never in
    // practice) is redone from its first state by the checked routes below.  Step probe (profiles/microbench/
    // step_probe.cu, one warp per scheduler): 228 cycles per step against 265 with the three votes of a checked
    // step; cycle counters in the kernel: the regular day 1,520 -> 1,100 cycles with one vote per day.
    if (pow2) {
      // (lanes without a chain to integrate — behind the batch, invalid data_offset, not in the retry round — tag along)
      int nf = !alive ? 0x7fffffff : (live && sg.fast_ok) ? (int)fmin((double)(nint - d), floor(tcut) - tday) : 0;
      nf = __reduce_min_sync(kFull, nf);
      if (nf == 0x7fffffff) nf = 0;
      if (nf > 0) {
        const bool hold = __all_sync(kFull, sg.hold);
        int j = 0;
        for (; j < nf; ++j) {
          double w[NS];
          const unsigned flag = hold ? rk4_day2w<true, NS>(L, sg, y, w, tday, h, hh2reg, h6reg, m)
                                     : rk4_day2w<false, NS>(L, sg, y, w, tday, h, hh2reg, h6reg, m);
          if (__any_sync(kFull, alive && (int)flag < 0)) break;
#pragma unroll
          for (int i = 0; i < NS; ++i) y[i] = w[i];
          day_out(d + j, alive);
          tday += 1.0;
        }
        d += j;
        if (j == nf) { --d; continue; }  // (the loop header steps to the day behind the run)
        live = alive && d > d0 && d < nint;  // day d of the run raised the flag: the checked route takes it
      }
    }
    const bool regular = pow2 && __all_sync(kFull, !live || tcut >= tday + 1.0);
    double pa = tday;
    bool general = false;
    if (regular && (m & 1) == 0 && __all_sync(kFull, live)) {
      // every chain of the warp integrates this day: sub-steps in pairs, ping-pong between two state arrays
      double z[NS];
      for (int s = 0; s < m; s += 2) {
        const double pm = pa + h, pb = pm + h;
        rk4_step2w_all<NS>(L, sg, y, z, pa, pm, h, hh2reg, h6reg, need_r);
        rk4_step2w_all<NS>(L, sg, z, y, pm, pb, h, hh2reg, h6reg, need_r);
        pa = pb;
      }
    } else if (PASS2 && pow2) {
      // A day some chain of the warp has a breakpoint in (or is not live): the chains walk their pieces together
      // (a lane with a cut sub-step has one piece more and the others wait for it), every piece by the FAST step on
      // the general (linear-segment) instance, flags accumulated, one vote at the end of the day — and the day redone
      // from its first state by the checked loop below if some chain raised the flag or sits on a segment with a
      // negative beta (never on a posterior).  Same pieces, same operations, same bits.  Cycle counters (8,192
      // chains, profiles/r2_summary.md): the checked loop took such a day through five pieces of ~930 cycles (three
      // votes, a division and a re-selection through global memory each) against ~390 for a regular sub-step; 17 %
      // of the days were 40 % of pass 2.
      double y0[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) y0[i] = y[i];
      unsigned flag = 0;
      bool fok = true;
      int s = 0;
      double sb = tday, pl = tday;  // start of the current sub-step, start of the next piece
      for (;;) {
        const bool pending = live && s < m;
        if (!__any_sync(kFull, pending)) break;
        const double b = sb + h;  // (exact for power-of-two m: == tday + (s + 1) * h)
        if (pending && pl >= tcut) sg = select_segment_tab(T, g, nbp, pl, &tcut, &kseg, bp_sorted);
        double pb = b;
        bool done = true;
        if (tcut < b) { pb = tcut; done = false; }  // a breakpoint strictly inside (pl, b): cut there
        const double hh = pending ? pb - pl : 0.0, hh2 = 0.5 * hh, h6 = div6(hh);
        fok = fok && (!pending || sg.fast_ok);
        double yn[NS];
        const unsigned f = rk4_raw2w<false, NS>(L, sg, y, yn, pl, pending ? pb : pl, hh, hh2, h6);
        if (pending) {
#pragma unroll
          for (int i = 0; i < NS; ++i) y[i] = yn[i];
          flag |= f;
          pl = pb;
          if (done) { ++s; sb = b; }
        }
      }
      if (__any_sync(kFull, (int)flag < 0 || !fok)) {
#pragma unroll
        for (int i = 0; i < NS; ++i) y[i] = y0[i];
        kseg = -2;  // (segment index unknown again: full scan)
        sg = select_segment_tab(T, g, nbp, tday, &tcut, &kseg, bp_sorted);
        general = true;
      }
    } else if (regular) {
      for (int s = 0; s < m; ++s) {
        const double pb = pa + h;
        rk4_step2w<NS>(L, sg, y, live, pa, pb, h, hh2reg, h6reg, need_r);
        pa = pb;
      }
    } else {
      general = true;
    }
    if (general) {
      int s = 0;
      for (;;) {
        const bool pending = live && s < m;
        if (!__any_sync(kFull, pending)) break;
        const double b = (s == m - 1) ? tday + 1.0 : tday + (double)(s + 1) * h;
        bool step_done = true;
        double pb = b;
        if constexpr (PASS2) {
          if (pending && pa >= tcut) sg = select_segment_tab(T, g, nbp, pa, &tcut, &kseg, bp_sorted);
          if (tcut < b) { pb = tcut; step_done = false; }  // a breakpoint strictly inside (pa, b): cut there
        }
        const double hh = pending ? pb - pa : 0.0, hh2 = 0.5 * hh, h6 = div6(hh);
        rk4_step2w<NS>(L, sg, y, pending, pa, pending ? pb : pa, hh, hh2, h6, need_r);
        if (pending) {
          pa = pb;
          if (step_done) ++s;
        }
      }
    }
    day_out(d, live);
  }
  // either lane may have met an undecidable R test: the chain is flagged if one of them did
  const unsigned nr = (__ballot_sync(kFull, need_r) >> L.shift) & 3u;
  if (flags_r && g == 0 && need_r_flag) need_r_flag[i] = (nr || force_need_r) ? 1 : 0;
}

// ------------------------------------------------ K_ode, four lanes per chain
// Small batches (the 8,192-chain shards of the strong-scaled BASELINE config, 65,536 chains over 8 GPUs) leave a
// thread-per-chain launch at ~2 warps per SM: every RK4 stage is then one long dependent FP64 chain per warp and the
// pipe idles (round 1: pass 2 0.71 ms thread-per-chain, 0.56 ms with two lanes per chain, for 1/8 of the work that
// takes 0.86 ms at full batch).  Here a chain is spread over FOUR lanes, two per age group:
//     lane (g, 0) holds (S_g, E1_g),   lane (g, 1) holds (E2_g, I_g)
// and a stage costs each lane two shuffles (its partner's second value: I_g resp. E1_g; and I_y, the only coupling
// between the groups, model-agen.cpp:136-144) and ONE instruction stream for both lane kinds:
//     inf = ((bA * I_g) + bB * I_y) * S_g        (bA = bB = 0 on the (E2, I) lanes)
//     w   = half ? u : inf
//     dv  = fma(c, v, w)        c = -a1 | -gamma     ->  inf - a1 E1   |  E2 - gamma I
//     du  = fma(a1h, pv, -w)    a1h = 0 | a1         ->  -inf          |  a1 E1 - E2
// Every operation is the explicitly rounded one of rhs<> (the y group's beta_yo = 0 term adds an exact +0), so the
// results are bit-identical to the thread-per-chain kernel (tests/test_gpu_parity.py).  R is not integrated; a
// chain whose guard decision could have depended on it is flagged for the R-tracking instance of k_ode, as there.
struct Seg4 {
  double tk, bA, sA, bB, sB;
  bool fast_ok, hold;
};

template <int FL>
__device__ __noinline__ Seg4 select_segment4(const DevModel &M, const double *__restrict__ th, int64_t ld, const double *bp, int nbp,
                                             double pa, int is_o, int half, double *tcut, double invy, double invo,
                                             int *kcur, bool sorted) {
  int k = -1;
  double tc = INFINITY;
  if (sorted && *kcur >= -1) {
    k = *kcur;
    while (k + 1 < nbp && bp[k + 1] <= pa) ++k;
    if (k + 1 < nbp) tc = bp[k + 1];
  } else {
    for (int i = 0; i < nbp; ++i) {
      if (bp[i] <= pa) k = i;
      if (bp[i] > pa && bp[i] < tc) tc = bp[i];
    }
  }
  *kcur = k;
  *tcut = tc;
  double by, bo, byo;
  age_beta<FL>(M, th, ld, k < 0 ? 0 : k, by, bo, byo);
  Seg4 sg;
  sg.tk = 0.0; sg.sA = 0.0; sg.sB = 0.0;
  sg.bA = is_o ? bo : by;
  sg.bB = is_o ? byo : 0.0;
  if (k >= 0 && age_linear(M, k)) {
    double ny, no, nyo;
    age_beta<FL>(M, th, ld, k + 1, ny, no, nyo);
    const double len = bp[k + 1] - bp[k];
    sg.tk = bp[k];
    sg.sA = is_o ? (no - bo) / len : (ny - by) / len;
    sg.sB = is_o ? (nyo - byo) / len : 0.0;
    sg.fast_ok = by >= 0 && bo >= 0 && byo >= 0 && ny >= 0 && no >= 0 && nyo >= 0 && len > 0;
    // hold <=> all three slopes of the chain are zero (the same test as select_segment, on the same roundings)
    sg.hold = __dmul_rn((ny - by) / len, invy) == 0.0 && __dmul_rn((no - bo) / len, invo) == 0.0 &&
              __dmul_rn((nyo - byo) / len, invy) == 0.0;
  } else {
    sg.fast_ok = by >= 0 && bo >= 0 && byo >= 0;
    sg.hold = true;
  }
  // the same normalisation (and the same roundings) as select_segment
  const double invA = is_o ? invo : invy;
  sg.bA = __dmul_rn(sg.bA, invA); sg.sA = __dmul_rn(sg.sA, invA);
  sg.bB = __dmul_rn(sg.bB, invy); sg.sB = __dmul_rn(sg.sB, invy);
  if (half) { sg.bA = 0.0; sg.sA = 0.0; sg.bB = 0.0; sg.sB = 0.0; }  // the (E2, I) lanes have no infection term
  return sg;
}

// per-lane constants of the four-lane kernel
struct Lane4 {
  double a1h, c, Ngrp;  // du = fma(a1h, pv, -w), dv = fma(c, v, w)
  int shift;            // first lane of this chain's quad
  int src_partner, src_iy, half;
  unsigned nb, lo_u, r_u;  // fast guard: v in [0, N) via nb = hi(N) - 1; u in [1, +inf) (S) or [0, N) via (lo_u, r_u)
};

// The kernel is WARP-SYNCHRONOUS: all 32 lanes (8 chains) execute every stage together and every shuffle / vote
// names the full warp.  (First version, profiles/r2_summary.md: shuffles over per-quad masks made ptxas emit
// MATCH.ANY + REDUX.OR + a divergence check in front of every RK4 step — 400 cycles per stage at 8,192 chains,
// twice the thread-per-chain kernel.)  Chains that have nothing to do in a piece — finished, not started, fewer
// pieces on a day a breakpoint cuts for a neighbour — run it with their state frozen (`commit` false).

template <bool GUARDED, bool HOLD>
__device__ __forceinline__ void rhs4(const Lane4 &L, const Seg4 &sg, double t, double u, double v, double &du,
                                     double &dv, unsigned &flag) {
  const double pv = __shfl_sync(kFull, v, L.src_partner);  // (S,E1) lane: I_g;  (E2,I) lane: E1_g
  const double iy = __shfl_sync(kFull, v, L.src_iy);       // I of the younger group
  bool bad = false;
  if constexpr (GUARDED) {
    // bounds guard over the chain's states, model-agen.cpp:109-124 (R is not integrated here: see below_one)
    const int bad_own = (u < 0 || v < 0 || u > L.Ngrp || v > L.Ngrp);
    bad = ((__ballot_sync(kFull, bad_own) >> L.shift) & 0xFu) != 0u;
    flag |= (!bad && !L.half && u < 1.0) ? 1u : 0u;  // the R test is undecidable here
  } else {
    const unsigned hu = (unsigned)__double2hiint(u) - L.lo_u, hv = (unsigned)__double2hiint(v);
    flag |= ((L.r_u - hu) | hu) | ((L.nb - hv) | hv);  // bit 31 set <=> maybe out of bounds
  }
  const double dt = HOLD ? 0.0 : t - sg.tk;
  const double bA = HOLD ? sg.bA : fma(dt, sg.sA, sg.bA), bB = HOLD ? sg.bB : fma(dt, sg.sB, sg.bB);
  const double inf = __dmul_rn(fma(bA, pv, __dmul_rn(bB, iy)), u);
  const double w = L.half ? u : inf;
  dv = fma(L.c, v, w);
  du = fma(L.a1h, pv, -w);
  if constexpr (GUARDED) {
    if (bad) { du = 0.0; dv = 0.0; }  // the guard freezes the whole state
  }
}

// returns (fast instance) whether no lane of the chain raised the maybe-out-of-bounds bit
template <bool GUARDED, bool HOLD>
__device__ __forceinline__ bool rk4_try4(const Lane4 &L, const Seg4 &sg, double u, double v, double &un, double &vn,
                                         double pa, double pb, double hh, double hh2, double h6, bool &need_r) {
  const double tm = pa + hh2;
  double ku, kv, au, av, ut, vt;
  unsigned flag = 0;
  rhs4<GUARDED, HOLD>(L, sg, pa, u, v, ku, kv, flag);
  au = ku; av = kv; ut = fma(hh2, ku, u); vt = fma(hh2, kv, v);
  rhs4<GUARDED, HOLD>(L, sg, tm, ut, vt, ku, kv, flag);
  au = fma(2.0, ku, au); av = fma(2.0, kv, av); ut = fma(hh2, ku, u); vt = fma(hh2, kv, v);
  rhs4<GUARDED, HOLD>(L, sg, tm, ut, vt, ku, kv, flag);
  au = fma(2.0, ku, au); av = fma(2.0, kv, av); ut = fma(hh, ku, u); vt = fma(hh, kv, v);
  rhs4<GUARDED, HOLD>(L, sg, pb, ut, vt, ku, kv, flag);
  un = fma(h6, __dadd_rn(au, ku), u);
  vn = fma(h6, __dadd_rn(av, kv), v);
  if constexpr (GUARDED) {
    need_r = (flag & 1u) != 0u;
    return true;
  }
  return ((__ballot_sync(kFull, (int)flag < 0) >> L.shift) & 0xFu) == 0u;
}

// one RK4 piece for all eight chains of the warp; a chain takes the result only when `commit`
__device__ __forceinline__ void rk4_step4(const Lane4 &L, const Seg4 &sg, double &u, double &v, bool commit, double pa,
                                          double pb, double hh, double hh2, double h6, bool &need_r) {
  double un, vn;
  bool nr = false;
  const bool hold = __all_sync(kFull, sg.hold);
  bool ok = hold ? rk4_try4<false, true>(L, sg, u, v, un, vn, pa, pb, hh, hh2, h6, nr)
                 : rk4_try4<false, false>(L, sg, u, v, un, vn, pa, pb, hh, hh2, h6, nr);
  ok = ok && sg.fast_ok;
  if (__any_sync(kFull, commit && !ok)) {
    // rare: the reference's guard evaluated exactly in FP64, by the whole warp; taken by the chains that need it
    double gu, gv;
    bool gnr = false;
    rk4_try4<true, false>(L, sg, u, v, gu, gv, pa, pb, hh, hh2, h6, gnr);
    if (!ok) { un = gu; vn = gv; nr = gnr; }
  }
  if (commit) { u = un; v = vn; need_r = need_r || nr; }
}

// MINB = resident blocks per SM the register allocation must allow: 4 (<= 128 registers, nothing spilled) while the
// whole batch fits at that occupancy anyway, 7 (<= 72 registers) above it
template <int FL, bool PASS2, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_ode_age4(const DevModel M, const double *__restrict__ theta, int64_t ld, int64_t n, const int *__restrict__ offset,
           const int *__restrict__ padv, const double *__restrict__ hist1, double *__restrict__ hist_out,
           double *__restrict__ ckpt, int window_only = 0, int ndays_limit = 0,
           const int *__restrict__ retry_status = nullptr, int *__restrict__ need_r_flag = nullptr,
           int force_need_r = 0) {
  bool need_r = false;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool exists = (tid >> 2) < n;
  const int64_t i = exists ? (tid >> 2) : n - 1;  // lanes behind the batch shadow its last chain and never store
  const int lane = threadIdx.x & 31;
  const int g = (lane >> 1) & 1;  // 0 = younger group, 1 = older group
  Lane4 L;
  L.half = lane & 1;              // 0: (S, E1), 1: (E2, I)
  L.shift = lane & ~3;
  L.src_partner = lane ^ 1;
  L.src_iy = L.shift | 1;
  L.Ngrp = g ? M.No : M.Ny;
  L.nb = (unsigned)__double2hiint(L.Ngrp) - 1u;
  L.lo_u = L.half ? 0u : 0x3ff00000u;
  L.r_u = L.half ? L.nb : 0x40000000u;
  L.a1h = L.half ? M.a1 : 0.0;
  L.c = L.half ? -M.gamma : -M.a1;
  const double *th = theta + i;
  const int ndays = PASS2 ? M.P : M.P1;
  int nint = ndays;  // days integrated here
  bool alive = exists;
  if constexpr (!PASS2) {  // two-round pass 1, see k_ode
    if (retry_status && !(retry_status[i] & kStRetry)) alive = false;
    if (ndays_limit > 0) nint = min(nint, ndays_limit);
  }
  const int pitch = PASS2 ? hist_pitch2(ndays) : hist_pitch(ndays), Q = pitch >> 2;
  double *out = hist_out + (i * 2 + g) * (int64_t)pitch;
  const bool writes_s = !L.half;
  // model-age-groups2q.R:157-158
  double u = L.half ? 0.0 : (g ? M.No : M.Ny - 1.0);
  double v = (L.half || g) ? 0.0 : 1.0;

  double bp[kMaxBp];
  int nbp = 0;
  double tcut = INFINITY;
  int t0 = 1;
  if constexpr (PASS2) {
    const int off = offset[i];
    if (off == EPI_INVALID_DATA_OFFSET) {
      // pass 2 skipped: the pass-1 rows are recycled into the P-length slices (:216-223)
      if (alive && writes_s) {
        const double *in = hist1 + (i * 2 + g) * (int64_t)hist_pitch(M.P1);
        for (int d = 0; d < M.P; ++d) out[hist_index<PASS2>(d, Q)] = in[d % M.P1];
      }
      if (alive && (lane & 3) == 0 && need_r_flag) need_r_flag[i] = 0;
      alive = false;
    } else {
      t0 = padv[i] + 1;
      nbp = breakpoints<FL>(M, th, ld, (double)off, bp);
      if (window_only) nint = min(nint, max(last_window_day(M, off, padv[i]) - padv[i], 1));
    }
  }
  const bool flags_r = alive;  // this chain reports need_r at the end
  const int cdays = ckpt_days(M.P1);
  // states of this lane in the checkpoint tile: q = g * 5 + half * 2 (+1)
  double *ck = ckpt ? ckpt + ((i / kTile) * (int64_t)cdays * 10 + g * 5 + L.half * 2) * kTile + (i % kTile) : nullptr;
  int d0 = 0;
  if constexpr (PASS2) {
    if (ck && alive) {
      double minbp = INFINITY;
      for (int k = 0; k < nbp; ++k) minbp = fmin(minbp, bp[k]);
      const double dd = floor(minbp) - (double)t0;
      d0 = dd > 0 ? (int)fmin(dd, (double)min(cdays - 1, ndays - 1)) : 0;
      if (d0 > 0) {
        if (writes_s) {
          const double *in = hist1 + (i * 2 + g) * (int64_t)hist_pitch(M.P1);
          for (int d = 0; d <= d0; ++d) out[hist_index<PASS2>(d, Q)] = in[d];
        }
        u = ck[((int64_t)d0 * 10) * kTile];
        v = ck[((int64_t)d0 * 10 + 1) * kTile];
      }
    }
  }
  bool bp_sorted = true;
  for (int k = 1; k < nbp; ++k) bp_sorted = bp_sorted && bp[k - 1] <= bp[k];
  int kseg = -2;
  const double invy = 1.0 / M.Ny, invo = 1.0 / M.No;
  Seg4 sg = select_segment4<FL>(M, th, ld, bp, nbp, (double)(t0 + d0), g, L.half, &tcut, invy, invo, &kseg, bp_sorted);

  const int m = M.m;
  const double h = 1.0 / (double)m;
  if (alive && d0 == 0 && writes_s) out[0] = u;
  if constexpr (!PASS2) {
    if (ck && alive) { ck[0] = u; ck[kTile] = v; }
  }
  // day range of the warp: every chain walks it, live only inside its own (d0, nint)
  const int d_first = __reduce_min_sync(kFull, alive ? d0 + 1 : 0x7fffffff);
  const int d_end = __reduce_max_sync(kFull, alive ? nint : 0);
  // the same route choices as k_ode (all of them give the same bits): a day without a breakpoint in any live lane
  // takes the constant-step loop, the others the general piece loop
  const bool pow2 = (m & (m - 1)) == 0;
  const double hh2reg = 0.5 * h, h6reg = h / 6.0;
  for (int d = d_first; d < d_end; ++d) {
    const bool live = alive && d > d0 && d < nint;
    const double tday = (double)(t0 + d - 1);
    const bool regular = pow2 && __all_sync(kFull, !live || tcut >= tday + 1.0);
    double pa = tday;
    if (regular) {
      for (int s = 0; s < m; ++s) {
        const double pb = pa + h;
        rk4_step4(L, sg, u, v, live, pa, pb, h, hh2reg, h6reg, need_r);
        pa = pb;
      }
    } else {
      int s = 0;
      for (;;) {
        const bool pending = live && s < m;
        if (!__any_sync(kFull, pending)) break;
        const double b = (s == m - 1) ? tday + 1.0 : tday + (double)(s + 1) * h;
        bool step_done = true;
        double pb = b;
        if constexpr (PASS2) {
          if (pending && pa >= tcut)
            sg = select_segment4<FL>(M, th, ld, bp, nbp, pa, g, L.half, &tcut, invy, invo, &kseg, bp_sorted);
          if (tcut < b) { pb = tcut; step_done = false; }  // a breakpoint strictly inside (pa, b): cut there
        }
        const double hh = pending ? pb - pa : 0.0, hh2 = 0.5 * hh, h6 = div6(hh);
        rk4_step4(L, sg, u, v, pending, pa, pending ? pb : pa, hh, hh2, h6, need_r);
        if (pending) {
          pa = pb;
          if (step_done) ++s;
        }
      }
    }
    if (live && writes_s) out[hist_index<PASS2>(d, Q)] = u;
    if constexpr (!PASS2) {
      if (ck && live && d < cdays) { ck[((int64_t)d * 10) * kTile] = u; ck[((int64_t)d * 10 + 1) * kTile] = v; }
    }
  }
  // any lane may have met an undecidable R test: the chain is flagged if one of them did
  const unsigned nr = (__ballot_sync(kFull, need_r) >> L.shift) & 0xFu;
  if (flags_r && (lane & 3) == 0 && need_r_flag) need_r_flag[i] = (nr || force_need_r) ? 1 : 0;
}

// ---------------------------------------------------------------------- K_mid
// Per-warp shared-memory carve-up shared by k_mid and k_score.
constexpr int kCdfStride = EPI_MAX_TAPS + 1;  // n + 1 cell edges per profile
struct WarpSmem {
  double *theta;   // EPI_MAX_PARAMS
  double *pp;      // NK * 4: per-profile (x0, shape|mean, scale|sd, lgamma(shape))
  double *cdf;     // NK * kCdfStride
  double *W;       // NK * EPI_MAX_TAPS
  double *S;       // NG * (Traits<FL>::kPad + hist_pitch(ndays))
  uint64_t *mbar;  // completion barrier of the bulk history copy
};

template <int FL>
__host__ __device__ constexpr int warp_smem_doubles(int ndays) {
  // even (16-byte aligned sub-regions); the last two doubles hold the warp's mbarrier
  return (EPI_MAX_PARAMS + Traits<FL>::NK * 4 + Traits<FL>::NK * kCdfStride + Traits<FL>::NK * EPI_MAX_TAPS +
          Traits<FL>::NG * (Traits<FL>::kPad + hist_pitch(ndays)) + 3) & ~1;
}

template <int FL>
__device__ __forceinline__ WarpSmem carve(double *base, int ndays) {
  WarpSmem w;
  w.theta = base;
  w.pp = w.theta + EPI_MAX_PARAMS;
  w.cdf = w.pp + Traits<FL>::NK * 4;
  w.W = w.cdf + Traits<FL>::NK * kCdfStride;
  w.S = w.W + Traits<FL>::NK * EPI_MAX_TAPS;
  w.mbar = reinterpret_cast<uint64_t *>(base + warp_smem_doubles<FL>(ndays) - 2);
  return w;
}

// latency (mean, sd) of profile q in MODEL space; age order: ycase, yhosp,
// ydied, ocase, ohosp, odied (model-age-groups2q.R:106-111); simple: hosp, died
// with the NEGATED latency and sd 5 (model-simple-ode-drj-cmp.R:56-57).
template <int FL>
__device__ __forceinline__ void profile_params(const double *th, int q, double &mean, double &sd) {
  if constexpr (Traits<FL>::k2i) {
    // model-age-groups2i.R:62-75,96-101 (CLsd is a parameter in this generation)
    const double CLsd = th[21], HLsd = th[22], DLsd = th[23];
    switch (q) {
      case 0: mean = th[10]; sd = CLsd; break;
      case 1: mean = th[12]; sd = HLsd; break;
      case 2: mean = th[13]; sd = DLsd; break;
      case 3: mean = th[15]; sd = CLsd; break;
      case 4: mean = th[17]; sd = HLsd; break;
      default: mean = th[18]; sd = DLsd; break;
    }
  } else if constexpr (Traits<FL>::kAge) {
    // CLsd: 5 in model-age-groups2q.R:103, 2.5 in model-age-groups2x.R:20; y.CL / o.CL are parameters 45, 46 resp. 47, 48
    const double CLsd = Traits<FL>::k2x ? 2.5 : 5.0, HLsd = th[19], DLsd = th[20];
    constexpr int iycl = Traits<FL>::k2x ? 46 : 43, iocl = Traits<FL>::k2x ? 47 : 44;
    switch (q) {
      case 0: mean = th[iycl]; sd = CLsd; break;
      case 1: mean = th[11]; sd = HLsd; break;
      case 2: mean = th[12]; sd = DLsd; break;
      case 3: mean = th[iocl]; sd = CLsd; break;
      case 4: mean = th[15]; sd = HLsd; break;
      default: mean = th[16]; sd = DLsd; break;
    }
  } else {
    mean = -(q == 0 ? th[5] : th[6]);
    sd = 5.0;
  }
}

// Profile extent in delay form: taps o = 0..n-1 act on x[t - dmin - o].
// Returns false when the profile is wider than EPI_MAX_TAPS (or NaN input).
template <int FL>
__device__ __forceinline__ bool profile_extent(double mean, double sd, int &dmin, int &n, double &x0) {
  if constexpr (Traits<FL>::kAge) {
    // calcGammaProfile, model-age-groups2q.R:22-23
    const double kb = fmax(0.0, ceil(mean - sd * 3));
    const double ke = fmax(kb + 1, floor(mean + sd * 3));
    if (!(ke - kb + 1 <= EPI_MAX_TAPS) || !(kb >= 0) || !(ke < 1e6)) return false;
    dmin = (int)kb; n = (int)ke - (int)kb + 1; x0 = kb - 0.5;
  } else {
    // calcConvolveProfile, model-simple-ode-drj-cmp.R:17-18 (latency < 0)
    const double ke = fmin(0.0, ceil(mean + sd * 1.5));
    const double kb = fmin(ke - 1, floor(mean - sd * 1.5));
    if (!(ke - kb + 1 <= EPI_MAX_TAPS) || !(kb > -1e6)) return false;
    dmin = -(int)ke; n = (int)ke - (int)kb + 1; x0 = kb - 0.5;
  }
  return true;
}

// Build all NK latency profiles of one chain into W[q*EPI_MAX_TAPS + o] (tap order o = 0..n-1).
// The CDF is needed at the n_q + 1 cell edges x0_q + i of every profile; all of them (~200 for
// the age model) form ONE flat work list spread over the lanes, so that a round of 32 lanes is
// always full (ncu, profiles/r1_v1: one profile at a time left the second round of each profile
// with 1-6 active lanes, and six inlined copies of the incomplete-gamma code thrashed the
// instruction cache: 43 % of the stalls were "no instruction").
template <int FL>
__device__ __forceinline__ void build_profiles(const WarpSmem &w, const int *n_, int lane) {
  constexpr int NK = Traits<FL>::NK;
  if constexpr (Traits<FL>::kAge) {
    if (lane < NK) w.pp[lane * 4 + 3] = lgamma(w.pp[lane * 4 + 1]);  // once per profile
    __syncwarp();
    // split every profile's edge list at x = (shape + 1) * scale: points below go to the
    // power-series list, the rest to the continued-fraction list
    int ns[NK], offA[NK + 1], offB[NK + 1];
    offA[0] = offB[0] = 0;
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      const double xs = (w.pp[q * 4 + 1] + 1.0) * w.pp[q * 4 + 2] - w.pp[q * 4 + 0];  // first i with x0 + i >= (a+1)*scale
      int k = (xs > 0.0) ? (int)ceil(xs) : 0;
      ns[q] = min(max(k, 0), n_[q] + 1);
      offA[q + 1] = offA[q] + ns[q];
      offB[q + 1] = offB[q] + (n_[q] + 1 - ns[q]);
    }
#pragma unroll
    for (int method = 0; method < 2; ++method) {
      const int *off = method ? offB : offA;
      for (int idx = lane; idx < off[NK]; idx += 32) {
        int q = 0, base = 0;
#pragma unroll
        for (int k = 1; k < NK; ++k) { const bool ge = idx >= off[k]; q += ge; base = ge ? off[k] : base; }
        int nsq = 0;
#pragma unroll
        for (int k = 0; k < NK; ++k) nsq = (k == q) ? ns[k] : nsq;
        const int i = idx - base + (method ? nsq : 0);
        const double x = w.pp[q * 4 + 0] + (double)i;
        const double c = method ? dev_pgamma_m<1>(x, w.pp[q * 4 + 1], w.pp[q * 4 + 2], w.pp[q * 4 + 3])
                                : dev_pgamma_m<0>(x, w.pp[q * 4 + 1], w.pp[q * 4 + 2], w.pp[q * 4 + 3]);
        w.cdf[q * kCdfStride + i] = c;
      }
    }
  } else {
    int off[NK + 1];
    off[0] = 0;
#pragma unroll
    for (int q = 0; q < NK; ++q) off[q + 1] = off[q] + n_[q] + 1;
    for (int idx = lane; idx < off[NK]; idx += 32) {
      int q = 0, base = 0;
#pragma unroll
      for (int k = 1; k < NK; ++k) { const bool ge = idx >= off[k]; q += ge; base = ge ? off[k] : base; }
      const int i = idx - base;
      w.cdf[q * kCdfStride + i] = dev_pnorm(w.pp[q * 4 + 0] + (double)i, w.pp[q * 4 + 1], w.pp[q * 4 + 2]);
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < NK; ++q) {
    const int n = n_[q];
    const double *cdf = w.cdf + q * kCdfStride;
    // sum(values) (:36) telescopes to F(kb - .5) - F(ke + .5); the explicit sum differs from it only by the
    // rounding of the individual differences (<= n ulp/2), well inside the 1e-13 agreement of the weights
    const double sum = cdf[0] - cdf[n];
    for (int o = lane; o < n; o += 32) {
      // age: values are reversed (rev(), :37): tap o <-> k = ke - o <-> cell n-1-o
      const int cell = Traits<FL>::kAge ? (n - 1 - o) : o;
      w.W[q * EPI_MAX_TAPS + o] = (cdf[cell] - cdf[cell + 1]) / sum;
    }
  }
  __syncwarp();
}

// stage one chain's S history (ndays values per group) behind Traits<FL>::kPad prefill slots: issue the
// bulk copies (one elected lane), fill the prefill slots meanwhile, then wait for the bytes
template <int FL>
__device__ __forceinline__ void stage_history_begin(double *S, uint64_t *mbar, const double *__restrict__ hist,
                                                    int64_t chain, int ndays, int lane) {
  constexpr int NG = Traits<FL>::NG;
  constexpr int PAD = Traits<FL>::kPad;
  const int pitch = hist_pitch(ndays), stride = PAD + pitch;
  if (lane == 0) mbar_init(mbar, 1);
  __syncwarp();
  if (lane == 0) {
    const unsigned bytes = (unsigned)pitch * 8u;
    mbar_expect_tx(mbar, bytes * NG);
#pragma unroll
    for (int g = 0; g < NG; ++g) bulk_g2s(S + g * stride + PAD, hist + (chain * NG + g) * (int64_t)pitch, bytes, mbar);
  }
}
template <int FL>
__device__ __forceinline__ void stage_history_end(const DevModel &M, double *S, uint64_t *mbar, int ndays, int lane) {
  constexpr int NG = Traits<FL>::NG;
  constexpr int PAD = Traits<FL>::kPad;
  const int stride = PAD + hist_pitch(ndays);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    // (age-2x: the staged rows hold the incidence, pre-filled with 0, model-age-groups2x.R:156-157)
    const double s0 = Traits<FL>::k2x ? 0.0 : Traits<FL>::kAge ? (g == 0 ? M.Ny - 1.0 : M.No) : M.N - 1.0;
    for (int j = lane; j < PAD; j += 32) S[g * stride + j] = s0;
  }
  mbar_wait(mbar, 0);
  __syncwarp();
}

// convolute(), model-age-groups2q.R:42-45, at padded day index j (sim time t =
// pad+1+j): sum_o f[o] * x[t - dmin - o]; NA (NaN) while t - dmin < n.
template <int PAD>
__device__ __forceinline__ double conv_at(const double *__restrict__ Sg, const double *__restrict__ f, int pad,
                                          int j, int dmin, int n) {
  const int t = pad + 1 + j;
  if (t - dmin < n) return NAN;
  const double *x = Sg + PAD + j - dmin;  // x[t - dmin]
  double z = 0.0;
#pragma unroll 4
  for (int o = 0; o < n; ++o) z = fma(f[o], x[-o], z);  // sequential order as in stats::filter; loads hoisted
  return z;
}

template <int FL>
__global__ void __launch_bounds__(128)
k_mid(const DevModel M, const double *__restrict__ theta, int64_t ld, int64_t n, const double *__restrict__ hist1,
      int *__restrict__ offset, int *__restrict__ padv, int *__restrict__ status, double *__restrict__ Wg,
      int2 *__restrict__ pmeta, int model_space, int ndays_search, int only_flagged) {
  constexpr int NK = Traits<FL>::NK, NG = Traits<FL>::NG;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (chain >= n) return;
  if (only_flagged && !(status[chain] & kStRetry)) return;  // retry round: only the chains the first round flagged
  const int P1 = M.P1;
  const int nsearch = ndays_search > 0 ? min(P1, ndays_search) : P1;  // days of pass 1 that exist in this round
  const WarpSmem w = carve<FL>(smem + (size_t)wib * warp_smem_doubles<FL>(P1), P1);
  stage_history_begin<FL>(w.S, w.mbar, hist1, chain, P1, lane);  // asynchronous; consumed after the profiles

  // theta -> MODEL space in shared memory (transformParams: simple exp(logHR), :136-142)
  int oob = 0;
  for (int p = lane; p < M.d; p += 32) {
    double v = theta[(int64_t)p * ld + chain];
    if (!model_space) {
      oob |= !(v >= M.box_min[p] && v <= M.box_max[p]);
      if (!Traits<FL>::kAge && p == 4) v = exp(v);
    }
    w.theta[p] = v;
  }
  oob = __any_sync(0xffffffffu, oob);
  __syncwarp();

  // per-profile CDF parameters (x0, shape | mean, scale | sd): one lane per profile, side by side
  if (lane < NK) {
    double mean, sd, x0;
    profile_params<FL>(w.theta, lane, mean, sd);
    int dmin = 0, nt = 1;
    profile_extent<FL>(mean, sd, dmin, nt, x0);
    w.pp[lane * 4 + 0] = x0;
    if constexpr (Traits<FL>::kAge) { const double scale = sd * sd / mean; w.pp[lane * 4 + 2] = scale; w.pp[lane * 4 + 1] = mean / scale; }
    else { w.pp[lane * 4 + 1] = mean; w.pp[lane * 4 + 2] = sd; }
  }
  // profiles + padding (:106-114 / simple :56-59)
  int st = oob ? EPI_ST_OUT_OF_BOX : 0;
  int pad = 0;
  int dmin_[NK], n_[NK];
  bool ok = true;
#pragma unroll
  for (int q = 0; q < NK; ++q) {
    double mean, sd, x0;
    profile_params<FL>(w.theta, q, mean, sd);
    int dmin = 0, nt = 1;
    const bool okq = profile_extent<FL>(mean, sd, dmin, nt, x0);
    ok = ok && okq;
    dmin_[q] = dmin; n_[q] = nt;
    // padding counts the case & died kernels only (age, :113-114); both kernels for simple (:59)
    const bool in_pad = Traits<FL>::kAge ? (q != 1 && q != 4) : true;
    if (in_pad && okq) pad = max(pad, dmin + nt - 1);
  }
  pad += 1;
  constexpr int PAD = Traits<FL>::kPad;
  if (!ok || pad > PAD) {
    if (lane == 0) { offset[chain] = EPI_INVALID_DATA_OFFSET; padv[chain] = 0; status[chain] = st | EPI_ST_UNSUPPORTED; }
    mbar_wait(w.mbar, 0);  // never leave with a bulk copy into this block's shared memory in flight
    return;
  }
  __syncwarp();
  build_profiles<FL>(w, n_, lane);
  // profiles to HBM already in the absolute-delay layout k_score convolves with,
  // Wabs[group][delay 0..kPad-1][KG] (column = profile within the group, zero outside its support; a block of
  // four delays is 4 KG consecutive doubles), so that the reader can take the whole table of a chain with one
  // bulk copy
  {
    constexpr int KG = NK / NG;
    int *meta = reinterpret_cast<int *>(w.cdf);  // the CDF scratch is free again: (dmin, n) per profile
#pragma unroll
    for (int q = 0; q < NK; ++q)
      if (lane == 0) { meta[2 * q] = dmin_[q]; meta[2 * q + 1] = n_[q]; }
    __syncwarp();
    double *dst = Wg + (int64_t)chain * NG * PAD * KG;
    for (int i = lane; i < NG * PAD * KG; i += 32) {
      const int grp = i / (PAD * KG), rem = i - grp * (PAD * KG), dly = rem / KG, c = rem - dly * KG, q = grp * KG + c;
      const unsigned o = (unsigned)(dly - meta[2 * q]);
      dst[i] = (o < (unsigned)meta[2 * q + 1]) ? w.W[q * EPI_MAX_TAPS + o] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < NK; ++q)
      if (lane == 0) pmeta[(int64_t)chain * NK + q] = make_int2(dmin_[q], n_[q]);
  }

  // pass-1 threshold search (:167-182 / simple :88-95)
  stage_history_end<FL>(M, w.S, w.mbar, P1, lane);
  const int stride = PAD + hist_pitch(P1);
  int first = -1;
  double thr;  // t0_morts (2q :69, 2i :71) / mort_lockdown_threshold
  if constexpr (Traits<FL>::k2i) thr = w.theta[19]; else if constexpr (Traits<FL>::kAge) thr = w.theta[17]; else thr = w.theta[7];
  if constexpr (Traits<FL>::k2x) {
    // model-age-groups2x.R:194-204: died = cumsum(conv(y.i) * y.died_rate) + cumsum(conv(o.i) * o.died_rate), first day
    // above t0_morts.  The daily terms are computed lane-parallel into the (free) CDF scratch, the two running sums
    // then sequentially in double-double (R's cumsum accumulates in long double and rounds every partial sum to
    // double; a NaN term makes every later sum NaN: no crossing)
    double *dy_ = w.cdf, *do_ = w.cdf + 64;
    for (int j = lane; j < nsearch; j += 32) {
      dy_[j] = conv_at<PAD>(w.S, w.W + 2 * EPI_MAX_TAPS, pad, j, dmin_[2], n_[2]) * M.ifr_y1;
      do_[j] = conv_at<PAD>(w.S + stride, w.W + 5 * EPI_MAX_TAPS, pad, j, dmin_[5], n_[5]) * M.ifr_o1;
    }
    __syncwarp();
    double hy = 0.0, ly = 0.0, ho = 0.0, lo_ = 0.0;
    for (int j = 0; j < nsearch && first < 0; ++j) {
      // two-sum of the running (hi, lo) pair with the next term
      double x = dy_[j], s1 = hy + x, bb = s1 - hy, e = (hy - (s1 - bb)) + (x - bb);
      e += ly; hy = s1 + e; ly = e - (hy - s1);
      x = do_[j]; s1 = ho + x; bb = s1 - ho; e = (ho - (s1 - bb)) + (x - bb);
      e += lo_; ho = s1 + e; lo_ = e - (ho - s1);
      if (hy + ho > thr) first = j;
    }
  }
  for (int base = 0; !Traits<FL>::k2x && base < nsearch && first < 0; base += 32) {
    const int j = base + lane;
    bool hit = false;
    if (j < nsearch) {
      double died;
      if constexpr (Traits<FL>::kAge) {
        const double cy = conv_at<PAD>(w.S, w.W + 2 * EPI_MAX_TAPS, pad, j, dmin_[2], n_[2]);
        const double co = conv_at<PAD>(w.S + stride, w.W + 5 * EPI_MAX_TAPS, pad, j, dmin_[5], n_[5]);
        died = (M.Ny - cy) * M.ifr_y1 + (M.No - co) * M.ifr_o1;
      } else {
        const double c = conv_at<PAD>(w.S, w.W + 1 * EPI_MAX_TAPS, pad, j, dmin_[1], n_[1]);
        died = (M.N - c) * 0.007;
      }
      hit = died > thr;
    }
    const unsigned b = __ballot_sync(0xffffffffu, hit);
    if (b) first = base + __ffs(b) - 1;
  }
  int off = EPI_INVALID_DATA_OFFSET;
  if (first >= 0) off = pad + 1 + first - M.lockdown_offset;
  if (!Traits<FL>::kAge && thr < 0) off = 1 - M.lockdown_offset;  // died[1] = 0 > threshold (simple prefill is 0, :70)
  if (off == EPI_INVALID_DATA_OFFSET) st |= EPI_ST_INVALID_OFFSET;
  if (off == EPI_INVALID_DATA_OFFSET && nsearch < P1) st |= kStRetry;  // not crossed in the first chunk: redo over P1 days
  if (lane == 0) { offset[chain] = off; padv[chain] = pad; status[chain] = st; }
  (void)NG;
}

// -------------------------------------------------------------------- K_score

__device__ __forceinline__ double prior_term(const PriorTerm &p, const double *th) {
  const double A = th[p.ia], B = p.ib >= 0 ? th[p.ib] : 0.0;
  double v;
  if (p.mode == 0) v = p.c0 + p.ca * A + p.cb * B;
  else if (p.mode == 1) v = A * B;
  else v = A / B;
  // dnorm / dlnorm with log(sd) taken from the table: one library log per log-normal term, none per normal term
  if (!p.is_ln) {
    const double z = (v - p.mean) / p.sd;
    return -(kLnSqrt2Pi + 0.5 * z * z + p.log_sd);
  }
  if (v != v) return v;
  if (v <= 0) return -INFINITY;
  const double lx = log(v), y = (lx - p.mean) / p.sd;
  return -(kLnSqrt2Pi + 0.5 * y * y + (lx + p.log_sd));
}

__device__ __forceinline__ void window(int offset, int n, int T, int &dstart) {
  // model-age-groups2q.R:497-512
  int ds = offset, de = offset + n - 1;
  if (de > T) { de = T; ds = de - n + 1; }
  if (ds < 1) ds = 1;
  dstart = ds;
}

// rare: a day with terms of several sizes (the last-8-days hospital weight, :552-554): slots 1..K-1 of that day.
// Kept out of line: inlined into every series of every day it was most of k_score's code (instruction cache).
__device__ __noinline__ double score_more_slots(const Slot *__restrict__ day, int K, const double2 *__restrict__ ltab, double mu) {
  double acc = 0.0;
  const double lmu = tlog_pos(mu, ltab);
  for (int k = 1; k < K; ++k) {
    const Slot z = ldg_slot(day + k);
    if (z.n == 0.0f) continue;
    const double x = (double)z.x;
    acc += x * lmu - fma((double)z.n, z.size, x) * tlog_pos(z.size + mu, ltab);
  }
  return acc;
}

// score the slots of scored series s that read model day t (value v)
// One scored series at model day t: x*log(mu) - (x + n*size)*log(size + mu) summed over the slots of
// that day (= size*log(size/(size+mu)) + x*log(mu/(size+mu)) minus its data-only part size*log(size)).
// `first` is slot 0, already loaded by the caller together with the other series' (ILP).
__device__ __forceinline__ double score_slots(const DevModel &M, const double2 *__restrict__ ltab, int s, int r, bool in_range,
                                              const Slot &first, double v) {
  // Branch-free up to the rare multi-slot days: the two or three series a lane scores per day then sit in ONE basic
  // block and their logarithm chains overlap (cycle counters: scoring was 60 % of k_score, 25 % of its stall samples
  // fixed-latency waits).  A day outside the window or an empty slot computes the two logarithms of a harmless mean
  // and discards them; tlog_pos() of a NaN / inf mean returns garbage that the NA select below replaces.
  const double mu = dev_pmax(M.floor_[s], v);
  const double x = (double)first.x;
  const double a = x * tlog_pos(mu, ltab) - fma((double)first.n, first.size, x) * tlog_pos(first.size + mu, ltab);
  double acc = (in_range && first.n != 0.0f) ? a : 0.0;
  acc = (!in_range || mu < 1e300) ? acc : NAN;  // NA (or overflow) in the model series: the reference's sum is NA
  if (in_range && r >= M.multi_begin[s]) acc += score_more_slots(M.slots[s] + (size_t)r * M.K[s], M.K[s], ltab, mu);
  return acc;
}

// all scored series of one model day: slot 0 of every series is fetched (DaySlots::load, issued
// by the caller BEFORE it computes the day's model values) and only then evaluated
template <int NS>
struct DaySlots {
  Slot first[NS];
  int r[NS];
  bool ok[NS];
  __device__ __forceinline__ void load(const DevModel &M, int t, const int *dstart, bool live) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      r[s] = t - dstart[s];
      ok[s] = live && r[s] >= 0 && r[s] < M.n_obs[s];
      first[s] = ok[s] ? ldg_slot(M.slots[s] + (size_t)r[s] * M.K[s]) : Slot{0.0, 0.0f, 0.0f};
    }
  }
  __device__ __forceinline__ void score(const DevModel &M, const double2 *__restrict__ ltab, const double *v, double *acc) const {
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] += score_slots(M, ltab, s, r[s], ok[s], first[s], v[s]);
  }
};

// ---- shared-memory carve-up of k_score / k_series (per warp) -------------------
// theta[EPI_MAX_PARAMS] | Wabs[NG][kPad][KG] | S[NG][4 phases][kPad/4 + Q]
// Wabs holds the KG = NK/NG latency profiles of one S group indexed by ABSOLUTE delay
// (zero outside a profile's support), so that one sweep over the delay feeds KG
// independent FMA chains per day from a single S value.
// S is kept in the four-phase layout of hist2 (hist_index): phase row r of a group holds the padded-index
// elements e = 4 a + r, a = 0 .. kPad/4 + Q - 1 (the first kPad/4 entries are the prefill x[1..padding], :118,127).
template <int FL>
__host__ __device__ constexpr int score_smem_doubles(int ndays) {
  // even, so that every warp's regions stay 16-byte aligned for the double2 weight loads and the bulk copies
  return (EPI_MAX_PARAMS + Traits<FL>::NG * Traits<FL>::kPad * (Traits<FL>::NK / Traits<FL>::NG) +
          Traits<FL>::NG * (Traits<FL>::kPad + hist_pitch2(ndays)) + 3) & ~1;
}

struct ScoreSmem {
  double *theta, *Wabs, *S;
  uint64_t *mbar;
};

template <int FL>
__device__ __forceinline__ ScoreSmem carve_score(double *base, int ndays) {
  ScoreSmem w;
  w.theta = base;
  w.Wabs = base + EPI_MAX_PARAMS;
  w.S = w.Wabs + Traits<FL>::NG * Traits<FL>::kPad * (Traits<FL>::NK / Traits<FL>::NG);
  w.mbar = reinterpret_cast<uint64_t *>(base + score_smem_doubles<FL>(ndays) - 2);
  return w;
}

// Stage theta (MODEL space), the profiles in absolute-delay form and the S history.
// Returns per-group sweep bounds [lo, hi] and per-profile (dmin, n).
template <int FL>
__device__ __forceinline__ void stage_score(const DevModel &M, const ScoreSmem &w, const double *__restrict__ theta,
                                            int64_t ld, int64_t chain, bool to_model_space,
                                            const double *__restrict__ hist2, const double *__restrict__ Wg,
                                            const int2 *__restrict__ pmeta, int P, int lane, int *dmin_, int *n_,
                                            int *lo, int *hi) {
  constexpr int NK = Traits<FL>::NK, NG = Traits<FL>::NG, KG = NK / NG, PAD = Traits<FL>::kPad, PADQ = PAD / 4;
  const int pitch = hist_pitch2(P), Q = pitch >> 2, AS = PADQ + Q;  // AS: phase-row stride in shared memory
  // asynchronous bulk copies (TMA unit): the four phase rows of every S group and the profile table, one mbarrier
  {
    if (lane == 0) mbar_init(w.mbar, 1);
    __syncwarp();
    if (lane == 0) {
      const unsigned rb = (unsigned)Q * 8u, wb = (unsigned)(NG * PAD * KG * 8);
      mbar_expect_tx(w.mbar, rb * 4 * NG + wb);
#pragma unroll
      for (int r = 0; r < 4 * NG; ++r)  // r = group * 4 + phase: rows are consecutive in HBM and AS apart here
        bulk_g2s(w.S + r * AS + PADQ, hist2 + chain * NG * (int64_t)pitch + (int64_t)r * Q, rb, w.mbar);
      bulk_g2s(w.Wabs, Wg + (int64_t)chain * NG * PAD * KG, wb, w.mbar);
    }
  }
  for (int p = lane; p < M.d; p += 32) {
    double v = theta[(int64_t)p * ld + chain];
    if (to_model_space && !Traits<FL>::kAge && p == 4) v = exp(v);  // transformParams, :136-142
    w.theta[p] = v;
  }
  __syncwarp();
#pragma unroll
  for (int g = 0; g < NG; ++g) { lo[g] = PAD; hi[g] = 0; }
#pragma unroll
  for (int q = 0; q < NK; ++q) {
    const int2 pm = pmeta[(int64_t)chain * NK + q];
    dmin_[q] = pm.x; n_[q] = pm.y;
    const int g = q / KG;
    lo[g] = min(lo[g], pm.x);
    hi[g] = max(hi[g], pm.x + pm.y - 1);
  }
  // prefill slots x[1..padding] = N - Initial (:118,127) in front of every phase row
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const double s0 = Traits<FL>::k2x ? 0.0 : Traits<FL>::kAge ? (g == 0 ? M.Ny - 1.0 : M.No) : M.N - 1.0;
    for (int j = lane; j < PAD; j += 32) w.S[(g * 4 + (j & 3)) * AS + (j >> 2)] = s0;
  }
  __syncwarp();
}
// ... and wait for the bulk copies stage_score issued (the caller does what needs only theta in between: cycle
// counters in k_score showed 5.5 k cycles per chain in this wait and 5 k in the prior that followed it)
__device__ __forceinline__ void stage_score_wait(const ScoreSmem &w) {
  mbar_wait(w.mbar, 0);
  __syncwarp();
}

// S at padded day index j (sim time padding + 1 + j) of group g in the staged phase layout
template <int PAD>
__device__ __forceinline__ double staged_s(const double *__restrict__ S, int g, int AS, int j) {
  return S[(g * 4 + (j & 3)) * AS + PAD / 4 + (j >> 2)];
}

// convolute() (model-age-groups2q.R:42-45) of the KG profiles of one S group for the FOUR consecutive padded day
// indices 4 G .. 4 G + 3 of this lane, in one sweep over the delay:
//   conv_q(j) = sum_d Wabs[d][q] * S[j - d]          acc[r * KG + q] = conv_q(4 G + r)
// The delays run in blocks of four (blk_lo .. blk_hi, delay = 4 blk + u; the table is zero outside a profile's
// support, and fma(+0, x, acc) == acc).  Within a block the sixteen (day, delay) pairs of a lane read only SEVEN
// distinct S values  v[k] = S[4 G + 3 - 4 blk - k], k = 0..6,  three of which are the previous block's last three:
// four new S loads (one per phase row, consecutive lanes -> consecutive addresses, no bank conflict) and 4 KG
// broadcast weights feed 16 KG DFMA — against 4 loads per 6 DFMA when a lane's days were 32 apart.
// The FMA chain of one (day, profile) still runs over the delays in ascending order.
template <int KG, int PAD>
__device__ __forceinline__ void conv_tile4(const double *__restrict__ Sg, int AS, const double *__restrict__ Wabs,
                                           int blk_lo, int blk_hi, int G, double *acc) {
#pragma unroll
  for (int i = 0; i < 4 * KG; ++i) acc[i] = 0.0;
  const double *p = Sg + PAD / 4 + G - blk_lo;  // phase-0 row, element a0 = PAD/4 + G - blk
  const double *wp = Wabs + blk_lo * 4 * KG;
  double v0 = p[3 * AS], v1 = p[2 * AS], v2 = p[AS];
  auto block = [&](const double *pb, const double *wb, double a0, double a1, double a2, double &b4, double &b5, double &b6) {
    const double v3 = pb[0];
    b4 = pb[3 * AS - 1]; b5 = pb[2 * AS - 1]; b6 = pb[AS - 1];
    double w[4 * KG];
#pragma unroll
    for (int i = 0; i < 2 * KG; ++i) {
      const double2 t = *reinterpret_cast<const double2 *>(wb + 2 * i);
      w[2 * i] = t.x; w[2 * i + 1] = t.y;
    }
    const double v[7] = {a0, a1, a2, v3, b4, b5, b6};
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < KG; ++q) acc[r * KG + q] = fma(w[u * KG + q], v[3 - r + u], acc[r * KG + q]);
  };
  int blk = blk_lo;
  // two blocks per trip: the second one continues from the first one's window, no register moves in between
#pragma unroll 1
  for (; blk + 1 <= blk_hi; blk += 2) {
    double t4, t5, t6;
    block(p, wp, v0, v1, v2, t4, t5, t6);
    block(p - 1, wp + 4 * KG, t4, t5, t6, v0, v1, v2);
    p -= 2;
    wp += 8 * KG;
  }
  if (blk <= blk_hi) {
    double t4, t5, t6;
    block(p, wp, v0, v1, v2, t4, t5, t6);
  }
}

// the three-slice rate map of model-age-groups2q.R:240-251 for the six series.
// For a valid data_offset the rate applied at sim day t is rate[i], i = 1 for t <= t1,
// t - t1 + 1 for t1 < t < t2, and length(rate) for t >= t2 (the later slice assignments overwrite
// the shared end points), where t2 = min(padding + period, t1 + length(rate) - 1); the whole map
// is skipped (series stays 0) unless t1 > padding && t2 > t1.  The index is computed branch-free so
// that the eleven base-vector loads of a day are issued back to back (profiles/r1_v7: the serial
// load -> branch -> load chains were 24 % long-scoreboard stalls).
template <int FL>
struct AgeMap {
  static constexpr int kIfrred = Traits<FL>::k2x ? 45 : 42;  // ifrred: parameter 43 of 2q, 46 of 2x
  int t1[6], t2[6];
  // age-2x: the case series are multiplied by symp.cases.factor on the days between t.symp and min(t.all, T)
  // (model-age-groups2x.R:278-279,293-294,340-341; R's a:b runs downwards when a > b); valid offsets only
  int symp_lo, symp_hi;
  __device__ __forceinline__ double case_factor(const DevModel &M, int t) const {
    if constexpr (Traits<FL>::k2x) return (valid && t >= symp_lo && t <= symp_hi) ? M.symp_factor : 1.0;
    return 1.0;
  }
  bool on[6];  // t1 > pad && t2 > t1, or invalid offset (then rate[1] everywhere, :250)
  bool valid, all_on;
  __device__ __forceinline__ void init(const DevModel &M, const double *th, int off_raw, int pad, int P) {
    valid = off_raw != EPI_INVALID_DATA_OFFSET;
    // y.DL is used for both groups (2q :290,306; 2i :264,280)
    const double ydl = Traits<FL>::k2i ? th[13] : th[12];
    // (y.CL, o.CL: parameters 44, 45 of 2q, 47, 48 of 2x)
    const double lat[6] = {Traits<FL>::k2i ? th[10] : Traits<FL>::k2x ? th[46] : th[43], Traits<FL>::k2i ? th[12] : th[11], 0,
                           Traits<FL>::k2i ? th[15] : Traits<FL>::k2x ? th[47] : th[44], Traits<FL>::k2i ? th[17] : th[15], 0};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const int shift = (q == 2 || q == 5) ? 0 : (int)rint(-ydl + lat[q]);
      t1[q] = off_raw + shift;
      t2[q] = min(pad + P, t1[q] + M.n_rate - 1);
      on[q] = !valid || (t1[q] > pad && t2[q] > t1[q]);
    }
    all_on = on[0] && on[1] && on[2] && on[3] && on[4] && on[5];
    symp_lo = 1; symp_hi = 0;
    if constexpr (Traits<FL>::k2x) {
      const int a = off_raw + M.d_symp, b = min(off_raw + M.d_all - 1, pad + P);
      symp_lo = min(a, b); symp_hi = max(a, b);
    }
  }
  // age-2x: a window that would start beyond the vector grows it in R (and moves calclogl's windows): unsupported
  __device__ __forceinline__ bool symp_unsupported(const DevModel &M, int off_raw, int pad, int P) const {
    if constexpr (Traits<FL>::k2x) {
      if (!valid) return false;
      const int a = off_raw + M.d_symp, b = min(off_raw + M.d_all - 1, pad + P);
      return a > pad + P || a < 1 || b < 1;
    }
    return false;
  }
  // 0-based index into the rate vectors for profile q at sim day t
  __device__ __forceinline__ int index(const DevModel &M, int q, int t) const {
    const int i = (t >= t2[q]) ? M.n_rate : max(t - t1[q] + 1, 1);
    return (valid ? i : 1) - 1;
  }
  // val[q] = inc[q] * rate_q(t) for the six profiles (ycase, yhosp, ydied, ocase, ohosp, odied)
  __device__ __forceinline__ void values(const DevModel &M, const double *th, int t, const double *inc, double *val) const {
    int k[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) k[q] = index(M, q, t);
    if constexpr (Traits<FL>::k2i) {
      // model-age-groups2i.R:212-304: the rates are the data vectors themselves; products are
      // left-associative as written there (s2i * yhosp_rate * y.hr[.]), and the InvalidDataOffset
      // branch of the case series has no pmin (:224,272)
      const double yifr0 = __ldg(M.y_ifr + k[0]), yhr1 = __ldg(M.y_hr + k[1]), yifr2 = __ldg(M.y_ifr + k[2]);
      const double oifr3 = __ldg(M.o_ifr + k[3]), ohr4 = __ldg(M.o_hr + k[4]), oifr5 = __ldg(M.o_ifr + k[5]);
      double r[6];
      if (valid) {
        double c0 = th[9] * yifr0, c3 = th[14] * oifr3;
        c0 = c0 > 1 ? 1.0 : c0; c3 = c3 > 1 ? 1.0 : c3;
        r[0] = inc[0] * c0; r[3] = inc[3] * c3;
      } else {
        r[0] = inc[0] * th[9] * M.ifr_y1; r[3] = inc[3] * th[14] * M.ifr_o1;
      }
      r[1] = inc[1] * th[11] * yhr1;
      r[2] = inc[2] * yifr2;
      r[4] = inc[4] * th[16] * ohr4;
      r[5] = inc[5] * oifr5;
#pragma unroll
      for (int q = 0; q < 6; ++q) val[q] = on[q] ? r[q] : 0.0;
      return;
    }
    // all loads first (model-age-groups2q.R:229-235)
    const double yifr0 = __ldg(M.y_ifr + k[0]), g0 = __ldg(M.g + k[0]);
    const double yhr1 = __ldg(M.y_hr + k[1]), g1 = __ldg(M.g + k[1]);
    const double yifr2 = __ldg(M.y_ifr + k[2]), g2 = __ldg(M.g + k[2]), gt2 = __ldg(M.gtrimp + k[2]);
    const double oifr3 = __ldg(M.o_ifr + k[3]);
    const double ohr4 = __ldg(M.o_hr + k[4]);
    const double oifr5 = __ldg(M.o_ifr + k[5]), gt5 = __ldg(M.gtrimp + k[5]);
    const double ifrred = th[kIfrred];
    double r[6];
    r[0] = th[9] * yifr0 * (1 + (th[21] - 1) * g0);   r[0] = r[0] > 1 ? 1.0 : r[0];  // gycr = pmin(1, .)
    r[1] = th[10] * yhr1 * (1 + (th[22] - 1) * g1);                                   // gyhr
    r[2] = yifr2 * (1 + (th[21] - 1) * g2) * (1 - ifrred * gt2);                      // gyifr
    r[3] = th[13] * oifr3;                            r[3] = r[3] > 1 ? 1.0 : r[3];  // gocr
    r[4] = th[14] * ohr4;                                                             // gohr
    r[5] = oifr5 * (1 - ifrred * gt5);                                                // goifr
    if (all_on) {  // per-chain, i.e. warp-uniform; the usual case
#pragma unroll
      for (int q = 0; q < 6; ++q) val[q] = inc[q] * r[q];
    } else {
#pragma unroll
      for (int q = 0; q < 6; ++q) val[q] = on[q] ? inc[q] * r[q] : 0.0;
    }
  }
  // the same for the three profiles of ONE group (G = 0: ycase, yhosp, ydied; G = 1: ocase, ohosp, odied):
  // val[c] = inc[c] * rate_{3 G + c}(t), the same operations in the same order as values()
  template <int G>
  __device__ __forceinline__ void values3(const DevModel &M, const double *th, int t, const double *inc, double *val) const {
    int k[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) k[c] = index(M, 3 * G + c, t);
    double r[3];
    if constexpr (Traits<FL>::k2i) {
      if constexpr (G == 0) {
        const double yifr0 = __ldg(M.y_ifr + k[0]), yhr1 = __ldg(M.y_hr + k[1]), yifr2 = __ldg(M.y_ifr + k[2]);
        if (valid) { double c0 = th[9] * yifr0; c0 = c0 > 1 ? 1.0 : c0; r[0] = inc[0] * c0; }
        else r[0] = inc[0] * th[9] * M.ifr_y1;
        r[1] = inc[1] * th[11] * yhr1;
        r[2] = inc[2] * yifr2;
      } else {
        const double oifr3 = __ldg(M.o_ifr + k[0]), ohr4 = __ldg(M.o_hr + k[1]), oifr5 = __ldg(M.o_ifr + k[2]);
        if (valid) { double c3 = th[14] * oifr3; c3 = c3 > 1 ? 1.0 : c3; r[0] = inc[0] * c3; }
        else r[0] = inc[0] * th[14] * M.ifr_o1;
        r[1] = inc[1] * th[16] * ohr4;
        r[2] = inc[2] * oifr5;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) val[c] = on[3 * G + c] ? r[c] : 0.0;
      return;
    }
    if constexpr (G == 0) {
      const double yifr0 = __ldg(M.y_ifr + k[0]), g0 = __ldg(M.g + k[0]);
      const double yhr1 = __ldg(M.y_hr + k[1]), g1 = __ldg(M.g + k[1]);
      const double yifr2 = __ldg(M.y_ifr + k[2]), g2 = __ldg(M.g + k[2]), gt2 = __ldg(M.gtrimp + k[2]);
      r[0] = th[9] * yifr0 * (1 + (th[21] - 1) * g0);   r[0] = r[0] > 1 ? 1.0 : r[0];  // gycr = pmin(1, .)
      r[1] = th[10] * yhr1 * (1 + (th[22] - 1) * g1);                                   // gyhr
      r[2] = yifr2 * (1 + (th[21] - 1) * g2) * (1 - th[kIfrred] * gt2);                 // gyifr
    } else {
      const double oifr3 = __ldg(M.o_ifr + k[0]);
      const double ohr4 = __ldg(M.o_hr + k[1]);
      const double oifr5 = __ldg(M.o_ifr + k[2]), gt5 = __ldg(M.gtrimp + k[2]);
      r[0] = th[13] * oifr3;                            r[0] = r[0] > 1 ? 1.0 : r[0];  // gocr
      r[1] = th[14] * ohr4;                                                             // gohr
      r[2] = oifr5 * (1 - th[kIfrred] * gt5);                                           // goifr
    }
    if (all_on) {
#pragma unroll
      for (int c = 0; c < 3; ++c) val[c] = inc[c] * r[c];
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) val[c] = on[3 * G + c] ? inc[c] * r[c] : 0.0;
    }
  }
};

// one scored series at one model day: the slot fetch (issued before the day's model value is computed) and its score
struct SeriesSlot {
  Slot first;
  int r;
  bool ok;
  __device__ __forceinline__ void load(const DevModel &M, int s, int t, int dstart_s) {
    r = t - dstart_s;
    ok = r >= 0 && r < M.n_obs[s];
    first = ok ? ldg_slot(M.slots[s] + (size_t)r * M.K[s]) : Slot{0.0, 0.0f, 0.0f};
  }
  __device__ __forceinline__ double score(const DevModel &M, const double2 *__restrict__ ltab, int s, double v) const {
    return score_slots(M, ltab, s, r, ok, first, v);
  }
};

// the table of tlog(): computed once per device by k_fill_log_table (first launch), copied into shared memory
// by every k_score block (round 1 recomputed its 128 divisions and logarithms in every block: 1 % of the kernel)
__device__ double2 g_log_table[kLogTabN];
__global__ void k_fill_log_table() { fill_log_table(g_log_table, threadIdx.x, blockDim.x); }

template <int FL>
__global__ void __launch_bounds__(128, 4)
k_score(const DevModel M, const PriorTerm *__restrict__ PT, int n_prior, const double *__restrict__ theta, int64_t ld, int64_t n,
        const double *__restrict__ hist2, const int *__restrict__ offset, const int *__restrict__ padv,
        const double *__restrict__ Wg, const int2 *__restrict__ pmeta, int *__restrict__ status,
        double *__restrict__ logl, double *__restrict__ logp, double *__restrict__ terms, int window_only) {
  constexpr int NK = Traits<FL>::NK, NG = Traits<FL>::NG, KG = NK / NG;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  double2 *ltab = reinterpret_cast<double2 *>(smem);  // block-shared log table, then the per-warp regions
  for (int i = threadIdx.x; i < kLogTabN; i += blockDim.x) ltab[i] = g_log_table[i];  // (filled once per device)
  __syncthreads();
  if (chain >= n) return;
  const int P = M.P;
  const ScoreSmem w = carve_score<FL>(smem + 2 * kLogTabN + (size_t)wib * score_smem_doubles<FL>(P), P);
  int st = status[chain];

  if (st & EPI_ST_UNSUPPORTED) {
    for (int p = lane; p < M.d; p += 32) w.theta[p] = theta[(int64_t)p * ld + chain];
    __syncwarp();
    double lp = 0.0;
    for (int k = lane; k < n_prior; k += 32) lp += prior_term(PT[k], w.theta);
    lp = warp_sum(lp);
    if (lane == 0) {
      if (logl) logl[chain] = NAN;
      if (logp) logp[chain] = lp;
      if (terms) for (int k = 0; k < 8; ++k) terms[chain * 8 + k] = NAN;
    }
    return;
  }

  int dmin_[NK], n_[NK], lo[NG], hi[NG];
  stage_score<FL>(M, w, theta, ld, chain, true, hist2, Wg, pmeta, P, lane, dmin_, n_, lo, hi);

  // prior on the UNTRANSFORMED vector (R/MCMC/fitMCMC-drj.R:51); no prior term reads
  // theta[4] in the simple flavour, so the staged model-space copy serves.  (While the S history and the profile
  // table are still in flight.)
  double lp = 0.0;
  for (int k = lane; k < n_prior; k += 32) lp += prior_term(PT[k], w.theta);
  lp = warp_sum(lp);
  stage_score_wait(w);

  const int pad = padv[chain];
  const int T = pad + P;
  const int off_raw = offset[chain];
  const int off = off_raw != EPI_INVALID_DATA_OFFSET ? off_raw : 1;  // :493-494
  constexpr int PAD = Traits<FL>::kPad;
  const int AS = PAD / 4 + (hist_pitch2(P) >> 2);  // phase-row stride of the staged S

  // calclogl reads the model series only inside the data windows [dstart_s, dstart_s + n_s - 1] (:497-590):
  // the sweep covers the union of the windows plus ONE day in front of it (the predecessor the first scored
  // day's increment needs), not the whole period.  Days before padding+1 read the 0 prefill.
  int dstart[kMaxSeries];
  int tlo = T, thi = 1;
  constexpr int NSER = Traits<FL>::kAge ? 5 : 2;  // == M.n_series (static trip counts keep dstart[] in registers)
#pragma unroll
  for (int s = 0; s < NSER; ++s) {
    window(off, M.n_obs[s], T, dstart[s]);
    tlo = min(tlo, dstart[s]);
    thi = max(thi, dstart[s] + M.n_obs[s] - 1);
  }

  double acc[kMaxSeries] = {0, 0, 0, 0, 0};
  double ysum = 0.0, osum = 0.0;
  // (window_only == 0, EPI_WINDOW_ONLY=0: sweep the whole period as the reference does — test switch)
  const int jlo = window_only ? tlo - 1 - (pad + 1) : min(tlo, pad + 1) - (pad + 1);
  const int jhi = window_only ? min(thi - (pad + 1), P - 1) : P - 1;

  // days in front of padding+1 (j < 0) hold the 0 prefill of every model series (:118-136): scored with the value 0
  for (int j = jlo + lane; j < 0; j += 32) {
    const int t = pad + 1 + j;
    if constexpr (Traits<FL>::kAge) {
      DaySlots<5> slots;
      slots.load(M, t, dstart, true);
      const double v5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      slots.score(M, ltab, v5, acc);
    } else {
      DaySlots<2> slots;
      slots.load(M, t, dstart, true);
      const double v2[2] = {0.0, 0.0};
      slots.score(M, ltab, v2, acc);
    }
  }

  // Sweeps of 128 days: lane L owns the four consecutive days 4 G .. 4 G + 3, G = base / 4 + L (conv_tile4).  Lanes
  // whose days lie behind the period convolve the last in-range quad again (their results are discarded).
  const int Gmax = (P - 1) >> 2;
  const int blo[2] = {lo[0] >> 2, lo[NG - 1] >> 2}, bhi[2] = {hi[0] >> 2, hi[NG - 1] >> 2};
  if constexpr (Traits<FL>::kAge) {
    AgeMap<FL> map;
    map.init(M, w.theta, off_raw, pad, P);
    if (map.symp_unsupported(M, off_raw, pad, P)) {
      if (lane == 0) {
        if (logl) logl[chain] = NAN;
        if (logp) logp[chain] = lp;
        status[chain] = st | EPI_ST_UNSUPPORTED;
        if (terms) for (int k = 0; k < 8; ++k) terms[chain * 8 + k] = NAN;
      }
      return;
    }
    double carry[6] = {0, 0, 0, 0, 0, 0};  // s2 of the day in front of the sweep (0 in front of day 0)
    // One S group at a time, so that only twelve convolution results are live while a group's days are scored
    // (with all 24 the day loop spilled: profiles/r2_summary.md): the y group scores y.casei and y.deadi and
    // parks its hospital series, the o group scores o.casei, o.deadi and hospi = y.hospi + o.hospi (:552) and the
    // sums of the ratio term (:563-565).  Per series the days still arrive in the same order as before.
    for (int base = max(jlo, 0) & ~3; base <= jhi; base += 128) {
      const int G = (base >> 2) + lane, Gc = min(G, Gmax);
      const int j0 = 4 * G, t0 = pad + 1 + j0;
      const bool mine = j0 + 3 >= jlo && j0 <= jhi;  // some day of this lane lies inside the swept range
      double yhosp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        double c[12];  // c[r * 3 + q]
        conv_tile4<3, PAD>(w.S + g * 4 * AS, AS, w.Wabs + g * PAD * 3, blo[g], bhi[g], Gc, c);
        // s2 = N - conv (NA while stats::filter has no full window: reachable only for the hospital profiles, the
        // other four define the padding (:113-114), so t - dmin >= n for every t > padding); 0 behind the period
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            double &x = c[r * 3 + q];
            if (q == 1 && t0 + r - dmin_[g * 3 + 1] < n_[g * 3 + 1]) x = NAN;
            // (age-2x: the convolution of the incidence IS the series, model-age-groups2x.R:282,301,315)
            if constexpr (Traits<FL>::k2x) x = (j0 + r < P) ? x : 0.0;
            else x = (j0 + r < P) ? (g ? M.No : M.Ny) - x : 0.0;
          }
        // c(s2[1], diff(s2)), :239: the predecessor of the lane's first day is the last day of the lane below (the
        // carry of the previous sweep for lane 0; the 0.0 in front of day 0, and s2 - 0.0 == s2 exactly)
        double prev[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          prev[q] = __shfl_up_sync(0xffffffffu, c[9 + q], 1);
          if (lane == 0) prev[q] = carry[g * 3 + q];
          carry[g * 3 + q] = __shfl_sync(0xffffffffu, c[9 + q], 31);
        }
        if (mine) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int j = j0 + r, t = t0 + r;
            if (j >= P) break;
            // scored series (profile order within a group: case, hosp, died):
            //   0 y.casei, 1 o.casei, 2 hospi = y.hospi + o.hospi, 3 y.deadi, 4 o.deadi
            SeriesSlot sc, sd, sh;
            sc.load(M, g, t, dstart[g]);
            sd.load(M, 3 + g, t, dstart[3 + g]);
            if (g == 1) sh.load(M, 2, t, dstart[2]);
            double inc[3], val[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              inc[q] = Traits<FL>::k2x ? c[r * 3 + q] : c[r * 3 + q] - prev[q];
              prev[q] = c[r * 3 + q];
            }
            if (g == 0) map.template values3<0>(M, w.theta, t, inc, val);
            else map.template values3<1>(M, w.theta, t, inc, val);
            if constexpr (Traits<FL>::k2x) val[0] = val[0] * map.case_factor(M, t);
            acc[g] += sc.score(M, ltab, g, val[0]);
            acc[3 + g] += sd.score(M, ltab, 3 + g, val[2]);
            if (g == 0) {
              yhosp[r] = val[1];
            } else {
              acc[2] += sh.score(M, ltab, 2, yhosp[r] + val[1]);
              if constexpr (!Traits<FL>::k2i)
                if (t >= dstart[2] + M.d_hosp_o1 && t <= dstart[2] + M.d_hosp_o2) { ysum += yhosp[r]; osum += val[1]; }  // :563-564
            }
          }
        }
      }
    }
  } else {
    // simple: hosp = (N - conv)*HR, died = (N - conv)*0.007 on the padded vector (0 before padding+1),
    // hospi/deadi = c(v[1], diff(v)) (model-simple-ode-drj-cmp.R:115-122)
    const double rate[2] = {w.theta[4], 0.007};
    double carry[2] = {0, 0};
    for (int base = max(jlo, 0) & ~3; base <= jhi; base += 128) {
      const int G = (base >> 2) + lane, Gc = min(G, Gmax);
      double c[8];  // c[r * 2 + q]
      conv_tile4<2, PAD>(w.S, AS, w.Wabs, blo[0], bhi[0], Gc, c);
      const int j0 = 4 * G, t0 = pad + 1 + j0;
#pragma unroll
      for (int i = 0; i < 8; ++i)  // both profiles define the padding (:59): never NA for t > padding
        c[i] = (j0 + (i >> 1) < P) ? (M.N - c[i]) * rate[i & 1] : 0.0;
      double prev[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        prev[q] = __shfl_up_sync(0xffffffffu, c[6 + q], 1);
        if (lane == 0) prev[q] = carry[q];  // 0 in front of day 0 (prefill)
        carry[q] = __shfl_sync(0xffffffffu, c[6 + q], 31);
      }
      if (j0 + 3 >= jlo && j0 <= jhi) {
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
          const int j = j0 + r, t = t0 + r;
          if (j >= P) break;
          DaySlots<2> slots;
          slots.load(M, t, dstart, true);
          const double v2[2] = {c[0] - prev[0], c[1] - prev[1]};
          prev[0] = c[0]; prev[1] = c[1];
#pragma unroll
          for (int i = 0; i < 6; ++i) c[i] = c[i + 2];
          slots.score(M, ltab, v2, acc);
        }
      }
    }
  }
  (void)KG;

  double tv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // (static indices only: a dynamically indexed acc[] / tv[] lives in local memory for the whole kernel)
#pragma unroll
  for (int s = 0; s < NSER; ++s) {
    const int term = M.term_of[s];
    const double v = warp_sum(acc[s]) + M.term_const[term];
#pragma unroll
    for (int k = 0; k < 8; ++k) tv[k] = (k == term) ? v : tv[k];
  }
  if constexpr (Traits<FL>::k2i) {
    tv[3] = 0.0;  // no hospital-ratio term in this generation (model-age-groups2i.R:526)
  } else if constexpr (Traits<FL>::kAge) {
    ysum = warp_sum(ysum);
    osum = warp_sum(osum);
    tv[3] = dev_dlnorm_log(ysum / (ysum + osum), log(56.7 / 100), log(1.05));  // :565
  } else {
    // loglLD, model-simple-ode-drj-cmp.R:203-204
    const double mu = dev_pmax(0.1, w.theta[7]);
    tv[0] = M.tdl * log(mu) - (M.tdl + M.mort_size) * log(M.mort_size + mu) + M.term_const[0];
  }
  double ll;
  if constexpr (Traits<FL>::k2i) ll = tv[0] + tv[1] + tv[2] + tv[4] + tv[5];  // 2i :526
  else if constexpr (Traits<FL>::kAge) ll = tv[0] + tv[1] + tv[2] + tv[3] + tv[4] + tv[5];  // :592
  else ll = tv[0] + tv[1] + tv[2];
  if (ll != ll) st |= EPI_ST_NA;
  if (lane == 0) {
    if (logl) logl[chain] = ll;
    if (logp) logp[chain] = lp;
    status[chain] = st;
    if (terms) for (int k = 0; k < 8; ++k) terms[chain * 8 + k] = tv[k];
  }
}

// ------------------------------------------------------------- trajectories
// calculateModel() output series for days padding+1..padding+P (epi_trajectories):
// age: y.S,o.S,y.casei,o.casei,y.hospi,o.hospi,y.deadi,o.deadi; simple: S,hospi,deadi.
template <int FL>
__global__ void __launch_bounds__(128, 4)
k_series(const DevModel M, const double *__restrict__ theta, int64_t ld, int64_t n, const double *__restrict__ hist2,
         const int *__restrict__ offset, const int *__restrict__ padv, const double *__restrict__ Wg,
         const int2 *__restrict__ pmeta, const int *__restrict__ status, double *__restrict__ out, int ns_total) {
  constexpr int NK = Traits<FL>::NK, NG = Traits<FL>::NG;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (chain >= n) return;
  const int P = M.P;
  const ScoreSmem w = carve_score<FL>(smem + (size_t)wib * score_smem_doubles<FL>(P), P);
  double *o = out + (int64_t)chain * ns_total * P;  // basic series first, ext series (if any) behind them
  if (status[chain] & EPI_ST_UNSUPPORTED) {
    for (int i = lane; i < ns_total * P; i += 32) o[i] = NAN;
    return;
  }
  // theta is already in MODEL space here (calculateModel is called with transformed params)
  int dmin_[NK], n_[NK], lo[NG], hi[NG];
  stage_score<FL>(M, w, theta, ld, chain, false, hist2, Wg, pmeta, P, lane, dmin_, n_, lo, hi);
  stage_score_wait(w);
  const int pad = padv[chain];
  const int off_raw = offset[chain];
  constexpr int PAD = Traits<FL>::kPad;
  const int AS = PAD / 4 + (hist_pitch2(P) >> 2);
  const int Gmax = (P - 1) >> 2;
  const int blo[2] = {lo[0] >> 2, lo[NG - 1] >> 2}, bhi[2] = {hi[0] >> 2, hi[NG - 1] >> 2};
  // the same sweep as k_score (four consecutive days per lane, conv_tile4) over the whole period
  if constexpr (Traits<FL>::kAge) {
    AgeMap<FL> map;
    map.init(M, w.theta, off_raw, pad, P);
    const int dst[6] = {2, 4, 6, 3, 5, 7};
    double carry[6] = {0, 0, 0, 0, 0, 0};
    for (int base = 0; base < P; base += 128) {
      const int G = (base >> 2) + lane, Gc = min(G, Gmax);
      double c[2][12];
      conv_tile4<3, PAD>(w.S, AS, w.Wabs, blo[0], bhi[0], Gc, c[0]);
      conv_tile4<3, PAD>(w.S + 4 * AS, AS, w.Wabs + PAD * 3, blo[1], bhi[1], Gc, c[1]);
      const int j0 = 4 * G, t0 = pad + 1 + j0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          double &x = c[q / 3][r * 3 + q % 3];
          if (t0 + r - dmin_[q] < n_[q]) x = NAN;
          if constexpr (Traits<FL>::k2x) x = (j0 + r < P) ? x : 0.0;
          else x = (j0 + r < P) ? (q < 3 ? M.Ny : M.No) - x : 0.0;
        }
      double prev[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const double last = c[q / 3][9 + q % 3];
        prev[q] = __shfl_up_sync(0xffffffffu, last, 1);
        if (lane == 0) prev[q] = carry[q];
        carry[q] = __shfl_sync(0xffffffffu, last, 31);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = j0 + r, t = t0 + r;
        if (j >= P) break;
        if constexpr (!Traits<FL>::k2x) {  // (age-2x: the staged rows hold the incidence; S is written by the ext ODE instance)
          o[0 * P + j] = staged_s<PAD>(w.S, 0, AS, j);
          o[1 * P + j] = staged_s<PAD>(w.S, 1, AS, j);
        }
        double inc[6], val[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double s2 = c[q / 3][r * 3 + q % 3];
          inc[q] = (Traits<FL>::k2x || j == 0) ? s2 : s2 - prev[q];
          prev[q] = s2;
        }
        map.values(M, w.theta, t, inc, val);
        if constexpr (Traits<FL>::k2x) {
          const double f = map.case_factor(M, t);
          val[0] = val[0] * f; val[3] = val[3] * f;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) o[dst[q] * P + j] = val[q];
      }
    }
  } else {
    const double rate[2] = {w.theta[4], 0.007};
    double carry[2] = {0, 0};
    for (int base = 0; base < P; base += 128) {
      const int G = (base >> 2) + lane, Gc = min(G, Gmax);
      double c[8];
      conv_tile4<2, PAD>(w.S, AS, w.Wabs, blo[0], bhi[0], Gc, c);
      const int j0 = 4 * G, t0 = pad + 1 + j0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (t0 + (i >> 1) - dmin_[i & 1] < n_[i & 1]) c[i] = NAN;
        c[i] = (j0 + (i >> 1) < P) ? (M.N - c[i]) * rate[i & 1] : 0.0;
      }
      double prev[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        prev[q] = __shfl_up_sync(0xffffffffu, c[6 + q], 1);
        if (lane == 0) prev[q] = carry[q];
        carry[q] = __shfl_sync(0xffffffffu, c[6 + q], 31);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = j0 + r;
        if (j >= P) break;
        o[j] = staged_s<PAD>(w.S, 0, AS, j);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          o[(1 + q) * P + j] = c[r * 2 + q] - prev[q];
          prev[q] = c[r * 2 + q];
        }
      }
    }
  }
}

// ------------------------------------------------------------------ quantiles
// quantileData (R/lib/libMCMC.R:151-172): one block per output day; the n draws' values of that
// day are gathered (takeAndPad alignment by each draw's data_offset), bitonic-sorted in shared
// memory with NA last, and the R type-7 quantiles are interpolated.
__global__ void __launch_bounds__(1024)
k_quantiles(const double *__restrict__ series, const int *__restrict__ offset, const int *__restrict__ padv, int64_t n,
            int NS, int simperiod, int startoffset, int sa, int sb, double prefill_a, double prefill_b,
            const double *__restrict__ probs, int n_probs, int npow2, double *__restrict__ out,
            double *__restrict__ gscratch) {
  // the day's values: in shared memory up to 8,192 draws, else in this block's slice of a global scratch buffer
  // (the same block-wide bitonic network, through L2)
  extern __shared__ double vsh[];
  double *v = gscratch ? gscratch + (size_t)blockIdx.x * npow2 : vsh;
  __shared__ int n_valid;
  const int j = blockIdx.x + 1;  // day, 1-based
  if (threadIdx.x == 0) n_valid = 0;
  __syncthreads();
  int mine = 0;
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    double x = INFINITY;  // NA sorts last
    if (i < n) {
      const int off = offset[i], pad = padv[i];
      const long long sidx = (long long)off - startoffset - 1 + j;  // sim index read by takeAndPad
      if (off != EPI_INVALID_DATA_OFFSET && sidx >= 1 && sidx <= simperiod && off - startoffset >= 1) {
        const double *base = series + (int64_t)i * NS * simperiod;
        const int k = (int)sidx - pad - 1;  // index into the stored days padding+1..padding+simperiod
        double a = k >= 0 ? base[(int64_t)sa * simperiod + k] : prefill_a;
        if (sb >= 0) a += k >= 0 ? base[(int64_t)sb * simperiod + k] : prefill_b;
        if (a == a) { x = a; ++mine; }
      }
    }
    v[i] = x;
  }
  atomicAdd(&n_valid, mine);
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1)
    for (int s = k >> 1; s > 0; s >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        const int p = i ^ s;
        if (p > i) {
          const bool up = (i & k) == 0;
          const double a = v[i], b = v[p];
          if ((a > b) == up) { v[i] = b; v[p] = a; }
        }
      }
      __syncthreads();
    }
  const int m = n_valid;
  for (int q = threadIdx.x; q < n_probs; q += blockDim.x) {
    double r = NAN;
    if (m > 0) {  // R quantile type 7
      const double hq = (double)(m - 1) * probs[q];
      const int lo = (int)floor(hq);
      const int hi = min(lo + 1, m - 1);
      r = v[lo] + (hq - (double)lo) * (v[hi] - v[lo]);
    }
    out[(int64_t)blockIdx.x * n_probs + q] = r;
  }
}

// AoS (n x d, one proposal per row) -> SoA (d x ld) transpose, optionally applying
// transformParams (used by the host-buffer entry points).
// marginalizeData (R/lib/libMCMC.R:174-188): per posterior draw v = takeAndPad(fun(state), offset - startoffset,
// simperiod) and ONE number marginfun(v).  marginfun here: max / sum / min / mean of v[j1..j2] (the reference's only
// call takes the maximum of Re over a 60-day window, R/MCMC/predictMCMC.R:714-717); like R's max() and sum() an NA in
// the range makes the result NA.  One warp per draw.
__global__ void __launch_bounds__(128)
k_marginal(const double *__restrict__ series, const int *__restrict__ offset, const int *__restrict__ padv, int64_t n,
           int NS, int simperiod, int startoffset, int sa, int sb, double prefill_a, double prefill_b, int kind, int j1,
           int j2, double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int off = offset[i], pad = padv[i];
  const double *base = series + i * NS * (int64_t)simperiod;
  double acc = kind == 0 ? -INFINITY : kind == 2 ? INFINITY : 0.0;
  bool na = false;
  for (int j = j1 + lane; j <= j2; j += 32) {
    double a = NAN;
    const long long sidx = (long long)off - startoffset - 1 + j;  // sim index read by takeAndPad
    if (off != EPI_INVALID_DATA_OFFSET && sidx >= 1 && sidx <= simperiod && off - startoffset >= 1) {
      const int k = (int)sidx - pad - 1;
      a = k >= 0 ? base[(int64_t)sa * simperiod + k] : prefill_a;
      if (sb >= 0) a += k >= 0 ? base[(int64_t)sb * simperiod + k] : prefill_b;
    }
    if (a != a) na = true;
    else acc = kind == 0 ? fmax(acc, a) : kind == 2 ? fmin(acc, a) : acc + a;
  }
  na = __any_sync(0xffffffffu, na);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double b = __shfl_xor_sync(0xffffffffu, acc, o);
    acc = kind == 0 ? fmax(acc, b) : kind == 2 ? fmin(acc, b) : acc + b;
  }
  if (kind == 3) acc /= (double)(j2 - j1 + 1);
  if (lane == 0) out[i] = na ? NAN : acc;
}

__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ soa, int64_t n, int d, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  const int64_t c = i / d;
  const int p = (int)(i % d);
  soa[(int64_t)p * ld + c] = aos[i];
}

// DFMA-chain microbenchmark: 8 independent FMA chains per thread (epi_fp64_peak)
__global__ void k_dfma(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace epi

// ------------------------------------------------------------ launch wrappers

namespace epi {

template <int FL>
static cudaError_t launch_eval_fl(const EvalArgs &a) {
  const DevModel &M = *a.M;
  const int64_t n = a.n;
  const unsigned ode_blocks = (unsigned)((n + 31) / 32);
  const unsigned warp_blocks = (unsigned)((n + 3) / 4);
  const size_t smem_mid = (size_t)4 * warp_smem_doubles<FL>(M.P1) * sizeof(double);
  const size_t smem_score = ((size_t)4 * score_smem_doubles<FL>(M.P) + 2 * kLogTabN) * sizeof(double);
  if (smem_mid > 200 * 1024 || smem_score > 200 * 1024) return cudaErrorInvalidConfiguration;  // period too long (checked at create)
  // function attributes and the SM count are per DEVICE: one process may hold handles on several GPUs
  constexpr int kMaxDev = 64;
  static bool attr_done[kMaxDev] = {};
  static int n_sm_of[kMaxDev] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int di = dev < kMaxDev ? dev : kMaxDev - 1;
  if (!attr_done[di] || dev >= kMaxDev) {
    cudaFuncSetAttribute(k_mid<FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_score<FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_series<FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if constexpr (Traits<FL>::kAge)
    {
      // four 128-thread blocks per SM (the register budget of the default instance) must also fit the SM's 227 KB of
      // shared memory, 1 KB of it reserved per block
      static_assert(4 * (seg_table_bytes<FL>(128) + 1024) <= 227 * 1024, "schedule table too large for four blocks per SM");
      cudaFuncSetAttribute(k_ode_age2<FL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_table_bytes<FL>(128));
      cudaFuncSetAttribute(k_ode_age2<FL, true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_table_bytes<FL>(128));
      // (without it the driver sizes the carve-out for ONE block per SM and a 16,384-chain shard runs in two waves)
      cudaFuncSetAttribute(k_ode_age2<FL, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(k_ode_age2<FL, true, 5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    cudaDeviceGetAttribute(&n_sm_of[di], cudaDevAttrMultiProcessorCount, dev);
    k_fill_log_table<<<1, kLogTabN, 0, a.stream>>>();
    cudaStreamSynchronize(a.stream);  // once per device and process: other streams may launch k_score next
    attr_done[di] = true;
  }
  const int n_sm = n_sm_of[di];
  cudaStream_t s = a.stream;
  if (a.ev) cudaEventRecord(a.ev[0], s);
  const unsigned pair_blocks = (unsigned)((2 * n + 127) / 128);
  // Lanes per chain of the age ODE passes: two (one per age group, warp-synchronous k_ode_age2) at every batch size
  // since the kernel votes once per day and reads the chain's schedule from a shared-memory table — step ms on the
  // posterior cloud, one / two lanes (profiles/r2_summary.md): 8,192 chains 0.82 / 0.52, 24,576 1.45 / 1.20,
  // 32,768 1.71 / 1.59, 65,536 3.09 / 2.94 (1,024 blocks at four per SM: 1.7 waves).  Thread per chain (k_ode) stays
  // the kernel of the other flavours, of the trajectories with the extended series and of the R-tracking fallback;
  // the four-lane kernel (k_ode_age4) is kept for EPI_ODE_LANES = 4.  EPI_ODE_LANES = 1 | 2 | 4 forces one mapping.
  int lanes = 1;
  if constexpr (Traits<FL>::kAge) {
    // (age-2i, whose schedule is mostly linear ramps — three more FP64 operations per stage that two lanes pay twice —
    //  at 65,536 chains: on the posterior cloud of a sampler 3.14 ms thread per chain against 2.93 with two lanes; on
    //  the narrow cloud, whose warps run in lock-step, 3.15 against 3.29.  The sampler is the use case.)
    lanes = a.ode_lanes ? a.ode_lanes : 2;
    // age-2x (five states per lane, R integrated; k_ode_age2<.., NS = 5>, bit-identical, tests): thread per chain wins
    // at every size measured — 65,536 chains 4.51 ms against 5.06, 8,192 chains 1.27 against 1.30 (posterior cloud) —
    // so two lanes there only on request
    if (Traits<FL>::k2x) lanes = a.ode_lanes == 2 || a.ode_lanes == 4 ? 2 : 1;
  }
  constexpr bool kFallback = !Traits<FL>::k2x;  // the R-tracking fallback instance exists where the hot one does not integrate R
  const bool use_pair = lanes == 2, use_quad = lanes == 4;
  const bool use_pair1 = use_pair;
  // EPI_ODE_MINB=5: the 96-register instance of the two-lane kernel, five blocks per SM
  const bool pair_dense = a.ode_minb == 5;  // (measured: 65,536 chains 2.926 ms against 2.937 at four, 8,192 chains 0.550 against 0.523)
  const unsigned quad_blocks = (unsigned)((4 * n + 127) / 128);
  const bool quad_roomy = (int64_t)quad_blocks <= (int64_t)n_sm * 4;  // all blocks resident at 128 registers
  // pass 1 that covers the whole period (simple / age-2i) is integrated for its first kPass1Chunk days first:
  // the threshold is normally crossed there and everything behind the crossing is never read.  Chains that
  // did not cross are redone over the whole pass-1 period by a retry round (k_mid flags them).
  const int chunk = (a.pass1_chunk && M.P1 > kPass1Chunk + kPass1Chunk / 4) ? kPass1Chunk : 0;
  int nl = 0;  // kernels enqueued
  // thread-per-chain pass 1: the hot instance (R not integrated) and, right behind it, the R-tracking instance for
  // the chains the hot one flagged (exits at once otherwise)
  auto pass1 = [&](int limit, const int *retry) {
    if (use_quad) {
      if constexpr (Traits<FL>::kAge) {
        if (quad_roomy) k_ode_age4<FL, false, 4><<<quad_blocks, 128, 0, s>>>(M, a.theta, a.ld, n, nullptr, nullptr, nullptr, a.hist1, a.ckpt, 0, limit, retry, a.need_r, a.force_need_r);
        else k_ode_age4<FL, false, 7><<<quad_blocks, 128, 0, s>>>(M, a.theta, a.ld, n, nullptr, nullptr, nullptr, a.hist1, a.ckpt, 0, limit, retry, a.need_r, a.force_need_r);
      }
    } else if (use_pair1) {
      if constexpr (Traits<FL>::kAge) {
        if (pair_dense) k_ode_age2<FL, false, 5><<<pair_blocks, 128, 0, s>>>(M, a.theta, a.ld, n, nullptr, nullptr, nullptr, a.hist1, a.ckpt, 0, limit, retry, a.need_r, a.force_need_r);
        else k_ode_age2<FL, false><<<pair_blocks, 128, 0, s>>>(M, a.theta, a.ld, n, nullptr, nullptr, nullptr, a.hist1, a.ckpt, 0, limit, retry, a.need_r, a.force_need_r);
      }
    } else {
      k_ode<FL, false><<<ode_blocks, 32, 0, s>>>(M, a.theta, a.ld, n, nullptr, nullptr, nullptr, a.hist1, a.ckpt, nullptr, 0, 0, limit, retry, a.need_r, a.force_need_r);
    }
    if constexpr (kFallback) k_ode<FL, false, false, true><<<ode_blocks, 32, 0, s>>>(M, a.theta, a.ld, n, nullptr, nullptr, nullptr, a.hist1, a.ckpt, nullptr, 0, 0, 0, nullptr, a.need_r, 0);
    nl += kFallback ? 2 : 1;
  };
  pass1(chunk, nullptr);
  if (a.ev) cudaEventRecord(a.ev[1], s);
  k_mid<FL><<<warp_blocks, 128, smem_mid, s>>>(M, a.theta, a.ld, n, a.hist1, a.offset, a.pad, a.status, a.W, a.pmeta, a.model_space, chunk, 0);
  nl += 1;
  if (chunk) {
    pass1(0, a.status);
    k_mid<FL><<<warp_blocks, 128, smem_mid, s>>>(M, a.theta, a.ld, n, a.hist1, a.offset, a.pad, a.status, a.W, a.pmeta, a.model_space, 0, 1);
    nl += 1;
  }
  if (a.ev) cudaEventRecord(a.ev[2], s);
  if (a.ev_after_mid) cudaEventRecord(a.ev_after_mid, s);
  const bool want_ext = a.series_out && (a.ns_total > (Traits<FL>::kAge ? 8 : 3));  // (age-2x trajectories always ask for it)
  const int window_only = (a.series_out || !a.window_only) ? 0 : 1;  // scoring reads S(t) only up to the last data-window day
  if ((use_pair || use_quad) && !want_ext) {
    if constexpr (Traits<FL>::kAge) {
      if (use_quad && quad_roomy) k_ode_age4<FL, true, 4><<<quad_blocks, 128, 0, s>>>(M, a.theta, a.ld, n, a.offset, a.pad, a.hist1, a.hist2, a.ckpt, window_only, 0, nullptr, a.need_r, a.force_need_r);
      else if (use_quad) k_ode_age4<FL, true, 7><<<quad_blocks, 128, 0, s>>>(M, a.theta, a.ld, n, a.offset, a.pad, a.hist1, a.hist2, a.ckpt, window_only, 0, nullptr, a.need_r, a.force_need_r);
      else if (pair_dense) k_ode_age2<FL, true, 5><<<pair_blocks, 128, seg_table_bytes<FL>(128), s>>>(M, a.theta, a.ld, n, a.offset, a.pad, a.hist1, a.hist2, a.ckpt, window_only, 0, nullptr, a.need_r, a.force_need_r);
      else k_ode_age2<FL, true><<<pair_blocks, 128, seg_table_bytes<FL>(128), s>>>(M, a.theta, a.ld, n, a.offset, a.pad, a.hist1, a.hist2, a.ckpt, window_only, 0, nullptr, a.need_r, a.force_need_r);
    }
    k_ode<FL, true, false, true><<<ode_blocks, 32, 0, s>>>(M, a.theta, a.ld, n, a.offset, a.pad, a.hist1, a.hist2, a.ckpt, nullptr, 0, window_only, 0, nullptr, a.need_r, 0);
    nl += 2;
  } else if (want_ext) {
    // trajectories with the ext series: pass 2 also writes E, I, R and the Re/Rt diagnostics (no restart)
    k_ode<FL, true, true><<<ode_blocks, 32, 0, s>>>(M, a.theta, a.ld, n, a.offset, a.pad, a.hist1, a.hist2, nullptr,
                                                    a.series_out, a.ns_total);
    nl += 1;
  } else {
    k_ode<FL, true><<<ode_blocks, 32, 0, s>>>(M, a.theta, a.ld, n, a.offset, a.pad, a.hist1, a.hist2, a.ckpt, nullptr, 0, window_only, 0, nullptr, a.need_r, a.force_need_r);
    if constexpr (kFallback) k_ode<FL, true, false, true><<<ode_blocks, 32, 0, s>>>(M, a.theta, a.ld, n, a.offset, a.pad, a.hist1, a.hist2, a.ckpt, nullptr, 0, window_only, 0, nullptr, a.need_r, 0);
    nl += kFallback ? 2 : 1;
  }
  if (a.ev) cudaEventRecord(a.ev[3], s);
  if (a.series_out) {
    k_series<FL><<<warp_blocks, 128, smem_score, s>>>(M, a.theta, a.ld, n, a.hist2, a.offset, a.pad, a.W, a.pmeta,
                                                      a.status, a.series_out, a.ns_total);
  } else {
    k_score<FL><<<warp_blocks, 128, smem_score, s>>>(M, a.PT, a.n_prior, a.theta, a.ld, n, a.hist2, a.offset, a.pad, a.W, a.pmeta,
                                                     a.status, a.logl, a.logp, a.terms, window_only);
  }
  nl += 1;
  if (a.n_launches) *a.n_launches += nl;
  if (a.ev) cudaEventRecord(a.ev[4], s);
  return cudaGetLastError();
}

cudaError_t launch_eval(const EvalArgs &a) {
  switch (a.M->flavour) {
    case EPI_FLAVOUR_SIMPLE: return launch_eval_fl<EPI_FLAVOUR_SIMPLE>(a);
    case EPI_FLAVOUR_NE: return launch_eval_fl<EPI_FLAVOUR_NE>(a);
    case EPI_FLAVOUR_AGE2I: return launch_eval_fl<EPI_FLAVOUR_AGE2I>(a);
    case EPI_FLAVOUR_AGE2X: return launch_eval_fl<EPI_FLAVOUR_AGE2X>(a);
    default: return launch_eval_fl<EPI_FLAVOUR_AGE2Q>(a);
  }
}

cudaError_t launch_aos_to_soa(const double *aos, double *soa, int64_t n, int d, int64_t ld, cudaStream_t s) {
  const int64_t tot = n * d;
  k_aos_to_soa<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(aos, soa, n, d, ld);
  return cudaGetLastError();
}

cudaError_t launch_quantiles(const double *series, const int *offset, const int *pad, int64_t n, int NS, int simperiod,
                             int period, int startoffset, int sa, int sb, double prefill_a, double prefill_b,
                             const double *d_probs, int n_probs, double *d_out, double *d_scratch, cudaStream_t s) {
  int npow2 = 1;
  while (npow2 < n) npow2 <<= 1;
  const size_t smem = d_scratch ? 0 : sizeof(double) * (size_t)npow2;
  cudaFuncSetAttribute(k_quantiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_quantiles<<<period, d_scratch ? 1024 : 256, smem, s>>>(series, offset, pad, n, NS, simperiod, startoffset, sa, sb,
                                                         prefill_a, prefill_b, d_probs, n_probs, npow2, d_out, d_scratch);
  return cudaGetLastError();
}

cudaError_t launch_marginal(const double *series, const int *offset, const int *pad, int64_t n, int NS, int simperiod,
                            int startoffset, int sa, int sb, double prefill_a, double prefill_b, int kind, int j1, int j2,
                            double *d_out, cudaStream_t s) {
  k_marginal<<<(unsigned)((n + 3) / 4), 128, 0, s>>>(series, offset, pad, n, NS, simperiod, startoffset, sa, sb, prefill_a,
                                                     prefill_b, kind, j1, j2, d_out);
  return cudaGetLastError();
}

cudaError_t launch_dfma(double *out, int blocks, int threads, int iters, cudaStream_t s) {
  k_dfma<<<blocks, threads, 0, s>>>(out, iters);
  return cudaGetLastError();
}

}  // namespace epi
