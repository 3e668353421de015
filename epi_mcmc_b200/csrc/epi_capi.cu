// epi_capi.cu — host side of libepimcmc.so: the C ABI of include/epimcmc.h.
//
// Handle creation snapshots the model file's R globals (data + settings) into one
// immutable device-side DevModel: the observed series become a day-major
// likelihood term table (R's argument recycling in dnbinom() resolved here,
// R/models/model-age-groups2q.R:514-554,584-590), the priors become a uniform
// term table (:433-485), df_params becomes the box.  There is no CPU path: every
// compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/epimcmc.h"
#include "epi_handle.h"
#include "epi_launch.h"

using namespace epi;

static thread_local std::string g_last_error;

int epi_fail(epi_handle *h, int rc, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  g_last_error = buf;
  return rc;
}

#define CUDA_OK(h, call)                                                                       \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return epi_fail(h, (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? EPI_ERR_NO_DEVICE : EPI_ERR_CUDA, \
                      "%s: %s", #call, cudaGetErrorString(e_));                                \
  } while (0)

static double age_gamma() { return 1.0 / ((4.7 - (3.0 + 1.0)) * 2.0); }  // model-age-groups2q.R:8-15
static bool is_age(int flavour) { return flavour == EPI_FLAVOUR_AGE2Q || flavour == EPI_FLAVOUR_AGE2I || flavour == EPI_FLAVOUR_AGE2X; }
static bool fixed_pass1(int flavour) { return flavour == EPI_FLAVOUR_AGE2Q || flavour == EPI_FLAVOUR_AGE2X; }  // period1 = 60
static int pad_rows(int flavour) { return flavour == EPI_FLAVOUR_AGE2I ? EPI_MAX_PADDING : kPadMax; }  // Traits<FL>::kPad

// ------------------------------------------------------------- metadata tables
static const char *kAgeNames[45] = {  // fit.paramnames, model-age-groups2q.R:604-618
    "betay0", "betao0", "betayo0", "betay1", "betao1", "betayo1", "betay2", "betao2", "betayo2",
    "y.CR", "y.HR", "y.HL", "y.DL", "o.CR", "o.HR", "o.HL", "o.DL", "t0_morts", "t0o", "HLsd", "DLsd",
    "fyifr", "fyhr", "t3o", "t4o", "betay4", "betao4", "betayo4", "betay5", "betao5", "betayo5",
    "t6o", "betay6", "betao6", "betayo6", "t7o", "betay7", "betao7", "betayo7", "fy9", "fo9", "fyo9",
    "ifrred", "y.CL", "o.CL"};
static const char *kAge2iNames[36] = {  // fit.paramnames, model-age-groups2i.R:538-549
    "betay0", "betao0", "betayo0", "betay1", "betao1", "betayo1", "betay2", "betao2", "betayo2",
    "y.CR", "y.CL", "y.HR", "y.HL", "y.DL", "o.CR", "o.CL", "o.HR", "o.HL", "o.DL", "t0_morts", "t0o",
    "CRsd", "HLsd", "DLsd", "t3o", "t4o", "betay4", "betao4", "betayo4", "t6o", "betay6", "betao6", "betayo6",
    "fy7", "fo7", "fyo7"};
static const char *kAge2xNames[52] = {  // fit.paramnames, model-age-groups2x.R:683-698
    "betay0", "betao0", "betayo0", "betay1", "betao1", "betayo1", "betay2", "betao2", "betayo2",
    "y.CR", "y.HR", "y.HL", "y.DL", "o.CR", "o.HR", "o.HL", "o.DL", "t0_morts", "t0o", "HLsd", "DLsd",
    "fyifr", "fyhr", "t3o", "betay3", "betao3", "betayo3", "t4o", "betay4", "betao4", "betayo4",
    "betay5", "betao5", "betayo5", "t6o", "betay6", "betao6", "betayo6", "t7o", "betay7", "betao7", "betayo7",
    "betay9", "betao9", "betayo9", "ifrred", "y.CL", "o.CL", "t10o", "betay11", "betao11", "betayo11"};
static const char *kSimpleNames[9] = {  // model-simple-ode-drj-cmp.R:259-260 (+ Nef, README.md:91-93)
    "R0", "Rt", "Tinf0", "Tinft", "logHR", "HL", "DL", "lockdownmort", "Nef"};

static void df_params_host(int flavour, const epi_data &D, double *mn, double *mx, double *init) {
  if (flavour == EPI_FLAVOUR_AGE2Q) {  // model-age-groups2q.R:624-663
    const double g = age_gamma();
    const double lo = D.lockdown_offset, ltp = D.lockdown_transition_period;
    const double i_[45] = {2.7, 1.7, 1.7, 2.3, 0.8, 0.8, 0.6, 0.4, 0.1, 2000, 1, 10, 15, 50, 2.5, 10, 21,
                           16, -10, 5, 5, 0.15, 0.5, D.d3 - lo - ltp, (double)(D.d4 - D.d3), 1, 0.2, 0.05,
                           0.7, 0.2, 0.01, (double)(D.d6 - D.d5), 1.1, 0.4, 0.03, (double)(D.d7 - D.d6), 1.4, 0.4, 0.08,
                           0.85, 0.85, 0.85, 0.4, 10, 10};
    const double mn_[45] = {2 * g, 2 * g, 2 * g, 1 * g, 1 * g, 1 * g, 0.05 * g, 0.01 * g, 0.01 * g,
                            3, 0.5, 6, 10, 3, 0.5, 6, 10, 0, -30, 2, 2, 0.05, 0.05, 60,
                            10, 0.05 * g, 0.01 * g, 0.002 * g, 0.05 * g, 0.01 * g, 0.002 * g,
                            10, 0.05 * g, 0.01 * g, 0.002 * g, 10, 0.05 * g, 0.01 * g, 0.002 * g,
                            0.6, 0.6, 0.6, 0.1, 4, 4};
    const double mx_[45] = {8 * g, 8 * g, 8 * g, 5 * g, 5 * g, 5 * g, 2 * g, 2 * g, 2 * g,
                            5000, 1.5, 15, 25, 100, 6, 15, 25,
                            std::fmax(D.dmort_last / 10, D.total_deaths_at_lockdown * 10),
                            30, 9, 9, 1.5, 1.5, 90,
                            50, 3 * g, 3 * g, 3 * g, 1.5 * g, 1.5 * g, 0.5 * g,
                            50, 3 * g, 3 * g, 0.5 * g, 50, 3 * g, 3 * g, 0.5 * g,
                            1.1, 1.1, 1.1, 0.7, 14, 14};
    for (int i = 0; i < 45; ++i) { mn[i] = mn_[i]; mx[i] = mx_[i]; init[i] = i_[i]; }
  } else if (flavour == EPI_FLAVOUR_AGE2X) {  // model-age-groups2x.R:705-751
    const double g = age_gamma();
    const double i_[52] = {3.1, 0.52, 0.50, 1.5, 0.50, 0.44, 0.8, 0.35, 0.09, 550, 0.75, 14, 21, 11, 1.5, 14, 21,
                           14, -8, 5.2, 5.8, 0.49, 0.30, 92, 1.0, 0.18, 0.02, 29, 1.4, 0.4, 0.04, 0.96, 0.27, 0.007,
                           28, 1.2, 0.9, 0.02, 32, 2.0, 0.6, 0.10, 1.3, 0.3, 0.05, 0.34, 6, 11, 245, 1.9, 0.3, 0.04};
    const double mn_[52] = {2 * g, 0.5 * g, 0.5 * g, 1 * g, 0.5 * g, 0.5 * g, 0.1 * g, 0.1 * g, 0.01 * g,
                            500, 0.5, 9, 14, 3, 0.5, 9, 14, 0, -20, 2, 2, 0.05, 0.05,
                            60, 0.5 * g, 0.01 * g, 0.002 * g, 10, 0.5 * g, 0.01 * g, 0.002 * g,
                            0.1 * g, 0.01 * g, 0.002 * g, 10, 0.2 * g, 0.01 * g, 0.002 * g,
                            10, 0.2 * g, 0.01 * g, 0.002 * g, 0.2 * g, 0.01 * g, 0.002 * g, 0.1, 4, 4,
                            (double)D.d10 - 10, 0.2 * g, 0.01 * g, 0.002 * g};
    const double tdl = D.total_deaths_at_lockdown * 10, dml = D.dmort_last / 10;
    const double mx_[52] = {8 * g, 5 * g, 5 * g, 5 * g, 3 * g, 3 * g, 2 * g, 2 * g, 2 * g,
                            5000, 2, 18, 25, 100, 6, 18, 25, dml > tdl ? dml : tdl, 10, 9, 9, 1, 1,
                            100, 3 * g, 1 * g, 0.5 * g, 50, 3 * g, 1 * g, 0.5 * g,
                            2 * g, 1 * g, 0.5 * g, 50, 4 * g, 3 * g, 0.5 * g,
                            50, 4 * g, 3 * g, 0.5 * g, 3 * g, 3 * g, 0.5 * g, 0.7, 14, 17,
                            (double)D.duncertain - 2, 3 * g, 3 * g, 0.5 * g};
    for (int i = 0; i < 52; ++i) { mn[i] = mn_[i]; mx[i] = mx_[i]; init[i] = i_[i]; }
  } else if (flavour == EPI_FLAVOUR_AGE2I) {  // model-age-groups2i.R:552-587
    const double g = age_gamma();
    const double lo = D.lockdown_offset, ltp = D.lockdown_transition_period;
    const double i_[36] = {3.6 * g, 3.6 * g, 3.6 * g, 2.0 * g, 2.0 * g, 2.0 * g, 0.8 * g, 0.8 * g, 0.8 * g,
                           100, 10, 1, 10, 21, 20, 10, 1, 10, 21,
                           D.total_deaths_at_lockdown, -1, 5, 5, 5,
                           D.d3 - lo - ltp, (double)(D.d4 - D.d3), 0.8 * g, 0.8 * g, 0.8 * g,
                           4.7, 0.8 * g, 0.8 * g, 0.8 * g, 1.2, 1.01, 1.01};
    const double mn_[36] = {2 * g, 2 * g, 2 * g, 1 * g, 1 * g, 1 * g, 0.05 * g, 0.01 * g, 0.01 * g,
                            3, 3, 0.5, 3, 10, 3, 3, 0.5, 3, 10,
                            0, -30, 2, 2, 2, 60,
                            10, 0.05 * g, 0.01 * g, 0.01 * g,
                            3, 0.05 * g, 0.01 * g, 0.01 * g, 1, 1, 1};
    const double mx_[36] = {8 * g, 8 * g, 8 * g, 5 * g, 5 * g, 5 * g, 2 * g, 2 * g, 2 * g,
                            5000, 25, 2, 25, 50, 100, 25, 2, 25, 50,
                            std::fmax(D.dmort_last / 10, D.total_deaths_at_lockdown * 10),
                            30, 9, 9, 9, 90,
                            50, 3 * g, 3 * g, 3 * g,
                            12, 1.5 * g, 1.5 * g, 1.5 * g, 2, 6, 6};
    for (int i = 0; i < 36; ++i) { mn[i] = mn_[i]; mx[i] = mx_[i]; init[i] = i_[i]; }
  } else {  // model-simple-ode-drj-cmp.R:263-270
    const double tdl = D.total_deaths_at_lockdown;
    const double i_[9] = {2.9, 0.9, 3, 3, std::log(0.02), 10, 20, tdl, 0.7};
    const double mn_[9] = {0.1, 0.1, 0.1, 0.1, std::log(0.001), 5, 5, 0, 0.05};
    const double mx_[9] = {8, 8, 8, 8, std::log(0.5), 30, 40, std::fmax(D.dmort_last / 10, tdl * 10), 1.0};
    const int d = flavour == EPI_FLAVOUR_NE ? 9 : 8;
    for (int i = 0; i < d; ++i) { mn[i] = mn_[i]; mx[i] = mx_[i]; init[i] = i_[i]; }
  }
}

// ---------------------------------------------------------- likelihood table
struct NbCall {  // one sum(dnbinom(x[x1..x2], mu = pmax(floor, m[dstart+r1 .. dstart+r2]), size, log=T))
  int series, term;
  const double *x;
  int xlen, x1, x2;  // 1-based inclusive
  int r1, r2;        // model index relative to dstart
  const double *size;
  int ns;
};

static bool is_count(double x) { return x >= 0 && std::isfinite(x) && x == std::nearbyint(x); }

static int build_term_table(epi_handle *h, const std::vector<NbCall> &calls, const int *n_obs, int n_series) {
  std::vector<std::vector<std::vector<Slot>>> per(n_series);
  for (int s = 0; s < n_series; ++s) per[s].resize(n_obs[s]);
  long double cst[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (const NbCall &c : calls) {
    const int lx = c.x2 - c.x1 + 1, lm = c.r2 - c.r1 + 1;
    if (lx <= 0 || lm <= 0 || c.x1 < 1 || c.x2 > c.xlen || c.r1 < 0 || c.r2 >= n_obs[c.series])
      return epi_fail(h, EPI_ERR_ARG,
                      "inconsistent observation windows (series %d: x[%d..%d] of %d, model rel [%d..%d] of %d): "
                      "check d_reliable_* and the series lengths",
                      c.series, c.x1, c.x2, c.xlen, c.r1, c.r2, n_obs[c.series]);
    int len = lx > lm ? lx : lm;
    if (c.ns > len) len = c.ns;
    for (int k = 0; k < len; ++k) {  // R recycles every argument to the longest
      const double x = c.x[c.x1 - 1 + k % lx];
      const double size = c.size[k % c.ns];
      const int rel = c.r1 + k % lm;
      if (!is_count(x)) return epi_fail(h, EPI_ERR_ARG, "observed counts must be non-negative integers (got %g)", x);
      if (!(size > 0)) return epi_fail(h, EPI_ERR_ARG, "negative-binomial size must be > 0 (got %g)", size);
      // terms of one (series, day) that share a size merge into one slot:
      //   sum_k [x_k log(mu) - (x_k + size) log(size + mu)] = (sum x_k) log(mu) - (sum (x_k + size)) log(size + mu)
      bool merged = false;
      for (Slot &sl : per[c.series][rel])
        if (sl.size == size) { sl.x += (float)x; sl.n += 1.0f; merged = true; break; }
      if (!merged) per[c.series][rel].push_back(Slot{size, (float)x, 1.0f});
      cst[c.term] += lgammal((long double)x + size) - lgammal((long double)size) - lgammal((long double)x + 1.0L) +
                     (long double)size * logl((long double)size);
    }
  }
  for (int s = 0; s < n_series; ++s) {
    size_t K = 1;
    for (auto &v : per[s]) K = std::max(K, v.size());
    std::vector<Slot> flat((size_t)n_obs[s] * K, Slot{0.0, 0.0f, 0.0f});
    for (auto &v : per[s])
      for (Slot &sl : v)
        if (!(sl.x < 16777216.0f))
          return epi_fail(h, EPI_ERR_ARG, "observed counts of one day sum to %g >= 2^24: not representable in the term table", (double)sl.x);
    for (int r = 0; r < n_obs[s]; ++r)
      for (size_t k = 0; k < per[s][r].size(); ++k) flat[(size_t)r * K + k] = per[s][r][k];
    Slot *d = nullptr;
    CUDA_OK(h, cudaMalloc(&d, flat.size() * sizeof(Slot)));
    h->owned.push_back(d);
    CUDA_OK(h, cudaMemcpy(d, flat.data(), flat.size() * sizeof(Slot), cudaMemcpyHostToDevice));
    h->M.slots[s] = d;
    h->M.K[s] = (int)K;
    h->M.multi_begin[s] = n_obs[s];
    for (int r = n_obs[s] - 1; r >= 0; --r)
      if (per[s][r].size() > 1) h->M.multi_begin[s] = r;
    h->M.n_obs[s] = n_obs[s];
    h->n_terms += (int)flat.size();
  }
  for (int t = 0; t < 8; ++t) h->M.term_const[t] += (double)cst[t];
  return EPI_OK;
}

static int upload(epi_handle *h, const double *src, int n, const double **dst) {
  double *d = nullptr;
  CUDA_OK(h, cudaMalloc(&d, sizeof(double) * (size_t)n));
  h->owned.push_back(d);
  CUDA_OK(h, cudaMemcpy(d, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  *dst = d;
  return EPI_OK;
}

// ---------------------------------------------------------------- prior table
static PriorTerm pt_sum(int ia, int ib, double c0, double ca, double cb, double mean, double sd) {
  return PriorTerm{ia, ib, 0, 0, c0, ca, cb, mean, sd};
}
static PriorTerm pt_norm(int ia, double mean, double sd) { return pt_sum(ia, -1, 0, 1, 0, mean, sd); }
static PriorTerm pt_lnorm(int ia, double meanlog, double sdlog) { return PriorTerm{ia, -1, 0, 1, 0, 1, 0, meanlog, sdlog}; }
static PriorTerm pt_lratio(int ia, int ib, double sdlog) { return PriorTerm{ia, ib, 2, 1, 0, 0, 0, 0.0, sdlog}; }

static std::vector<PriorTerm> build_prior(int flavour, const epi_data &D) {
  std::vector<PriorTerm> p;
  if (flavour == EPI_FLAVOUR_AGE2Q) {  // calclogp, model-age-groups2q.R:433-485 (0-based theta indices)
    const double g = age_gamma(), lo_ltp = D.lockdown_offset + D.lockdown_transition_period;
    const double SD = std::log(2.0), lSD = std::log(1.5);
    p.push_back(pt_norm(18, 0, 10));
    p.push_back(pt_sum(23, -1, lo_ltp, 1, 0, D.d3, 10));
    p.push_back(pt_sum(23, 24, lo_ltp, 1, 1, D.d4, 10));
    p.push_back(pt_sum(31, -1, D.d5, 1, 0, D.d6, 5));
    p.push_back(pt_sum(31, 35, D.d5, 1, 1, D.d7, 4));
    const int r2[12][2] = {{0, 3}, {1, 4}, {2, 5}, {3, 6}, {4, 7}, {5, 8}, {6, 25}, {7, 26}, {8, 27}, {28, 32}, {29, 33}, {30, 34}};
    for (auto &r : r2) p.push_back(pt_lratio(r[0], r[1], SD));
    const int r15[3][2] = {{32, 36}, {33, 37}, {34, 38}};
    for (auto &r : r15) p.push_back(pt_lratio(r[0], r[1], lSD));
    p.push_back(pt_norm(37, 0.5 * g, 0.2 * g));
    p.push_back(PriorTerm{37, 40, 1, 0, 0, 0, 0, 0.5 * g, 0.2 * g});  // betao8 * fo9
    for (int i : {39, 40, 41}) p.push_back(pt_lnorm(i, std::log(0.85), std::log(1.07)));
    p.push_back(pt_norm(19, 5, 1));
    p.push_back(pt_norm(20, 5, 1));
    p.push_back(pt_norm(43, 10, 3));
    p.push_back(pt_norm(44, 10, 3));
    p.push_back(pt_norm(11, 13, 3));
    p.push_back(pt_norm(15, 13, 3));
    p.push_back(pt_norm(12, 21, 0.5));
    p.push_back(pt_norm(16, 21, 0.5));
    p.push_back(pt_lnorm(21, std::log(0.3), std::log(2.0)));
    p.push_back(pt_lnorm(22, std::log(0.3), std::log(2.0)));
    p.push_back(pt_lnorm(42, std::log(0.35), std::log(1.3)));
    p.push_back(pt_lnorm(10, 0, lSD));
    p.push_back(pt_lnorm(14, 0, lSD));
  } else if (flavour == EPI_FLAVOUR_AGE2X) {  // calclogp, model-age-groups2x.R:485-543 (0-based theta indices)
    const double lo_ltp = D.lockdown_offset + D.lockdown_transition_period;
    const double SD = std::log(2.0), lSD = std::log(1.3), llSD = std::log(1.2);
    p.push_back(pt_norm(18, 0, 10));
    p.push_back(pt_sum(23, -1, lo_ltp, 1, 0, D.d3, 10));
    p.push_back(pt_sum(23, 27, lo_ltp, 1, 1, D.d4, 10));
    p.push_back(pt_sum(34, -1, D.d5, 1, 0, D.d6, 5));
    p.push_back(pt_sum(34, 38, D.d5, 1, 1, (double)D.d7 - D.lockdown_offset, 2));
    p.push_back(pt_norm(48, D.d10, 3));
    // beta triples (0-based first index): 0: 0, 1: 3, 2: 6, 3: 24, 4: 28, 5: 31, 6: 35, 7: 39, 9: 42, 11: 49
    for (int c = 0; c < 3; ++c) p.push_back(pt_lratio(0 + c, 3 + c, llSD));
    const int pr[5][2] = {{3, 6}, {6, 24}, {24, 28}, {31, 35}, {35, 39}};
    for (auto &r : pr)
      for (int c = 0; c < 3; ++c) p.push_back(pt_lratio(r[0] + c, r[1] + c, SD));
    for (int c = 0; c < 3; ++c) p.push_back(PriorTerm{39 + c, 42 + c, 2, 1, 0, 0, 0, std::log(1.3), lSD});
    for (int c = 0; c < 3; ++c) p.push_back(pt_lratio(42 + c, 49 + c, lSD));
    p.push_back(pt_norm(19, 4, 1));
    p.push_back(pt_norm(20, 5, 1));
    p.push_back(pt_norm(46, 7, 3));
    p.push_back(pt_norm(47, 7, 3));
    p.push_back(pt_norm(11, 13, 1));
    p.push_back(pt_norm(15, 13, 1));
    p.push_back(pt_norm(12, 21, 0.5));
    p.push_back(pt_norm(16, 21, 0.5));
    p.push_back(pt_lnorm(21, std::log(0.6), std::log(1.5)));
    p.push_back(pt_lnorm(45, std::log(0.35), std::log(1.3)));
    p.push_back(pt_lnorm(10, 0, lSD));
    p.push_back(pt_lnorm(14, 0, lSD));
  } else if (flavour == EPI_FLAVOUR_AGE2I) {  // calclogp, model-age-groups2i.R:410-435 (0-based theta indices)
    const double g = age_gamma(), lo_ltp = D.lockdown_offset + D.lockdown_transition_period;
    p.push_back(pt_norm(20, 0, 10));
    p.push_back(pt_sum(24, -1, lo_ltp, 1, 0, D.d3, 10));
    p.push_back(pt_sum(24, 25, lo_ltp, 1, 1, D.d4, 10));
    p.push_back(pt_norm(29, 4.7, 3));  // dnorm(t6o, mean = G, sd = 3)
    const int df2[6][2] = {{0, 3}, {1, 4}, {2, 5}, {3, 6}, {4, 7}, {5, 8}};
    for (auto &r : df2) p.push_back(pt_sum(r[0], r[1], 0, 1, -1, 0, 2 * g));
    const int df05[3][2] = {{6, 26}, {7, 27}, {8, 28}};  // beta3 = beta2
    for (auto &r : df05) p.push_back(pt_sum(r[0], r[1], 0, 1, -1, 0, 0.5 * g));
    for (int i : {21, 22, 23}) p.push_back(pt_norm(i, 5, 1));
    p.push_back(pt_norm(13, 21, 4));
    p.push_back(pt_norm(18, 21, 4));
    p.push_back(pt_norm(30, 0.6, 0.2));
    p.push_back(pt_norm(31, 0.22, 0.2));
    p.push_back(pt_norm(32, 0.26, 0.2));
    p.push_back(PriorTerm{33, 30, 1, 0, 0, 0, 0, 0.85, 0.2});   // betay7 = fy7 * betay6
    p.push_back(PriorTerm{35, 31, 1, 0, 0, 0, 0, 0.19, 0.05});  // betao7 = params[36] * betao6
    p.push_back(PriorTerm{34, 32, 1, 0, 0, 0, 0, 0.02, 0.02});  // betayo7 = params[35] * betayo6
  } else {  // calclogp, model-simple-ode-drj-cmp.R:156-179: G0 = Tinc + Tinf0/2, Gt = Tinc + Tinft/2
    p.push_back(pt_norm(6, 21, 4));
    p.push_back(pt_sum(2, -1, 3.0, 0.5, 0, 5, 1));
    p.push_back(pt_sum(2, 3, 0.0, 0.5, -0.5, 0, 1.5));
  }
  return p;
}

// ------------------------------------------------------------ schedule table
// The beta(t) schedule of an age model file as data (Schedule, epi_device.cuh): breakpoints as running sums of
// data constants and parameters, the beta triple of every breakpoint, which segments interpolate linearly.
// 0-based theta indices; data_offset is added on the device.
static void sched_bp(Schedule &S, int k, double c, int ia = -1, int ib = -1, bool clamp = false, double post = 0.0) {
  S.bp[k] = SchedBp{c, post, (short)ia, (short)ib, (short)(clamp ? 1 : 0), 0};
}
static void sched_beta(Schedule &S, int k, int first, int my = -1, int mo = -1, int myo = -1) {
  S.beta[k] = SchedBeta{{(short)first, (short)(first + 1), (short)(first + 2)}, {(short)my, (short)mo, (short)myo}};
}
static void build_schedule(int flavour, const epi_data &D, Schedule &S) {
  std::memset(&S, 0, sizeof S);
  const double lo = D.lockdown_offset, t2 = lo + D.lockdown_transition_period;
  if (flavour == EPI_FLAVOUR_AGE2Q) {
    // breakpoints, model-age-groups2q.R:184-197; segment kinds, model-agen.cpp:78-99; betas, :52-99
    S.n = 10; S.linear = 0x10F;
    sched_bp(S, 0, lo, 18);                 // t0 = data_offset + lockdown_offset + t0o
    sched_bp(S, 1, lo, 18, -1, true);       // t1 = ... + max(t0o, 0)
    sched_bp(S, 2, t2);
    sched_bp(S, 3, t2, 23);                 // t3 = t2 + t3o
    sched_bp(S, 4, t2, 23, 24);             // t4 = t3 + t4o
    sched_bp(S, 5, D.d5);
    sched_bp(S, 6, D.d5, 31);               // t6 = t5 + t6o
    sched_bp(S, 7, D.d5, 31, 35);           // t7 = t6 + t7o
    sched_bp(S, 8, D.d8);
    sched_bp(S, 9, (double)D.d8 + 7);       // t9 = t8 + 7
    sched_beta(S, 0, 0); sched_beta(S, 1, 3); sched_beta(S, 2, 6);
    S.beta[3] = SchedBeta{{6, 26, 27}, {-1, -1, -1}};  // betay3 = betay2, betao3 = betao4, betayo3 = betayo4
    sched_beta(S, 4, 25); sched_beta(S, 5, 28); sched_beta(S, 6, 32);
    sched_beta(S, 7, 36); sched_beta(S, 8, 36);        // beta8 = beta7
    sched_beta(S, 9, 36, 39, 40, 41);                  // beta9 = f9 * beta8
  } else if (flavour == EPI_FLAVOUR_AGE2I) {
    // model-age-groups2i.R:169-183, model-agef.cpp:66-87, :52-93
    S.n = 8; S.linear = 0x03F;
    sched_bp(S, 0, lo, 20);
    sched_bp(S, 1, lo, 20, -1, true);
    sched_bp(S, 2, t2);
    sched_bp(S, 3, t2, 24);
    sched_bp(S, 4, t2, 24, 25);
    sched_bp(S, 5, D.d5);
    sched_bp(S, 6, D.d5, 29);
    sched_bp(S, 7, D.d7);
    sched_beta(S, 0, 0); sched_beta(S, 1, 3); sched_beta(S, 2, 6); sched_beta(S, 3, 6);  // beta3 = beta2
    sched_beta(S, 4, 26); sched_beta(S, 5, 26);                                          // beta5 = beta4
    sched_beta(S, 6, 30);
    sched_beta(S, 7, 30, 33, 35, 34);  // betay7 = fy7 betay6, betao7 = params[36] betao6, betayo7 = params[35] betayo6 (:91-93)
  } else if (flavour == EPI_FLAVOUR_AGE2X) {
    // model-age-groups2x.R:206-219, model-ager.cpp:88-124, :56-119
    S.n = 13; S.linear = 0xD0F;
    sched_bp(S, 0, lo, 18);
    sched_bp(S, 1, lo, 18, -1, true);
    sched_bp(S, 2, t2);
    sched_bp(S, 3, t2, 23);
    sched_bp(S, 4, t2, 23, 27);
    sched_bp(S, 5, D.d5);
    sched_bp(S, 6, D.d5, 34);
    sched_bp(S, 7, D.d5, 34, 38);
    sched_bp(S, 8, D.d8);
    sched_bp(S, 9, D.d9);
    sched_bp(S, 10, 0.0, 48);                    // t10 = data_offset + t10o
    sched_bp(S, 11, 0.0, 48, -1, false, 3.0);    // t11 = t10 + 3
    sched_bp(S, 12, D.d12);
    const int first[13] = {0, 3, 6, 24, 28, 31, 35, 39, 39, 42, 42, 49, 42};  // beta8 = beta7, beta10 = beta12 = beta9
    for (int k = 0; k < 13; ++k) sched_beta(S, k, first[k]);
  }
}

// -------------------------------------------------------------------- create
static int fill_model(epi_handle *h, const epi_model_desc &desc, const epi_data &D) {
  DevModel &M = h->M;
  std::memset(&M, 0, sizeof M);
  M.flavour = desc.flavour;
  M.P = desc.period;
  M.m = desc.substeps;
  M.lockdown_offset = D.lockdown_offset;
  M.ltp = D.lockdown_transition_period;
  std::vector<NbCall> calls;
  int n_obs[kMaxSeries] = {0, 0, 0, 0, 0};
  if (is_age(desc.flavour)) {
    Schedule S;
    build_schedule(desc.flavour, D, S);
    Schedule *dS = nullptr;
    CUDA_OK(h, cudaMalloc(&dS, sizeof(Schedule)));
    h->owned.push_back(dS);
    CUDA_OK(h, cudaMemcpy(dS, &S, sizeof(Schedule), cudaMemcpyHostToDevice));
    M.sched = dS;
  }
  if (desc.flavour == EPI_FLAVOUR_AGE2Q) {
    M.d = 45;
    M.P1 = 60;  // period1, model-age-groups2q.R:160
    if (desc.period < 60) return epi_fail(h, EPI_ERR_ARG, "age-2q needs period >= 60 (pass 1 integrates 60 days)");
    if (!(D.y_N > 1 && D.o_N > 0)) return epi_fail(h, EPI_ERR_ARG, "y_N / o_N missing");
    if (D.n_rate < 1 || !D.y_ifr || !D.o_ifr || !D.y_hr || !D.o_hr || !D.g || !D.gtrimp)
      return epi_fail(h, EPI_ERR_ARG, "rate vectors y_ifr/o_ifr/y_hr/o_hr/g/gtrimp missing");
    if (!D.y_dcasei || !D.o_dcasei || !D.dhospi || !D.y_dmorti || !D.o_dmorti || !D.o_wdmorti)
      return epi_fail(h, EPI_ERR_ARG, "observed series missing");
    M.Ny = D.y_N; M.No = D.o_N; M.N = D.y_N + D.o_N;
    M.a1 = 1.0 / 3.0; M.a2 = 1.0 / 1.0; M.gamma = age_gamma();
    M.ifr_y1 = D.y_ifr[0]; M.ifr_o1 = D.o_ifr[0];
    M.d5 = D.d5; M.d8 = D.d8;
    M.n_rate = D.n_rate;
    M.d_hosp_o1 = D.d_hosp_o1; M.d_hosp_o2 = D.d_hosp_o2;
    int rc;
    if ((rc = upload(h, D.y_ifr, D.n_rate, &M.y_ifr))) return rc;
    if ((rc = upload(h, D.o_ifr, D.n_rate, &M.o_ifr))) return rc;
    if ((rc = upload(h, D.y_hr, D.n_rate, &M.y_hr))) return rc;
    if ((rc = upload(h, D.o_hr, D.n_rate, &M.o_hr))) return rc;
    if ((rc = upload(h, D.g, D.n_rate, &M.g))) return rc;
    if ((rc = upload(h, D.gtrimp, D.n_rate, &M.gtrimp))) return rc;
    M.n_series = 5;
    const int nc = D.n_dcasei, nh = D.n_dhospi, nhh = D.n_dhosp, nm = D.n_dmorti;
    const int r = D.d_reliable_cases, rh = D.d_reliable_hosp;
    if (D.d_hosp_o1 < 0 || D.d_hosp_o2 < D.d_hosp_o1 || D.d_hosp_o2 >= nh)
      return epi_fail(h, EPI_ERR_ARG, "d_hosp_o1/o2 outside the hospital series");
    n_obs[0] = nc; n_obs[1] = nc; n_obs[2] = nh; n_obs[3] = nm; n_obs[4] = nm;
    const int term_of[5] = {0, 1, 2, 4, 5};
    const double fl[5] = {0.1, 0.1, 0.1, 0.001, 0.001};
    for (int s = 0; s < 5; ++s) { M.term_of[s] = term_of[s]; M.floor_[s] = fl[s]; }
    h->sizes = {D.case_nbinom_size1, D.case_nbinom_sizey2, D.case_nbinom_sizeo2, D.hosp_nbinom_size1,
                D.hosp_nbinom_size2, D.hosp_nbinom_size2 * D.hosp_last7, D.mort_nbinom_size * 2};
    h->size_vec.resize(nm);
    for (int i = 0; i < nm; ++i) h->size_vec[i] = D.o_wdmorti[i] * D.mort_nbinom_size;
    const double *sz = h->sizes.data();
    // cases, model-age-groups2q.R:514-526 (note the x/mu length mismatch -> recycling)
    calls.push_back({0, 0, D.y_dcasei, nc, 1, r - 1, 0, r, sz + 0, 1});
    calls.push_back({0, 0, D.y_dcasei, nc, r, nc, r + 1, nc - 1, sz + 1, 1});
    calls.push_back({1, 1, D.o_dcasei, nc, 1, r - 1, 0, r, sz + 0, 1});
    calls.push_back({1, 1, D.o_dcasei, nc, r, nc, r + 1, nc - 1, sz + 2, 1});
    // hosp, :544-554
    calls.push_back({2, 2, D.dhospi, nh, 1, rh - 1, 0, rh, sz + 3, 1});
    calls.push_back({2, 2, D.dhospi, nh, rh, nhh, rh + 1, nh - 1, sz + 4, 1});
    calls.push_back({2, 2, D.dhospi, nh, nhh - 7, nhh, nh - 1 - 7, nh - 1, sz + 5, 1});
    // deaths, :584-590
    calls.push_back({3, 4, D.y_dmorti, nm, 1, nm, 0, nm - 1, sz + 6, 1});
    calls.push_back({4, 5, D.o_dmorti, nm, 1, nm, 0, nm - 1, h->size_vec.data(), nm});
    M.d3 = D.d3; M.d4 = D.d4; M.d5d = D.d5; M.d6 = D.d6; M.d7 = D.d7;
    M.lo_ltp = D.lockdown_offset + D.lockdown_transition_period;
  } else if (desc.flavour == EPI_FLAVOUR_AGE2X) {
    M.d = 52;
    M.P1 = 60;  // period1, model-age-groups2x.R:183
    if (desc.period < 60) return epi_fail(h, EPI_ERR_ARG, "age-2x needs period >= 60 (pass 1 integrates 60 days)");
    if (!(D.y_N > 1 && D.o_N > 0)) return epi_fail(h, EPI_ERR_ARG, "y_N / o_N missing");
    if (D.n_rate < 1 || !D.y_ifr || !D.o_ifr || !D.y_hr || !D.o_hr || !D.g || !D.gtrimp)
      return epi_fail(h, EPI_ERR_ARG, "rate vectors y_ifr/o_ifr/y_hr/o_hr/g/gtrimp missing");
    if (!D.y_dcasei || !D.o_dcasei || !D.dhospi || !D.y_dmorti || !D.o_dmorti || !D.o_wdmorti)
      return epi_fail(h, EPI_ERR_ARG, "observed series missing");
    if (!(D.symp_cases_factor > 0) || D.d_symp_cases < 1 || D.d_all_cases < 1 || D.d9 < 1 || D.d10 < 1 || D.d12 < 1 || D.duncertain < 1)
      return epi_fail(h, EPI_ERR_ARG, "age-2x globals missing (d9, d10, d12, duncertain, d_symp_cases, d_all_cases, symp_cases_factor)");
    M.Ny = D.y_N; M.No = D.o_N; M.N = D.y_N + D.o_N;
    M.a1 = 1.0 / 3.0; M.a2 = 1.0 / 1.0; M.gamma = age_gamma();
    M.eta = 1.0 / (0.75 * 356);  // model-age-groups2x.R:19 (sic)
    M.ifr_y1 = D.y_ifr[0]; M.ifr_o1 = D.o_ifr[0];
    M.d5 = D.d5; M.d8 = D.d8; M.d9 = D.d9; M.d12 = D.d12;
    M.d_symp = D.d_symp_cases; M.d_all = D.d_all_cases; M.symp_factor = D.symp_cases_factor;
    M.n_rate = D.n_rate;
    M.d_hosp_o1 = D.d_hosp_o1; M.d_hosp_o2 = D.d_hosp_o2;
    int rc;
    if ((rc = upload(h, D.y_ifr, D.n_rate, &M.y_ifr))) return rc;
    if ((rc = upload(h, D.o_ifr, D.n_rate, &M.o_ifr))) return rc;
    if ((rc = upload(h, D.y_hr, D.n_rate, &M.y_hr))) return rc;
    if ((rc = upload(h, D.o_hr, D.n_rate, &M.o_hr))) return rc;
    if ((rc = upload(h, D.g, D.n_rate, &M.g))) return rc;
    if ((rc = upload(h, D.gtrimp, D.n_rate, &M.gtrimp))) return rc;
    M.n_series = 5;
    const int nc = D.n_dcasei, nh = D.n_dhospi, nhh = D.n_dhosp, nm = D.n_dmorti;
    const int r = D.d_reliable_cases, rh = D.d_reliable_hosp;
    if (D.d_hosp_o1 < 0 || D.d_hosp_o2 < D.d_hosp_o1 || D.d_hosp_o2 >= nh)
      return epi_fail(h, EPI_ERR_ARG, "d_hosp_o1/o2 outside the hospital series");
    n_obs[0] = nc; n_obs[1] = nc; n_obs[2] = nh; n_obs[3] = nm; n_obs[4] = nm;
    const int term_of[5] = {0, 1, 2, 4, 5};
    const double fl[5] = {0.1, 0.1, 0.1, 0.001, 0.001};
    for (int s = 0; s < 5; ++s) { M.term_of[s] = term_of[s]; M.floor_[s] = fl[s]; }
    h->sizes = {D.case_nbinom_size1, D.case_nbinom_sizey2, D.case_nbinom_sizeo2, D.hosp_nbinom_size1,
                D.hosp_nbinom_size2, D.hosp_nbinom_size2 * D.hosp_last7, D.mort_nbinom_size * 5,
                D.case_nbinom_sizeo2 * D.hosp_last7, D.mort_nbinom_size * D.hosp_last7};
    h->size_vec.resize(nm);
    for (int i = 0; i < nm; ++i) h->size_vec[i] = D.o_wdmorti[i] * D.mort_nbinom_size;
    const double *sz = h->sizes.data();
    // cases, model-age-groups2x.R:572-588 (the x/mu length mismatch of 2q, plus the last eight days of o.dcasei again)
    calls.push_back({0, 0, D.y_dcasei, nc, 1, r - 1, 0, r, sz + 0, 1});
    calls.push_back({0, 0, D.y_dcasei, nc, r, nc, r + 1, nc - 1, sz + 1, 1});
    calls.push_back({1, 1, D.o_dcasei, nc, 1, r - 1, 0, r, sz + 0, 1});
    calls.push_back({1, 1, D.o_dcasei, nc, r, nc, r + 1, nc - 1, sz + 2, 1});
    calls.push_back({1, 1, D.o_dcasei, nc, nc - 7, nc, nc - 1 - 7, nc - 1, sz + 7, 1});
    // hosp, :606-617
    calls.push_back({2, 2, D.dhospi, nh, 1, rh - 1, 0, rh, sz + 3, 1});
    calls.push_back({2, 2, D.dhospi, nh, rh, nhh, rh + 1, nh - 1, sz + 4, 1});
    calls.push_back({2, 2, D.dhospi, nh, nhh - 7, nhh, nh - 1 - 7, nh - 1, sz + 5, 1});
    // deaths, :646-655
    calls.push_back({3, 4, D.y_dmorti, nm, 1, nm, 0, nm - 1, sz + 6, 1});
    calls.push_back({4, 5, D.o_dmorti, nm, 1, nm, 0, nm - 1, h->size_vec.data(), nm});
    calls.push_back({4, 5, D.o_dmorti, nm, nm - 7, nm, nm - 1 - 7, nm - 1, sz + 8, 1});
    M.d3 = D.d3; M.d4 = D.d4; M.d5d = D.d5; M.d6 = D.d6; M.d7 = D.d7;
    M.lo_ltp = D.lockdown_offset + D.lockdown_transition_period;
  } else if (desc.flavour == EPI_FLAVOUR_AGE2I) {
    M.d = 36;
    M.P1 = desc.period;  // pass 1 covers the whole period, model-age-groups2i.R:148-152
    if (!(D.y_N > 1 && D.o_N > 0)) return epi_fail(h, EPI_ERR_ARG, "y_N / o_N missing");
    if (D.n_rate < 1 || !D.y_ifr || !D.o_ifr || !D.y_hr || !D.o_hr)
      return epi_fail(h, EPI_ERR_ARG, "rate vectors y_ifr/o_ifr/y_hr/o_hr missing");
    if (!D.y_dcasei || !D.o_dcasei || !D.dhospi || !D.y_dmorti || !D.o_dmorti)
      return epi_fail(h, EPI_ERR_ARG, "observed series missing");
    M.Ny = D.y_N; M.No = D.o_N; M.N = D.y_N + D.o_N;
    M.a1 = 1.0 / 3.0; M.a2 = 1.0 / 1.0; M.gamma = age_gamma();
    M.ifr_y1 = D.y_ifr[0]; M.ifr_o1 = D.o_ifr[0];
    M.d5 = D.d5;
    M.n_rate = D.n_rate;
    int rc;
    if ((rc = upload(h, D.y_ifr, D.n_rate, &M.y_ifr))) return rc;
    if ((rc = upload(h, D.o_ifr, D.n_rate, &M.o_ifr))) return rc;
    if ((rc = upload(h, D.y_hr, D.n_rate, &M.y_hr))) return rc;
    if ((rc = upload(h, D.o_hr, D.n_rate, &M.o_hr))) return rc;
    M.n_series = 5;
    const int nc = D.n_dcasei, nh = D.n_dhospi, nm = D.n_dmorti;
    const int r = D.d_reliable_cases;
    n_obs[0] = nc; n_obs[1] = nc; n_obs[2] = nh; n_obs[3] = nm; n_obs[4] = nm;
    const int term_of[5] = {0, 1, 2, 4, 5};
    const double fl[5] = {0.1, 0.1, 0.001, 0.001, 0.001};  // :471,496,516
    for (int s = 0; s < 5; ++s) { M.term_of[s] = term_of[s]; M.floor_[s] = fl[s]; }
    h->sizes = {D.case_nbinom_size1, D.case_nbinom_sizey2, D.case_nbinom_sizeo2, D.hosp_nbinom_size, D.mort_nbinom_size};
    const double *sz = h->sizes.data();
    // cases, model-age-groups2i.R:466-478 (same x/mu length mismatch as 2q)
    calls.push_back({0, 0, D.y_dcasei, nc, 1, r - 1, 0, r, sz + 0, 1});
    calls.push_back({0, 0, D.y_dcasei, nc, r, nc, r + 1, nc - 1, sz + 1, 1});
    calls.push_back({1, 1, D.o_dcasei, nc, 1, r - 1, 0, r, sz + 0, 1});
    calls.push_back({1, 1, D.o_dcasei, nc, r, nc, r + 1, nc - 1, sz + 2, 1});
    calls.push_back({2, 2, D.dhospi, nh, 1, nh, 0, nh - 1, sz + 3, 1});     // :496-498
    calls.push_back({3, 4, D.y_dmorti, nm, 1, nm, 0, nm - 1, sz + 4, 1});  // :516-522
    calls.push_back({4, 5, D.o_dmorti, nm, 1, nm, 0, nm - 1, sz + 4, 1});
    M.d3 = D.d3; M.d4 = D.d4; M.d5d = D.d5; M.d7 = D.d7;
    M.lo_ltp = D.lockdown_offset + D.lockdown_transition_period;
  } else if (desc.flavour == EPI_FLAVOUR_SIMPLE || desc.flavour == EPI_FLAVOUR_NE) {
    M.d = desc.flavour == EPI_FLAVOUR_NE ? 9 : 8;
    M.P1 = desc.period;  // pass 1 covers the whole period, model-simple-ode-drj-cmp.R:82-86
    if (!(D.N > 1)) return epi_fail(h, EPI_ERR_ARG, "N missing");
    if (!D.dhospi || !D.dmorti || D.n_dhospi < 1 || D.n_dmorti < 1) return epi_fail(h, EPI_ERR_ARG, "dhospi/dmorti missing");
    M.N = D.N;
    M.a_inc = 1.0 / 3.0;
    M.n_series = 2;
    n_obs[0] = D.n_dhospi; n_obs[1] = D.n_dmorti;
    M.term_of[0] = 1; M.term_of[1] = 2;
    M.floor_[0] = 0.1; M.floor_[1] = 0.1;
    h->sizes = {D.hosp_nbinom_size, D.mort_nbinom_size};
    const double *sz = h->sizes.data();
    calls.push_back({0, 1, D.dhospi, D.n_dhospi, 1, D.n_dhospi, 0, D.n_dhospi - 1, sz + 0, 1});
    calls.push_back({1, 2, D.dmorti, D.n_dmorti, 1, D.n_dmorti, 0, D.n_dmorti - 1, sz + 1, 1});
    // loglLD, :203-204
    const double x = D.total_deaths_at_lockdown, size = D.mort_nbinom_size;
    if (!is_count(x) || !(size > 0)) return epi_fail(h, EPI_ERR_ARG, "total_deaths_at_lockdown must be a count, sizes > 0");
    M.tdl = x; M.mort_size = size;
    M.term_const[0] = (double)(lgammal((long double)x + size) - lgammal((long double)size) - lgammal((long double)x + 1.0L) +
                               (long double)size * logl((long double)size));
  } else {
    return epi_fail(h, EPI_ERR_ARG, "unknown flavour %d", desc.flavour);
  }
  if (desc.period < 2 || desc.period > (is_age(desc.flavour) ? 2048 : 4096))
    return epi_fail(h, EPI_ERR_ARG, "period out of range (2..4096, age flavours 2..2048)");
  if (desc.substeps < 1 || desc.substeps > 64) return epi_fail(h, EPI_ERR_ARG, "substeps out of range");
  for (int s = 0; s < M.n_series; ++s)
    if (n_obs[s] > desc.period) return epi_fail(h, EPI_ERR_ARG, "observed series longer than FitTotalPeriod");
  int rc = build_term_table(h, calls, n_obs, M.n_series);
  if (rc) return rc;
  df_params_host(desc.flavour, D, h->box_min, h->box_max, h->init);
  if ((rc = upload(h, h->box_min, M.d, &M.box_min))) return rc;
  if ((rc = upload(h, h->box_max, M.d, &M.box_max))) return rc;
  std::vector<PriorTerm> pr = build_prior(desc.flavour, D);
  for (PriorTerm &t : pr) t.log_sd = std::log(t.sd);
  PriorTerm *dp = nullptr;
  CUDA_OK(h, cudaMalloc(&dp, pr.size() * sizeof(PriorTerm)));
  h->owned.push_back(dp);
  CUDA_OK(h, cudaMemcpy(dp, pr.data(), pr.size() * sizeof(PriorTerm), cudaMemcpyHostToDevice));
  h->PT = dp;
  h->n_prior = (int)pr.size();
  return EPI_OK;
}

extern "C" int epi_abi_version(void) { return EPI_ABI_VERSION; }

extern "C" int epi_create(const epi_model_desc *desc, const epi_data *data, int device, epi_handle **out) {
  if (out) *out = nullptr;
  if (!desc || !data || !out) return epi_fail(nullptr, EPI_ERR_ARG, "null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return epi_fail(nullptr, EPI_ERR_NO_DEVICE, "no CUDA device (%s); libepimcmc has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return epi_fail(nullptr, EPI_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
  epi_handle *h = new epi_handle();
  h->device = device;
  h->desc = *desc;
  int rc = EPI_OK;
  do {
    if (cudaSetDevice(device) != cudaSuccess) { rc = epi_fail(h, EPI_ERR_CUDA, "cudaSetDevice(%d) failed", device); break; }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = epi_fail(h, EPI_ERR_CUDA, "stream"); break; }
    if (cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) != cudaSuccess) { rc = epi_fail(h, EPI_ERR_CUDA, "stream"); break; }
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    for (int k = 0; k < epi_handle::kMaxWays - 1; ++k) {
      if (cudaStreamCreateWithFlags(&h->split_stream[k], cudaStreamNonBlocking) != cudaSuccess) { rc = epi_fail(h, EPI_ERR_CUDA, "stream"); break; }
      cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming);
    }
    if (rc) break;
    { const char *env = getenv("EPI_PASS2_RESTART"); h->restart = !(env && env[0] == '0'); }
    { const char *env = getenv("EPI_WINDOW_ONLY"); h->window_only = !(env && env[0] == '0'); }
    { const char *env = getenv("EPI_PASS1_CHUNK"); h->pass1_chunk = !(env && env[0] == '0'); }
    { const char *env = getenv("EPI_FORCE_R_FALLBACK"); h->force_need_r = (env && env[0] == '1'); }
    { const char *env = getenv("EPI_ODE_LANES"); h->ode_lanes = (env && (env[0] == '1' || env[0] == '2' || env[0] == '4')) ? env[0] - '0' : 0; }
    { const char *env = getenv("EPI_ODE_MINB"); h->ode_minb = (env && (env[0] == '4' || env[0] == '5')) ? env[0] - '0' : 0; }
    { const char *env = getenv("EPI_SPLIT_WAYS"); h->split_ways = (env && env[0] >= '1' && env[0] <= '8') ? env[0] - '0' : 0; }
    for (auto &ev : h->ev) cudaEventCreate(&ev);
    rc = fill_model(h, *desc, *data);
  } while (0);
  if (rc) {
    std::string msg = h->err;
    epi_destroy(h);
    g_last_error = msg;
    return rc;
  }
  *out = h;
  return EPI_OK;
}

static void free_workspace(Workspace &w) {
  cudaFree(w.theta); cudaFree(w.aos); cudaFree(w.hist1); cudaFree(w.hist2); cudaFree(w.ckpt); cudaFree(w.W); cudaFree(w.pmeta);
  cudaFree(w.offset); cudaFree(w.pad); cudaFree(w.status); cudaFree(w.need_r); cudaFree(w.logl); cudaFree(w.logp); cudaFree(w.terms);
  cudaFree(w.series);
  w = Workspace();
}

extern "C" void epi_destroy(epi_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  epi_sampler_free(h);
  free_workspace(h->ws);
  free_workspace(h->ws_traj);
  for (void *p : h->owned) cudaFree(p);
  for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->ev_ws) cudaEventDestroy(h->ev_ws);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (int k = 0; k < epi_handle::kMaxWays - 1; ++k) {
    if (h->split_stream[k]) { cudaStreamSynchronize(h->split_stream[k]); cudaStreamDestroy(h->split_stream[k]); }
    if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]);
  }
  for (int s = 0; s < 2; ++s) {
    cudaFree(h->stage[s]);
    if (h->ev_h2d[s]) { cudaEventDestroy(h->ev_h2d[s]); cudaEventDestroy(h->ev_free[s]); cudaEventDestroy(h->ev_done[s]); }
  }
  delete h;
}

extern "C" const char *epi_last_error(const epi_handle *h) { return h ? h->err.c_str() : g_last_error.c_str(); }
extern "C" int epi_n_params(const epi_handle *h) { return h ? h->M.d : 0; }
extern "C" int epi_n_series(const epi_handle *h) { return h ? (is_age(h->M.flavour) ? 8 : 3) : 0; }
extern "C" const char *epi_param_name(const epi_handle *h, int i) {
  if (!h || i < 0 || i >= h->M.d) return nullptr;
  return h->M.flavour == EPI_FLAVOUR_AGE2Q ? kAgeNames[i] : h->M.flavour == EPI_FLAVOUR_AGE2I ? kAge2iNames[i]
         : h->M.flavour == EPI_FLAVOUR_AGE2X ? kAge2xNames[i] : kSimpleNames[i];
}
extern "C" int epi_df_params(const epi_handle *h, double *mn, double *mx, double *init) {
  if (!h) return EPI_ERR_ARG;
  for (int i = 0; i < h->M.d; ++i) {
    if (mn) mn[i] = h->box_min[i];
    if (mx) mx[i] = h->box_max[i];
    if (init) init[i] = h->init[i];
  }
  return EPI_OK;
}
extern "C" int64_t epi_launch_count(const epi_handle *h) { return h ? h->launches : 0; }

extern "C" int64_t epi_fallback_count(epi_handle *h) {
  if (!h || !h->ws.need_r || h->ws.ld <= 0) return h ? 0 : -1;
  if (cudaSetDevice(h->device) != cudaSuccess) return -1;
  std::vector<int> host((size_t)h->ws.ld);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
  if (cudaMemcpy(host.data(), h->ws.need_r, sizeof(int) * host.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  int64_t c = 0;
  for (int v : host) c += v != 0;
  return c;
}

// ----------------------------------------------------------------- workspace
int epi_ensure_workspace(epi_handle *h, Workspace &w, const DevModel &M, int64_t n, int want_series) {
  // want_series: 0 = none, else the number of calculateModel series per chain (basic or basic + ext)
  const int NG = is_age(M.flavour) ? 2 : 1, NK = is_age(M.flavour) ? 6 : 2;
  if (n <= w.cap && M.P <= w.P && want_series <= w.ns) return EPI_OK;
  cudaStreamSynchronize(h->stream);
  if (h->ws_busy) cudaEventSynchronize(h->ev_ws);  // an evaluation on a caller's stream may still be using the buffers
  const int64_t cap = std::max<int64_t>(n, w.cap);
  const int P = std::max(M.P, w.P), P1 = std::max(M.P1, w.P1);
  const int NS = std::max(want_series, w.ns);
  const bool series = NS > 0;
  free_workspace(w);
  const int64_t tiles = (cap + kTile - 1) / kTile;
  const int64_t ld = tiles * kTile;
  CUDA_OK(h, cudaMalloc(&w.theta, sizeof(double) * (size_t)(M.d * ld)));
  CUDA_OK(h, cudaMalloc(&w.aos, sizeof(double) * (size_t)(M.d * ld)));
  // S histories: [chain][group][pitch], pitch = days rounded up to even (16-byte aligned rows for the bulk copies)
  CUDA_OK(h, cudaMalloc(&w.hist1, sizeof(double) * (size_t)(ld * NG * ((P1 + 1) & ~1))));
  CUDA_OK(h, cudaMalloc(&w.hist2, sizeof(double) * (size_t)(ld * NG * ((P + 7) & ~7))));  // four phase rows, hist_pitch2
  {
    const int NEQ = is_age(M.flavour) ? 10 : 4;
    CUDA_OK(h, cudaMalloc(&w.ckpt, sizeof(double) * (size_t)(ld * NEQ * (P1 < 64 ? P1 : 64))));
  }
  CUDA_OK(h, cudaMalloc(&w.W, sizeof(double) * (size_t)(ld * NG * pad_rows(M.flavour) * (is_age(M.flavour) ? 3 : 2))));  // [chain][group][delay][KG]
  CUDA_OK(h, cudaMalloc(&w.pmeta, sizeof(int2) * (size_t)(ld * NK)));
  CUDA_OK(h, cudaMalloc(&w.offset, sizeof(int) * (size_t)ld));
  CUDA_OK(h, cudaMalloc(&w.pad, sizeof(int) * (size_t)ld));
  CUDA_OK(h, cudaMalloc(&w.status, sizeof(int) * (size_t)ld));
  CUDA_OK(h, cudaMalloc(&w.need_r, sizeof(int) * (size_t)ld));
  CUDA_OK(h, cudaMemset(w.need_r, 0, sizeof(int) * (size_t)ld));
  CUDA_OK(h, cudaMalloc(&w.logl, sizeof(double) * (size_t)ld));
  CUDA_OK(h, cudaMalloc(&w.logp, sizeof(double) * (size_t)ld));
  CUDA_OK(h, cudaMalloc(&w.terms, sizeof(double) * (size_t)ld * 8));
  if (series) CUDA_OK(h, cudaMalloc(&w.series, sizeof(double) * (size_t)ld * NS * P));
  w.cap = cap; w.ld = ld; w.P = P; w.P1 = P1; w.ns = NS;
  return EPI_OK;
}

static EvalArgs make_args(epi_handle *h, Workspace &w, const DevModel &M, const double *d_theta, int64_t ld, int64_t n,
                          int64_t first, int model_space, double *d_logl, double *d_logp, int *d_status,
                          double *d_terms, double *d_series, cudaStream_t stream, int ns_total) {
  // the chain range [first, first + n) of a batch; `first` is a multiple of the 32-chain tile
  const int NG = is_age(M.flavour) ? 2 : 1, NK = is_age(M.flavour) ? 6 : 2;
  const int NS = ns_total > 0 ? ns_total : (is_age(M.flavour) ? 8 : 3);
  EvalArgs a;
  a.M = &M; a.PT = h->PT; a.n_prior = h->n_prior;
  a.theta = d_theta + first; a.ld = ld; a.n = n; a.model_space = model_space;
  a.hist1 = w.hist1 + first * NG * (int64_t)((M.P1 + 1) & ~1);
  a.hist2 = w.hist2 + first * NG * (int64_t)((M.P + 7) & ~7);
  a.W = w.W + first * NG * pad_rows(M.flavour) * (is_age(M.flavour) ? 3 : 2);
  // (`first` is a multiple of the 32-chain checkpoint tile)
  a.ckpt = h->restart ? w.ckpt + first * (is_age(M.flavour) ? 10 : 4) * (int64_t)(M.P1 < 64 ? M.P1 : 64) : nullptr;
  a.pmeta = w.pmeta + first * NK;
  a.offset = w.offset + first; a.pad = w.pad + first;
  a.status = (d_status ? d_status : w.status) + first;
  a.logl = d_logl ? d_logl + first : nullptr;
  a.logp = d_logp ? d_logp + first : nullptr;
  a.terms = d_terms ? d_terms + first * 8 : nullptr;
  a.series_out = d_series ? d_series + first * NS * M.P : nullptr;
  a.ns_total = NS;
  a.window_only = h->window_only ? 1 : 0;
  a.pass1_chunk = h->pass1_chunk ? 1 : 0;
  a.ode_lanes = h->ode_lanes;
  a.ode_minb = h->ode_minb;
  a.need_r = w.need_r + first;
  a.force_need_r = h->force_need_r ? 1 : 0;
  a.n_launches = &h->launches;
  a.stream = stream;
  a.ev = nullptr;
  return a;
}

int epi_launch_eval_ws(epi_handle *h, Workspace &w, const DevModel &M, const double *d_theta, int64_t ld, int64_t n,
                       int model_space, double *d_logl, double *d_logp, int *d_status, double *d_terms,
                       double *d_series, cudaStream_t stream, int ns_total) {
  // every whole-batch evaluation uses the SAME scratch (histories, checkpoints, profiles, offsets): one that runs on
  // another stream than its predecessor waits for the predecessor's last kernel
  if (!h->ev_ws) cudaEventCreateWithFlags(&h->ev_ws, cudaEventDisableTiming);
  if (h->ws_busy && h->ws_stream != stream) cudaStreamWaitEvent(stream, h->ev_ws, 0);
  EvalArgs a = make_args(h, w, M, d_theta, ld, n, 0, model_space, d_logl, d_logp, d_status, d_terms, d_series, stream, ns_total);
  a.ev = h->timing ? h->ev : nullptr;
  const cudaError_t e = launch_eval(a);
  if (e != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  cudaEventRecord(h->ev_ws, stream);
  h->ws_stream = stream;
  h->ws_busy = true;
  return EPI_OK;
}

// The chains [first, first + n) of a batch, with the workspace rows of that range: ranges of one batch that do not
// overlap may be in flight on different streams at the same time (the caller orders them against whatever produced
// theta and consumes the results).  `first` must be a multiple of 128.
int epi_launch_eval_range(epi_handle *h, Workspace &w, const DevModel &M, const double *d_theta, int64_t ld, int64_t first,
                          int64_t n, double *d_logl, double *d_logp, int *d_status, cudaStream_t stream) {
  EvalArgs a = make_args(h, w, M, d_theta, ld, n, first, 0, d_logl, d_logp, d_status, nullptr, nullptr, stream, 0);
  const cudaError_t e = launch_eval(a);
  if (e != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  return EPI_OK;
}

extern "C" int epi_eval_device_range(epi_handle *h, const double *d_theta_soa, int64_t ld, int64_t n_total, int64_t first,
                                     int64_t n, double *d_logl, double *d_logp, int32_t *d_status, void *stream) {
  if (!h || !d_theta_soa || n < 0 || first < 0 || first + n > n_total || ld < n_total || (first & 127))
    return epi_fail(h, EPI_ERR_ARG, "bad argument (first must be a multiple of 128, first + n <= n_total <= ld)");
  if (n == 0) return EPI_OK;
  CUDA_OK(h, cudaSetDevice(h->device));
  if (n_total > h->ws.cap) {
    // growing the workspace frees what other ranges may still be using: the caller sizes it first (epi_reserve)
    return epi_fail(h, EPI_ERR_STATE, "workspace holds %lld chains, the batch has %lld: call epi_reserve first",
                    (long long)h->ws.cap, (long long)n_total);
  }
  return epi_launch_eval_range(h, h->ws, h->M, d_theta_soa, ld, first, n, d_logl, d_logp, d_status,
                               stream ? (cudaStream_t)stream : h->stream);
}

extern "C" int epi_reserve(epi_handle *h, int64_t n) {
  if (!h || n < 1) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  CUDA_OK(h, cudaSetDevice(h->device));
  return epi_ensure_workspace(h, h->ws, h->M, n, false);
}

// ---------------------------------------------------------------------- eval
extern "C" int epi_eval_device(epi_handle *h, const double *d_theta_soa, int64_t n, double *d_logl, double *d_logp,
                               int32_t *d_status, void *stream) {
  if (!h || !d_theta_soa || n < 0) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  if (n == 0) return EPI_OK;
  CUDA_OK(h, cudaSetDevice(h->device));
  int rc = epi_ensure_workspace(h, h->ws, h->M, n, false);
  if (rc) return rc;
  return epi_launch_eval_ws(h, h->ws, h->M, d_theta_soa, n, n, 0, d_logl, d_logp, d_status, nullptr, nullptr,
                            stream ? (cudaStream_t)stream : h->stream);
}

static int stage_theta(epi_handle *h, Workspace &w, const double *theta, int64_t n, int layout, int d) {
  if (layout == EPI_LAYOUT_SOA) {
    CUDA_OK(h, cudaMemcpy2DAsync(w.theta, sizeof(double) * (size_t)w.ld, theta, sizeof(double) * (size_t)n,
                                 sizeof(double) * (size_t)n, d, cudaMemcpyHostToDevice, h->stream));
  } else if (layout == EPI_LAYOUT_AOS) {
    CUDA_OK(h, cudaMemcpyAsync(w.aos, theta, sizeof(double) * (size_t)(n * d), cudaMemcpyHostToDevice, h->stream));
    cudaError_t e = launch_aos_to_soa(w.aos, w.theta, n, d, w.ld, h->stream);
    h->launches += 1;
    if (e != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "transpose: %s", cudaGetErrorString(e));
  } else {
    return epi_fail(h, EPI_ERR_ARG, "unknown layout %d", layout);
  }
  return EPI_OK;
}

extern "C" int epi_eval(epi_handle *h, const double *theta, int64_t n, int layout, double *logl, double *logp,
                        int32_t *status) {
  if (!h || !theta || n < 0) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  if (n == 0) return EPI_OK;
  CUDA_OK(h, cudaSetDevice(h->device));
  int rc = epi_ensure_workspace(h, h->ws, h->M, n, false);
  if (rc) return rc;
  Workspace &w = h->ws;
  if ((rc = stage_theta(h, w, theta, n, layout, h->M.d))) return rc;
  if ((rc = epi_launch_eval_ws(h, w, h->M, w.theta, w.ld, n, 0, w.logl, w.logp, nullptr, nullptr, nullptr, h->stream))) return rc;
  if (logl) CUDA_OK(h, cudaMemcpyAsync(logl, w.logl, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  if (logp) CUDA_OK(h, cudaMemcpyAsync(logp, w.logp, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  if (status) CUDA_OK(h, cudaMemcpyAsync(status, w.status, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return EPI_OK;
}

extern "C" int epi_eval_submit(epi_handle *h, int slot, const double *theta, int64_t n, int layout, double *logl,
                               double *logp, int32_t *status) {
  if (!h || !theta || n < 1 || slot < 0 || slot > 1) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  if (layout != EPI_LAYOUT_AOS && layout != EPI_LAYOUT_SOA) return epi_fail(h, EPI_ERR_ARG, "unknown layout %d", layout);
  if (h->slot_pending[slot]) return epi_fail(h, EPI_ERR_STATE, "slot %d has a submission that was not waited for", slot);
  CUDA_OK(h, cudaSetDevice(h->device));
  if (n > h->ws.cap) {  // growing the workspace frees buffers the other slot's kernels may still use
    for (int s = 0; s < 2; ++s)
      if (h->slot_pending[s]) CUDA_OK(h, cudaEventSynchronize(h->ev_done[s]));
  }
  int rc = epi_ensure_workspace(h, h->ws, h->M, n, false);
  if (rc) return rc;
  Workspace &w = h->ws;
  const int d = h->M.d;
  if (!h->ev_h2d[slot]) {
    cudaEventCreateWithFlags(&h->ev_h2d[slot], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_free[slot], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_done[slot], cudaEventDisableTiming);
  }
  if (n * d > h->stage_cap[slot]) {
    if (h->slot_used[slot]) CUDA_OK(h, cudaEventSynchronize(h->ev_free[slot]));
    cudaFree(h->stage[slot]);
    h->stage[slot] = nullptr;
    h->stage_cap[slot] = 0;
    CUDA_OK(h, cudaMalloc(&h->stage[slot], sizeof(double) * (size_t)(n * d)));
    h->stage_cap[slot] = n * d;
  }
  cudaStream_t cs = h->stream2, ks = h->stream;
  // copy stream: the staging buffer is free once the previous user's re-layout has read it
  if (h->slot_used[slot]) CUDA_OK(h, cudaStreamWaitEvent(cs, h->ev_free[slot], 0));
  CUDA_OK(h, cudaMemcpyAsync(h->stage[slot], theta, sizeof(double) * (size_t)(n * d), cudaMemcpyHostToDevice, cs));
  CUDA_OK(h, cudaEventRecord(h->ev_h2d[slot], cs));
  // compute stream: re-layout into the SoA workspace (ordered after the previous batch's kernels), evaluate, return
  CUDA_OK(h, cudaStreamWaitEvent(ks, h->ev_h2d[slot], 0));
  if (layout == EPI_LAYOUT_AOS) {
    cudaError_t e = launch_aos_to_soa(h->stage[slot], w.theta, n, d, w.ld, ks);
    h->launches += 1;
    if (e != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "transpose: %s", cudaGetErrorString(e));
  } else {
    CUDA_OK(h, cudaMemcpy2DAsync(w.theta, sizeof(double) * (size_t)w.ld, h->stage[slot], sizeof(double) * (size_t)n,
                                 sizeof(double) * (size_t)n, d, cudaMemcpyDeviceToDevice, ks));
  }
  CUDA_OK(h, cudaEventRecord(h->ev_free[slot], ks));
  h->slot_used[slot] = true;
  if ((rc = epi_launch_eval_ws(h, w, h->M, w.theta, w.ld, n, 0, w.logl, w.logp, nullptr, nullptr, nullptr, ks))) return rc;
  if (logl) CUDA_OK(h, cudaMemcpyAsync(logl, w.logl, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ks));
  if (logp) CUDA_OK(h, cudaMemcpyAsync(logp, w.logp, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ks));
  if (status) CUDA_OK(h, cudaMemcpyAsync(status, w.status, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ks));
  CUDA_OK(h, cudaEventRecord(h->ev_done[slot], ks));
  h->slot_pending[slot] = true;
  return EPI_OK;
}

extern "C" int epi_eval_wait(epi_handle *h, int slot) {
  if (!h || slot < 0 || slot > 1) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  if (!h->slot_pending[slot]) return epi_fail(h, EPI_ERR_STATE, "nothing submitted on slot %d", slot);
  CUDA_OK(h, cudaSetDevice(h->device));
  CUDA_OK(h, cudaEventSynchronize(h->ev_done[slot]));
  h->slot_pending[slot] = false;
  return EPI_OK;
}

extern "C" int epi_eval_detail(epi_handle *h, const double *theta, int64_t n, int layout, int32_t *offset,
                               int32_t *padding, double *terms) {
  if (!h || !theta || n < 0) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  if (n == 0) return EPI_OK;
  CUDA_OK(h, cudaSetDevice(h->device));
  int rc = epi_ensure_workspace(h, h->ws, h->M, n, false);
  if (rc) return rc;
  Workspace &w = h->ws;
  if ((rc = stage_theta(h, w, theta, n, layout, h->M.d))) return rc;
  if ((rc = epi_launch_eval_ws(h, w, h->M, w.theta, w.ld, n, 0, w.logl, w.logp, nullptr, w.terms, nullptr, h->stream))) return rc;
  if (offset) CUDA_OK(h, cudaMemcpyAsync(offset, w.offset, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  if (padding) CUDA_OK(h, cudaMemcpyAsync(padding, w.pad, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  if (terms) CUDA_OK(h, cudaMemcpyAsync(terms, w.terms, sizeof(double) * (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return EPI_OK;
}

static int n_series_of(int flavour, bool ext) { return is_age(flavour) ? (ext ? 18 : 8) : (ext ? 8 : 3); }
extern "C" int epi_n_series_ext(const epi_handle *h) { return h ? n_series_of(h->M.flavour, true) : 0; }

static int trajectories_impl(epi_handle *h, const double *theta, int64_t n, int layout, int32_t period, double *out,
                             int32_t *offset, int32_t *padding, bool ext) {
  if (!h || !theta || !out || n < 0) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  if (n == 0) return EPI_OK;
  const bool age = is_age(h->M.flavour), fixed_p1 = fixed_pass1(h->M.flavour);  // 2q / 2x: period1 = 60
  if (period < (fixed_p1 ? 60 : 2) || period > (age ? 2048 : 4096)) return epi_fail(h, EPI_ERR_ARG, "period out of range");
  CUDA_OK(h, cudaSetDevice(h->device));
  DevModel M = h->M;  // calculateModel(params, period): same data, another horizon
  M.P = period;
  if (!fixed_p1) M.P1 = period;
  // age-2x: the histories between the kernels hold the infection incidence, S comes from the ext instance of the
  // ODE kernel — the ext pass always runs and the basic call copies its first eight series out
  const int NS = n_series_of(M.flavour, ext), NSrun = M.flavour == EPI_FLAVOUR_AGE2X ? n_series_of(M.flavour, true) : NS;
  int rc = epi_ensure_workspace(h, h->ws_traj, M, n, NSrun);
  if (rc) return rc;
  Workspace &w = h->ws_traj;
  if ((rc = stage_theta(h, w, theta, n, layout, M.d))) return rc;
  if ((rc = epi_launch_eval_ws(h, w, M, w.theta, w.ld, n, 1, nullptr, nullptr, nullptr, nullptr, w.series, h->stream, NSrun))) return rc;
  if (NSrun == NS)
    CUDA_OK(h, cudaMemcpyAsync(out, w.series, sizeof(double) * (size_t)n * NS * period, cudaMemcpyDeviceToHost, h->stream));
  else
    CUDA_OK(h, cudaMemcpy2DAsync(out, sizeof(double) * (size_t)NS * period, w.series, sizeof(double) * (size_t)NSrun * period,
                                 sizeof(double) * (size_t)NS * period, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  if (offset) CUDA_OK(h, cudaMemcpyAsync(offset, w.offset, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  if (padding) CUDA_OK(h, cudaMemcpyAsync(padding, w.pad, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return EPI_OK;
}

extern "C" int epi_trajectories(epi_handle *h, const double *theta, int64_t n, int layout, int32_t period, double *out,
                                int32_t *offset, int32_t *padding) {
  return trajectories_impl(h, theta, n, layout, period, out, offset, padding, false);
}
extern "C" int epi_trajectories_ext(epi_handle *h, const double *theta, int64_t n, int layout, int32_t period, double *out,
                                    int32_t *offset, int32_t *padding) {
  return trajectories_impl(h, theta, n, layout, period, out, offset, padding, true);
}

extern "C" int epi_quantiles(epi_handle *h, const double *theta, int64_t n, int layout, int32_t period, int32_t startoffset,
                             int32_t series_a, int32_t series_b, const double *probs, int32_t n_probs, double *out) {
  if (!h || !theta || !probs || !out || n < 1 || n_probs < 1 || period < 1) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  // up to 8,192 draws a day's values are sorted in shared memory, beyond that in a global scratch buffer of
  // period x 2^ceil(log2 n) doubles (capped at 4 GiB)
  int64_t npow2 = 1;
  while (npow2 < n) npow2 <<= 1;
  const bool big = n > 8192;
  if (big && (double)npow2 * period * 8.0 > 4294967296.0)
    return epi_fail(h, EPI_ERR_ARG, "epi_quantiles: %lld draws x %d days need more than 4 GiB of sort scratch", (long long)n, period);
  const bool age = is_age(h->M.flavour), fixed_p1 = fixed_pass1(h->M.flavour);
  const int NSb = n_series_of(h->M.flavour, false), NSx = n_series_of(h->M.flavour, true);
  if (series_a < 0 || series_a >= NSx || series_b >= NSx) return epi_fail(h, EPI_ERR_ARG, "series index out of range");
  const int NS = (series_a >= NSb || series_b >= NSb || h->M.flavour == EPI_FLAVOUR_AGE2X) ? NSx : NSb;  // an ext series asked for (age-2x: always): run the ext pass
  const int simperiod = 3 * period;  // libMCMC.R:153
  if (simperiod < (fixed_p1 ? 60 : 2) || simperiod > (age ? 2048 : 4096)) return epi_fail(h, EPI_ERR_ARG, "period out of range");
  CUDA_OK(h, cudaSetDevice(h->device));
  DevModel M = h->M;
  M.P = simperiod;
  if (!fixed_p1) M.P1 = simperiod;
  int rc = epi_ensure_workspace(h, h->ws_traj, M, n, NS);
  if (rc) return rc;
  Workspace &w = h->ws_traj;
  if ((rc = stage_theta(h, w, theta, n, layout, M.d))) return rc;
  if ((rc = epi_launch_eval_ws(h, w, M, w.theta, w.ld, n, 1, nullptr, nullptr, nullptr, nullptr, w.series, h->stream, NS))) return rc;
  double *d_probs = nullptr, *d_out = nullptr;
  CUDA_OK(h, cudaMalloc(&d_probs, sizeof(double) * (size_t)n_probs));
  CUDA_OK(h, cudaMalloc(&d_out, sizeof(double) * (size_t)period * n_probs));
  CUDA_OK(h, cudaMemcpyAsync(d_probs, probs, sizeof(double) * (size_t)n_probs, cudaMemcpyHostToDevice, h->stream));
  // values before padding+1: the pre-filled state (model-age-groups2q.R:118-134)
  // (ext: y.E / o.E are first assigned on (padding+1):..., so 1..padding is NA there, :216-223;
  //  simple E is pre-filled with Initial, model-simple-ode-drj-cmp.R:66)
  auto prefill = [&](int sidx) -> double {
    if (sidx < 0) return 0.0;
    if (age) return sidx == 0 ? M.Ny - 1.0 : sidx == 1 ? M.No : (sidx == 8 || sidx == 11) ? NAN : 0.0;
    return sidx == 0 ? M.N - 1.0 : sidx == 3 ? 1.0 : 0.0;
  };
  double *d_scratch = nullptr;
  if (big) CUDA_OK(h, cudaMalloc(&d_scratch, sizeof(double) * (size_t)npow2 * period));
  cudaError_t e = launch_quantiles(w.series, w.offset, w.pad, n, NS, simperiod, period, startoffset, series_a, series_b,
                                   prefill(series_a), prefill(series_b), d_probs, n_probs, d_out, d_scratch, h->stream);
  h->launches += 1;
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)period * n_probs, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_probs);
  cudaFree(d_out);
  cudaFree(d_scratch);
  if (e != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "epi_quantiles: %s", cudaGetErrorString(e));
  return EPI_OK;
}

extern "C" int epi_marginalize(epi_handle *h, const double *theta, int64_t n, int layout, int32_t period, int32_t startoffset,
                               int32_t series_a, int32_t series_b, int32_t kind, int32_t j1, int32_t j2, double *out) {
  if (!h || !theta || !out || n < 1 || period < 1) return epi_fail(h, EPI_ERR_ARG, "bad argument");
  const bool age = is_age(h->M.flavour), fixed_p1 = fixed_pass1(h->M.flavour);
  const int NSb = n_series_of(h->M.flavour, false), NSx = n_series_of(h->M.flavour, true);
  if (series_a < 0 || series_a >= NSx || series_b >= NSx) return epi_fail(h, EPI_ERR_ARG, "series index out of range");
  const int simperiod = 3 * period;  // libMCMC.R:176
  if (kind < 0 || kind > 3 || j1 < 1 || j2 < j1 || j2 > simperiod) return epi_fail(h, EPI_ERR_ARG, "bad marginal (kind 0..3, 1 <= j1 <= j2 <= 3 * period)");
  if (simperiod < (fixed_p1 ? 60 : 2) || simperiod > (age ? 2048 : 4096)) return epi_fail(h, EPI_ERR_ARG, "period out of range");
  const int NS = (series_a >= NSb || series_b >= NSb || h->M.flavour == EPI_FLAVOUR_AGE2X) ? NSx : NSb;
  CUDA_OK(h, cudaSetDevice(h->device));
  DevModel M = h->M;
  M.P = simperiod;
  if (!fixed_p1) M.P1 = simperiod;
  int rc = epi_ensure_workspace(h, h->ws_traj, M, n, NS);
  if (rc) return rc;
  Workspace &w = h->ws_traj;
  if ((rc = stage_theta(h, w, theta, n, layout, M.d))) return rc;
  if ((rc = epi_launch_eval_ws(h, w, M, w.theta, w.ld, n, 1, nullptr, nullptr, nullptr, nullptr, w.series, h->stream, NS))) return rc;
  auto prefill = [&](int sidx) -> double {  // values before padding + 1, as in epi_quantiles
    if (sidx < 0) return 0.0;
    if (age) return sidx == 0 ? M.Ny - 1.0 : sidx == 1 ? M.No : (sidx == 8 || sidx == 11) ? NAN : 0.0;
    return sidx == 0 ? M.N - 1.0 : sidx == 3 ? 1.0 : 0.0;
  };
  double *d_out = nullptr;
  CUDA_OK(h, cudaMalloc(&d_out, sizeof(double) * (size_t)n));
  cudaError_t e = launch_marginal(w.series, w.offset, w.pad, n, NS, simperiod, startoffset, series_a, series_b,
                                  prefill(series_a), prefill(series_b), kind, j1, j2, d_out, h->stream);
  h->launches += 1;
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_out);
  if (e != cudaSuccess) return epi_fail(h, EPI_ERR_CUDA, "epi_marginalize: %s", cudaGetErrorString(e));
  return EPI_OK;
}

// --------------------------------------------------------------- measurement
extern "C" int epi_set_timing(epi_handle *h, int enable) {
  if (!h) return EPI_ERR_ARG;
  h->timing = enable != 0;
  return EPI_OK;
}

extern "C" int epi_last_kernel_ms(epi_handle *h, float ms[4]) {
  if (!h || !ms) return EPI_ERR_ARG;
  if (!h->timing) return epi_fail(h, EPI_ERR_STATE, "timing not enabled");
  CUDA_OK(h, cudaSetDevice(h->device));
  CUDA_OK(h, cudaEventSynchronize(h->ev[4]));
  for (int i = 0; i < 4; ++i) CUDA_OK(h, cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
  return EPI_OK;
}

extern "C" int epi_fp64_peak(epi_handle *h, double *tflops) {
  if (!h || !tflops) return EPI_ERR_ARG;
  CUDA_OK(h, cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CUDA_OK(h, cudaGetDeviceProperties(&prop, h->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double *buf = nullptr;
  CUDA_OK(h, cudaMalloc(&buf, sizeof(double) * (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0, h->stream);
    launch_dfma(buf, blocks, threads, iters, h->stream);
    cudaEventRecord(e1, h->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0 && ms > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    h->launches += 1;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *tflops = best;
  return EPI_OK;
}
