/* epi_oracle.h — declarations of the CPU restatement (TEST INFRASTRUCTURE, see
 * epi_oracle.c).  PARITY UNPINNED vs executed reference output (no R here). */
#ifndef EPI_ORACLE_H
#define EPI_ORACLE_H
#include <stdint.h>
#include "../include/epimcmc.h"
#ifdef __cplusplus
extern "C" {
#endif
double oracle_pgamma(double q, double shape, double scale);
double oracle_pnorm(double q, double mean, double sd);
double oracle_dnorm_log(double x, double mean, double sd);
double oracle_dlnorm_log(double x, double meanlog, double sdlog);
double oracle_dnbinom_mu_log(double x, double size, double mu);
double oracle_age_gamma(void);
void oracle_derivs_age(const double *parms45, double t, const double *y, double *ydot);
void oracle_derivs_agef(const double *parms37, double t, const double *y, double *ydot);
void oracle_derivs_ager(const double *parms60, double t, const double *y, double *ydot, double *yout9);
int oracle_n_params(int flavour);
int oracle_n_series(int flavour);
int oracle_df_params(const epi_model_desc *desc, const epi_data *D, double *mn, double *mx, double *init);
void oracle_eval_one(const epi_model_desc *desc, const epi_data *D, const double *theta, double *logl,
                     double *logp, int32_t *status, int32_t *offset, int32_t *padding, double *terms8);
int oracle_eval(const epi_model_desc *desc, const epi_data *D, const double *theta, int64_t n,
                int n_threads, double *logl, double *logp, int32_t *status, int32_t *offset,
                int32_t *padding, double *terms8);
int oracle_trajectory(const epi_model_desc *desc, const epi_data *D, const double *theta_model,
                      int period, double *out, double *states, double *ext, int32_t *offset, int32_t *padding);
void oracle_yout_age(const double *parms45, double t, const double *y, double *yout);
void oracle_yout_agef(const double *parms37, double t, const double *y, double *yout);
void oracle_set_jitter(int ulps, unsigned seed);
/* tests only: evaluate the ODE scheme in the CUDA kernels' operation order (explicit fma), see epi_oracle.c */
void oracle_set_mirror(int on);
int oracle_profile(int kind, double mean, double sd, double *values, int *kbegin, int *kend);
#ifdef __cplusplus
}
#endif
#endif
