/*
 * oracle/ossg_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * C interface of the CPU restatement (see ossg_oracle.c for the parity status).
 */
#ifndef OSSG_ORACLE_H
#define OSSG_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int Nx, Ny;                      /* size[1:2]  (src/tripolar_grid.jl:78)  */
    int Hx, Hy;                      /* halo[1:2]  (src/tripolar_grid.jl:79)  */
    double southernmost_latitude;    /* default -80 */
    double north_poles_latitude;     /* default  55 */
    double first_pole_longitude;     /* default  70 */
    double radius;                   /* default R_Earth = 6371.0e3 */
} ossg_oracle_params;

int  ossg_oracle_check_params(const ossg_oracle_params *p);
void ossg_oracle_coordinate_vectors(const ossg_oracle_params *p, double *lamF, double *lamC,
                                    double *phiF, double *phiC);
int  ossg_oracle_coords(const ossg_oracle_params *p, double *const coords[8]);
int  ossg_oracle_metrics(const ossg_oracle_params *p, const double *const coords[8],
                         double *const metrics[12]);
int  ossg_oracle_build_f64(const ossg_oracle_params *p, double *const out[20]);
int  ossg_oracle_build_f32(const ossg_oracle_params *p, float *const out[20]);

void ossg_oracle_zipper_f64(double *c, int Nx, int Ny, int Nz, int Hx, int Hy, int Hz,
                            int locx, int locy, double sign, int k0, int nk);
void ossg_oracle_zipper_f32(float *c, int Nx, int Ny, int Nz, int Hx, int Hy, int Hz,
                            int locx, int locy, double sign, int k0, int nk);
void ossg_oracle_periodic_x_f64(double *c, int Nx, int Ny, int Nz, int Hx, int Hy, int Hz);
void ossg_oracle_periodic_x_f32(float *c, int Nx, int Ny, int Nz, int Hx, int Hy, int Hz);

int  ossg_oracle_uses_libm(void);
int  ossg_oracle_max_threads(void);
void ossg_oracle_set_threads(int n);
void ossg_oracle_set_lambda_f32(int on);   /* 1: 1-D longitudes pre-rounded to Float32 (FT = Float32 reading) */
int  ossg_oracle_math_eval(int fn, const double *x, double *y, size_t n);
void ossg_oracle_cartesian_nodes(const double *lam, const double *phi, size_t n, double radius,
                                 double *x, double *y, double *z);

#ifdef __cplusplus
}
#endif
#endif
