/* mithra_oracle.h -- CPU restatement of the MITHRA time-march (TEST INFRASTRUCTURE, never shipped or measured
 * as the product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it).
 *
 * Plain C99, scalar, single thread, compiled with -ffp-contract=off so that it rounds like the reference's
 * x86-64 build.  It is pinned against the unmodified reference (oracle/_ref/ref_dump, built from
 * /root/reference/src) by tests/test_oracle_vs_reference.py and against the committed fixtures in tests/golden/.
 *
 * It takes the same parameter block as the CUDA library (include/mithra_gpu.h) and keeps the state in the
 * reference's own layouts: potentials double[nodes][3], node m = N1*N0*k + N1*i + j; particles double[n][11].
 */
#ifndef MITHRA_ORACLE_H_
#define MITHRA_ORACLE_H_

#include <stddef.h>
#include "../include/mithra_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct Oracle Oracle;

Oracle* oracle_create  (const MithraGpuParams* p);
void    oracle_destroy (Oracle* o);

/* raw access to the state arrays (owned by the oracle) */
double* oracle_anp1 (Oracle* o);   /* new potential after field_update, current J after current_update (fdtd.cpp:27,244) */
double* oracle_an   (Oracle* o);
double* oracle_anm1 (Oracle* o);
double* oracle_fnp1 (Oracle* o);
double* oracle_fn   (Oracle* o);
double* oracle_fnm1 (Oracle* o);
float*  oracle_en   (Oracle* o);
float*  oracle_bn   (Oracle* o);
unsigned char* oracle_pic (Oracle* o);

void    oracle_set_particles (Oracle* o, const double* aos11, size_t n);
size_t  oracle_num_particles (Oracle* o);
double* oracle_particles     (Oracle* o);
void    oracle_set_time (Oracle* o, double time, double time_bunch, unsigned int n_time);
double  oracle_time (Oracle* o);
double  oracle_time_bunch (Oracle* o);

/* the reference methods of the time march */
void oracle_field_update   (Oracle* o);          /* FdTd::fieldUpdate / FdTdSC::fieldUpdate                 */
void oracle_field_evaluate (Oracle* o, long m);  /* FdTd::fieldEvaluate                                    */
void oracle_field_sample (Oracle* o, const double* pos3, double* out9);   /* FdTd::fieldSample fdtd.cpp:851-913 (one point) */
void oracle_bunch_update   (Oracle* o);          /* rnm = rnp, then nUpdateBunch x Solver::bunchUpdate     */
void oracle_screen_profile (Oracle* o);          /* Solver::screenProfile                                  */
void oracle_power_sample   (Oracle* o);          /* Solver::powerSample                                    */
void oracle_field_shift    (Oracle* o);
void oracle_current_reset  (Oracle* o);
void oracle_current_update (Oracle* o);
void oracle_advance_time   (Oracle* o);
void oracle_step           (Oracle* o, int nsteps);

/* cell index of every particle as the push computes it (solver.cpp:1464-1469), -1 when it gathers nothing */
void oracle_push_cells    (Oracle* o, long* m_out);
/* cell indices (ip,jp,kp,im,jm,km) as the deposit computes them (fdtd.cpp:70-77) */
void oracle_deposit_cells (Oracle* o, int* ijk6_out);

/* outputs */
void          oracle_power_visualize (Oracle* o);              /* Solver::powerVisualize radiation.cpp:324-391 */
const double* oracle_power_map    (Oracle* o);                 /* pL[i*N1 + j], 0 when power_map is disabled  */
size_t        oracle_power_rows   (Oracle* o);                 /* rows of N*Nl doubles recorded so far      */
const double* oracle_power_data   (Oracle* o);
size_t        oracle_screen_count (Oracle* o, int screen);
const double* oracle_screen_data  (Oracle* o, int screen);    /* records of 6 doubles                      */

/* seed potential at a node (Seed::fields) and the TF box initial condition (solver.cpp:828-839) */
void oracle_seed_fields  (Oracle* o, double x, double y, double z, double time, double a[3]);
void oracle_seed_initial (Oracle* o);

#ifdef __cplusplus
}
#endif

#endif
