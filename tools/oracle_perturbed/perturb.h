/* tools/oracle_perturbed/perturb.h -- DEBUG ONLY (never part of the oracle proper): force-included into a private build
 * of oracle/oracle.c so that the outputs of the cubic solver's libm calls can be moved, to study which verdicts depend on
 * their last bits (tools/fragile_cpu.py). orc_pert_mode: 0 = relative factor on pow / absolute offset on cos; 1 = also
 * pow(x, 1./3) answered by cbrt(x) (what the device library evaluates). */
#include <math.h>
extern double orc_pert_a, orc_pert_d;
extern int orc_pert_mode, orc_trace_on;
#include <stdio.h>
#define ORC_TRACE 1
static inline double orc_pow_(double x, double y) { return (orc_pert_mode == 1 ? cbrt(x) : pow(x, y)) * orc_pert_a; }
static inline double orc_cos_(double x) { return cos(x) + orc_pert_d; }
#define pow orc_pow_
#define cos orc_cos_
