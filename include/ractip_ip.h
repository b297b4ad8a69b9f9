/*
 * ractip_ip.h -- C ABI of the host-side consumer of the probability stage: RactIP's integer
 * programme (model build + decoding) and the free-energy evaluation of the predicted joint
 * structure.  SURVEY.md section 8(f) rows 1 and 2; everything here is host code (no GPU).
 *
 * The model is built EXACTLY as RactIP::solve builds it (reference src/ractip.cpp:552-1225):
 * same binary variables in the same creation order (= column order), same constraint rows in the
 * same order, same float arithmetic for the objective weights.  The library does not contain a
 * MIP solver (the reference links GLPK / Gurobi / CPLEX / SCIP / HiGHS behind src/ip.cpp:49-622;
 * none of them is in this image): rp_ip_export hands the model out as plain arrays, the caller
 * solves it (ractip_b200/ip.py uses scipy.optimize.milp = HiGHS), and rp_ip_decode turns the
 * column values into the two dot-bracket strings of src/ractip.cpp:1227-1316.
 */
#ifndef RACTIP_IP_H
#define RACTIP_IP_H

#include <stddef.h>
#include <stdint.h>

#include "ractip_prob.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the members of class RactIP that shape the model (src/ractip.cpp:159-183, defaults of
 * src/cmdline.c:151-186 as mapped in RactIP::parse_options, src/ractip.cpp:1474-1498) */
typedef struct rp_ip_opts {
  float alpha;         /* -a, weight for hybridization, default 0.7            */
  float beta;          /* -b, weight for accessibility, default 0.0            */
  float th_ss;         /* -t, default 0.5                                      */
  float th_hy;         /* -u, default 0.1                                      */
  float th_ac;         /* -s, default 0.003                                    */
  int max_w, min_w;    /* --max-w 15, --min-w 5                                */
  int acc_max;         /* --acc-max                                            */
  int acc_max_ss;      /* --acc-max-ss                                         */
  int acc_num;         /* --acc-num, default 1                                 */
  int in_pk;           /* !--no-pk, default 1                                  */
  int stacking;        /* !--allow-isolated, default 1                         */
} rp_ip_opts;

void rp_ip_opts_default(rp_ip_opts* o);

/* row kinds of IP::BoundType (src/ip.h): only the bound a kind names is used */
enum { RP_IP_FR = 0, RP_IP_LO = 1, RP_IP_UP = 2, RP_IP_DB = 3, RP_IP_FX = 4 };

typedef struct rp_ip_model rp_ip_model;

/* Joint model of RactIP::solve (src/ractip.cpp:552-1225) from the DENSE matrices in the
 * reference's layouts (the output sections of rp_run_dense for one pair):
 *   bp1/bp2 : bp[offset[i]+j] upper triangles, (L+1)(L+2)/2 floats   (:314-317)
 *   up1/up2 : L x max(1,max_w) floats row-major                        (:370-375)
 *   hp      : (n1+1) x (n2+1) floats, 1-based                          (:404-405,451-453)
 * Structure constraints (-c / --force-constraint, :655-713,1170-1222) are not modelled. */
int rp_ip_build(const rp_ip_opts* o, int n1, int n2, const float* bp1, const float* bp2,
                const float* up1, const float* up2, const float* hp, rp_ip_model** out);

/* The same model from the thresholded variable lists the sparse path of the probability stage
 * emits (rp_run_sparse / rp_batch_fetch_sparse: x, y, z in creation order, v/w as (i, j, p) with
 * j the 0-based window-length index of src/ractip.cpp:622).  Column for column identical to
 * rp_ip_build on the dense matrices the lists were cut from. */
int rp_ip_build_sparse(const rp_ip_opts* o, int n1, int n2, const rp_rec* x, int nx, const rp_rec* y, int ny,
                       const rp_rec* z, int nz, const rp_rec* v, int nv, const rp_rec* w, int nw,
                       rp_ip_model** out);

/* Single-sequence model of RactIP::solve_ss (src/ractip.cpp:1366-1465).  usable[i] != 0 marks
 * the bases that may pair (NULL: all), as the --acc-max-ss branch passes them (:1263-1271). */
int rp_ip_build_ss(const rp_ip_opts* o, int n, const float* bp, const unsigned char* usable,
                   rp_ip_model** out);

int rp_ip_dims(const rp_ip_model* m, int* n_cols, int* n_rows, int* n_nonzeros);
/* obj[n_cols] (maximise); row_kind/row_lo/row_hi[n_rows]; triplets ia (row), ja (col), ar of
 * n_nonzeros entries, 0-based, in IP::add_constraint call order.  All columns are binary. */
int rp_ip_export(const rp_ip_model* m, double* obj, int* row_kind, double* row_lo, double* row_hi,
                 int* ia, int* ja, double* ar);
/* Column values -> dot-bracket strings (src/ractip.cpp:1227-1250,1286-1295): r1 (n1+1 bytes),
 * r2 (n2+1 bytes, ignored for a solve_ss model), NUL-terminated.  For a joint model with
 * --acc-max the chosen accessible regions are returned through used1/used2 (n1 / n2 bytes,
 * 1 = inside a chosen region; may be NULL). */
int rp_ip_decode(const rp_ip_model* m, const double* col_values, char* r1, char* r2,
                 unsigned char* used1, unsigned char* used2);
void rp_ip_free(rp_ip_model* m);

/* Free energy (kcal/mol) of `structure` on `seq` under the INTEGER energy tables of *model with
 * ViennaRNA's energy_of_structure semantics at dangles = 2 (d2: every stem of an exterior or
 * multi-loop takes both neighbouring bases when they exist), as RactIP calls it at
 * src/ractip.cpp:1254,1299,1457 (cut_point = -1) and, through energy_of_duplex, at :1528-1559
 * (cut_point = |s1|+1 on the concatenation, '[' ']' turned into '(' ')').  Characters other than
 * '(' and ')' count as unpaired.  cut_point <= 0: single strand.  Returns RP_OK or an error;
 * the energy goes to *energy (float, as the reference holds it). */
int rp_energy_of_structure(const rp_model* model, const char* seq, const char* structure, int n,
                           int cut_point, float* energy);
/* RactIP::energy_of_duplex (src/ractip.cpp:1528-1559). */
int rp_energy_of_duplex(const rp_model* model, const char* s1, int n1, const char* s2, int n2,
                        const char* r1, const char* r2, float* energy);

#ifdef __cplusplus
}
#endif
#endif /* RACTIP_IP_H */
