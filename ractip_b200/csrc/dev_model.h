// dev_model.h -- fp64 Boltzmann tables as the kernels consume them, and the
// host routine that builds them from the integer model.
//
// Replaces ViennaRNA's get_scaled_pf_parameters()/scale_parameters() as used
// inside pf_fold / co_pf_fold (reference call sites src/ractip.cpp:356,444) and
// by src/pf_duplex.c:78-81.  Plain-old-data: uploaded to HBM once per context.
#ifndef RP_DEV_MODEL_H
#define RP_DEV_MODEL_H

#include <stdint.h>

#include "ractip_prob.h"

namespace rp {

constexpr int TURN = RP_TURN;
constexpr int MAXLOOP = RP_MAXLOOP;

// One term of the factorised interior-loop sum: inner pair sits `dd` diagonals
// below the closing pair and `po` positions to the right of it
// (dd = u1+u2+2, po = u1+1); g = loop weight that depends on (u1,u2) only,
// already multiplied by scale[u1+u2+2].
struct Tap {
  int16_t dd, po;
  int32_t u2;
  double g;
};

enum { TAP_GENERIC = 0, TAP_1N = 1, TAP_BULGE = 2, TAP_CLASSES = 3 };
constexpr int MAX_TAPS = 384;

struct DevModel {
  double pf_scale, scale1, mlb1;  // scale1 = 1/pf_scale, mlb1 = expMLbase*scale1
  double kT, lxc;
  double expMLclosing, expMLintern, expTermAU, expMLbase;
  double scale_small[40];  // scale1^k
  double exphairpin[31], expbulge[31], expinternal[31], expninio[MAXLOOP + 1];
  double expstack[8][8];
  double mmI[8][5][5], mmH[8][5][5], mmM[8][5][5], mmExt[8][5][5], mm1n[8][5][5], mm23[8][5][5];
  double dangle5[8][5], dangle3[8][5];
  double int11[8][8][5][5];
  double int21[8][8][5][5][5];
  double int22[8][8][5][5][5][5];
  // special hairpins: k-mers as base-8 codes of the 1..4 encoding
  int n_tetra, n_tri, n_hex;
  int tetra_code[200], tri_code[40], hex_code[200];
  double exptetra[200], exptri[40], exphex[200];
  int special_hp;
  // factorised interior loops, per class, sorted by dd ascending
  int ntaps[TAP_CLASSES];
  int tap_prefix[TAP_CLASSES][MAXLOOP + 4];  // #taps with dd <= x
  Tap taps[TAP_CLASSES][MAX_TAPS];
  // integer view for pf_duplex (scale_parameters semantics at 37 C)
  int i_dangle5[8][5], i_dangle3[8][5], i_mmExt[8][5][5];
  int i_stack[8][8], i_bulge[31], i_internal[31], i_mmI[8][5][5], i_mm1n[8][5][5], i_mm23[8][5][5];
  int i_int11[8][8][5][5], i_int21[8][8][5][5][5], i_int22[8][8][5][5][5][5];
  int i_TermAU, i_ninio, i_MAX_NINIO, i_DuplexInit;
};

// Host: integer model -> DevModel.  Returns RP_OK or an error code.
int build_dev_model(const rp_model& m, DevModel* out);

}  // namespace rp
#endif
