// ORACLE (test infrastructure only; never linked into or imported by the product).
// Sequential restatement of the post-solve outlier loops of the reference:
//   PyrPoseOptim  /root/reference/src/optimizer.cc:1236-1302
//   PyrBA         /root/reference/src/optimizer.cc:1616-1684
// written with the reference's own running counters (FeatNum_tmp / num_badFeat reset at the end of each object),
// so that it also shows what the reference does when the caller's bookkeeping is inconsistent.
#include <cmath>
#include <cstdint>
#include <cstddef>
#include "../include/tslam_b200.h"

extern "C" int tso_gate_residuals(const double* FinalResidual, int num_s_residual, int num_t_residual, const int32_t* vIdx2vTextsGood,
                                  const int32_t* vSizeEachObj, int n_obj, const tslam_gate_options* g, uint8_t* pt_bad, uint8_t* tf_bad,
                                  uint8_t* obj_bad, int32_t* counts) {
  int nBadS = 0, nBadFeat = 0, nBadT = 0;
  if (g->gate_points) {   // :1238-1257 / :1618-1637
    double chi2MonoUse = g->chi2_mono;
    if (g->relax_below_text_blocks > 0 && num_t_residual < g->relax_below_text_blocks) chi2MonoUse = g->chi2_mono + g->relax_amount;
    const double weight_S_x = g->w_point[0], weight_S_y = g->w_point[1];
    for (int ieval_s = 0; ieval_s < num_s_residual; ++ieval_s) {
      const double chix = (FinalResidual[ieval_s * 2] / weight_S_x) * (FinalResidual[ieval_s * 2] / weight_S_x);
      const double chiy = (FinalResidual[ieval_s * 2 + 1] / weight_S_y) * (FinalResidual[ieval_s * 2 + 1] / weight_S_y);
      pt_bad[ieval_s] = 0;
      if (chix > chi2MonoUse || chiy > chi2MonoUse) { pt_bad[ieval_s] = 1; nBadS++; }
    }
  }
  if (g->gate_text) {     // :1259-1302 / :1639-1684
    for (int o = 0; o < n_obj; ++o) obj_bad[o] = 0;
    const size_t ieva_t_begin = (size_t)num_s_residual * 2;
    int FeatNum_tmp = 0, num_badFeat = 0;
    const double weight_T = g->w_text, chi2Text = g->chi2_text;
    for (int ieval_t = 0; ieval_t < num_t_residual; ++ieval_t) {
      bool any = false;
      for (int k = 0; k < 8; ++k) {
        const double IntenErro = FinalResidual[ieva_t_begin + (size_t)ieval_t * 8 + k] / weight_T;
        if (std::abs(IntenErro) > chi2Text) any = true;
      }
      tf_bad[ieval_t] = 0;
      if (any) { tf_bad[ieval_t] = 1; num_badFeat++; nBadFeat++; }
      FeatNum_tmp++;
      const int idx_obj = vIdx2vTextsGood[ieval_t];
      if (idx_obj < 0 || idx_obj >= n_obj) return -1;
      if (FeatNum_tmp > vSizeEachObj[idx_obj]) return -2;   // the reference asserts here
      if (FeatNum_tmp == vSizeEachObj[idx_obj]) {
        const double RatioBad = (double)num_badFeat / (double)vSizeEachObj[idx_obj];
        if (RatioBad > g->text_ratio) { obj_bad[idx_obj] = 1; nBadT++; }
        FeatNum_tmp = 0; num_badFeat = 0;
      }
    }
  }
  if (counts) { counts[0] = nBadS; counts[1] = nBadFeat; counts[2] = nBadT; }
  return 0;
}
