// fpt_warp_host.h — host-side helpers of the warp-autonomous kernel shared by fpt_api.cu and tests/emu: which
// geometries the kernel serves and the fields of ScoreParams only it reads.
#pragma once
#include <cmath>

#include "fpt_internal.h"

namespace fpt {
namespace wk {

inline bool warp_geometry_ok(int hw, int shw, int ktrim, int wh_max, bool combine, bool want_win_out) {
    return combine && hw == kFastHalfWin && ((shw == 50 && ktrim == 1) || shw == 0) && wh_max <= kFastMaxScaleHalfWin &&
           !want_win_out;
}

// false when the requested window half-widths are more than three distinct values (the CTA-tiled path serves those)
inline bool warp_params_finish(ScoreParams &p, const fpt_score_args *a, int wh_max) {
    auto al = [](const void *q, unsigned m) { return (reinterpret_cast<uintptr_t>(q) & m) == 0; };
    const bool windows = a->winp_out && a->n_scales > 0;
    p.wh_max = windows ? wh_max : 0;
    p.vec_ok = al(a->exp_out, 31) && al(a->obs_out, 31) && al(a->pval_out, 31);
    p.cuts_vec = al(a->cuts_plus, 15) && al(a->cuts_minus, 15);
    p.winp_vec = 0;
    for (int h = 0; h <= kFastMaxScaleHalfWin; ++h) {
        p.h_rows[h] = 0;
        p.inv_sqrt_k[h] = 1.0 / std::sqrt((double)(2 * h + 1));
    }
    p.wmode = 0;
    p.n_win_h = 0;
    p.win_h[0] = p.win_h[1] = p.win_h[2] = -1;
    p.k_vec = 0;
    for (int k = 0; k < 3; ++k) { p.k_off[k] = 0; p.k_extra[k] = 0; }
    if (windows) {
        for (int s = 0; s < a->n_scales; ++s) {
            p.win_row_off[s] = (long long)s * (long long)a->total;
            if (al(a->winp_out + (size_t)s * (size_t)a->total, 31)) p.winp_vec |= 1u << s;
            p.h_rows[a->win_half_width[s]] |= 1u << s;
        }
        for (int h = 0; h <= kFastMaxScaleHalfWin; ++h)
            if (p.h_rows[h]) {
                if (p.n_win_h == 3) return false;
                int first = 0;
                while (!((p.h_rows[h] >> first) & 1u)) ++first;
                p.k_off[p.n_win_h] = p.win_row_off[first];
                if ((p.winp_vec >> first) & 1u) p.k_vec |= 1u << p.n_win_h;
                p.k_extra[p.n_win_h] = p.h_rows[h] & (p.h_rows[h] - 1);
                p.win_h[p.n_win_h++] = h;
            }
        p.wmode = (p.n_win_h == 1 && p.win_h[0] == 3) ? 1
                  : (p.n_win_h == 3 && p.win_h[0] == 3 && p.win_h[1] == 5 && p.win_h[2] == 7) ? 2 : 3;
    }
    return true;
}

}  // namespace wk
}  // namespace fpt
