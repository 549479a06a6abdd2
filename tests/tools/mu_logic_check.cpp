// Host check of the scalar mu rules shared by the host solver and the device-resident outer loop of the batch engine
// (abip_b200/csrc/lp_logic.h: lp_mu_rule, lp_update_barrier, lp_update_barrier_dynamic(_2), lp_inner_stopper).  Reads one case
// per line from stdin, prints the new state with 17 digits; tests/test_lp_logic_host.py compares with the oracle's restatement
// of src/abip.c:753-992, 2251-2277 (oracle/lp_oracle.py).
#include <cstdio>
#include "../../abip_b200/csrc/lp_logic.h"

int main() {
    double mu, sigma, gamma, dyn, eps, sp, spr, second, dynx, thresh, rp, rd, rg, minxs, sumxs;
    int fc, dc, hybrid;
    long np1;
    while (scanf("%lf %lf %lf %lf %d %d %lf %lf %lf %lf %lf %lf %d %ld %lf %lf %lf %lf %lf", &mu, &sigma, &gamma, &dyn, &fc, &dc, &eps,
                 &sp, &spr, &second, &dynx, &thresh, &hybrid, &np1, &rp, &rd, &rg, &minxs, &sumxs) == 19) {
        LpMuState st{mu, sigma, gamma, dyn, fc, dc};
        const LpMuParams p{eps, sp, spr, second, dynx, thresh, hybrid, np1};
        LpResid r{};
        r.res_pri = rp; r.res_dual = rd; r.rel_gap = rg;
        const int rule = lp_mu_rule(&st, p);
        int rc = 0;
        if (rule == 1) lp_update_barrier(&st, p, r);
        else if (rule == 2) lp_update_barrier_dynamic_2(&st, p);
        else if (rule == 3) rc = lp_update_barrier_dynamic(&st, p, minxs, sumxs);
        const double spmin = sp < spr ? sp : spr;
        printf("%d %d %.17g %.17g %.17g %.17g %d %d %ld\n", rule, rc, st.mu, st.sigma, st.gamma, st.dynamic_sigma, st.final_check,
               st.double_check, lp_inner_stopper(spmin, st.mu, 1000000L));
    }
    return 0;
}
