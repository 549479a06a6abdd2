import sys, numpy as np
sys.path.insert(0, '/root/repo')
from abip_b200 import problems
from abip_b200.api import LpEngine, SC
from oracle import lp_oracle as O
p = problems.cfg1()
st = O.Settings(eps=1e-4)
w = O.Work(p.csc(), p.b, p.c, st)
A = w.A.tocsc(); A.sort_indices()
e = LpEngine(A); e.set_problem(w.b, w.c, w.D, w.E)
e.cold_start(1.0, 1.0); e.outer_prologue(0)
# one inner iteration
w.u_prev[:] = w.u
O.project_lin_sys(w, 0); O.project_barrier(w); O.update_dual_vars(w); O.compute_avg(w, 0)
sc = e.admm_iter(0, 0, 1.0, 1.0)
print('u diff', np.abs(e.get('U')-w.u).max(), 'v diff', np.abs(e.get('V')-w.v).max())
O.update_barrier_dynamic_2(w); print('mu', w.mu)
for idx in (0,1):
    O.reinitialize_vars(w, idx); e.reinit(idx, w.sigma, 0)
print('u diff', np.abs(e.get('U')-w.u).max(), 'v diff', np.abs(e.get('V')-w.v).max())
# oracle BB with full per-round dots
m, l, a = w.m, w.l, st.alpha
u_prev, v_prev = w.u.copy(), w.v.copy()
beta_prev = 1.0
v = np.zeros(l); v_next = np.zeros(l)
e.bb_begin()
g_beta_prev, carry = 1.0, 0
for rnd in range(20):
    ut, u, t = O.bb_half_step(w, u_prev, v_prev, beta_prev, 1)
    v[m:] = v_prev[m:] + (u[m:] - t - v_prev[m:])
    ut_next, u_next, t2 = O.bb_half_step(w, u, v, beta_prev, 1)
    v_next[m:] = v[m:] + (u_next[m:] - t2 - v[m:])
    dut = 2.0*v + u_next - u - v_next - v_prev; du = u - u_next; dv = (u_next-u)*(a-1.0) + v_next - v
    od = [dut@dut, dut@dv, du@du, dv@dv, du@dv]
    beta = O.bb_beta_from_scalars(st, *od, beta_prev)
    sc = e.bb_round(carry, 1, w.mu, g_beta_prev)
    gd = [sc[SC['BB_UTUT']+q] for q in range(5)]
    g_beta = O.bb_beta_from_scalars(st, *gd, g_beta_prev)
    print(rnd, 'oracle dots', ['%.6e'%x for x in od], 'beta %.6f'%beta)
    print(rnd, 'gpu    dots', ['%.6e'%x for x in gd], 'beta %.6f'%g_beta, 'cg', sc[0], sc[1])
    print('   ut diff', np.abs(e.get('BB_UT')-ut).max(), 'u', np.abs(e.get('BB_U')-u).max(), 'v', np.abs(e.get('BB_V')-v).max(), 'utn', np.abs(e.get('BB_UTNEXT')-ut_next).max(), 'un', np.abs(e.get('BB_UNEXT')-u_next).max(), 'vn', np.abs(e.get('BB_VNEXT')-v_next).max())
    d = abs(beta-beta_prev)
    if 0 < d <= st.eps_pen: break
    elif d > st.eps_pen:
        beta_prev = beta; u_prev = u.copy(); v_prev = v_prev.copy(); v_prev[:m]=v[:m]; v_prev[m:] = (w.mu/beta_prev)/u_prev[m:]
    else:
        u_prev = u.copy(); v_prev = v.copy()
    d = abs(g_beta-g_beta_prev)
    if d > st.eps_pen: g_beta_prev, carry = g_beta, 1
    else: carry = 2
