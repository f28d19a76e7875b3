"""Randomised parity census on the GPU (dev tool; the oracle is only the checker): for several horizons and steering-rate
limits, solve fresh synthetic batches through the product path (literal and early-polish termination) and compare the
first control and the predicted horizon with the independent KKT-Newton certificate of oracle/mpc_numpy.py.
usage: python tools/parity_census.py [problems per case, default 96] -> one line per case + a JSON summary"""
import dataclasses, json, os, sys, time
from concurrent.futures import ProcessPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import oracle_params, product_params
from oracle import mpc_numpy as O

def newton(args):
    x0, ref, up, N, du = args
    u0, X, U, _ = O.solve_kkt_newton(x0, ref, up, oracle_params(N, du))
    return u0, X

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    from rrt_mpc_b200 import MPCController, SolverSettings
    from rrt_mpc_b200.synthetic import make_batch
    cases = [(5, 0.15), (10, 0.05), (15, 0.15), (20, 0.15), (20, 0.02), (30, 0.05), (50, 0.02), (50, 0.15), (64, 0.02)]
    out = []
    with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        for N, du in cases:
            x0, ref, up = make_batch(B, N, seed=1000 + N + int(1000 * du))
            t = time.time()
            cert = list(pool.map(newton, [(x0[b], ref[b], up[b], N, du) for b in range(B)], chunksize=4))
            u0c = np.array([c[0] for c in cert]); Xc = np.array([c[1] for c in cert])
            row = {"horizon": N, "du": du, "problems": B, "oracle_s": round(time.time() - t, 1)}
            for name, early in (("literal", False), ("early", True)):
                ctl = MPCController(product_params(N, du), SolverSettings(eps_abs=1e-6, eps_rel=1e-6, polish_passes=5, polish_retry=2,
                                                                            early_polish=early), max_batch=B)
                r = ctl.solve_batch(x0, ref, u_prev=up)
                Xp = r.Xp if r.Xp.shape == Xc.shape else np.transpose(r.Xp, (0, 2, 1))
                row[name] = {"solved": int((r.status == 1).sum()), "max_u0_err": float(np.abs(r.u0 - u0c).max()),
                             "max_X_err": float(np.abs(Xp - Xc).max()), "mean_iters": float(r.iters.mean())}
            print(json.dumps(row), flush=True)
            out.append(row)
    worst = max(max(r["literal"]["max_u0_err"], r["early"]["max_u0_err"]) for r in out)
    print(json.dumps({"cases": len(out), "problems": B * len(out), "worst_u0_err": worst,
                      "all_solved": all(r[m]["solved"] == r["problems"] for r in out for m in ("literal", "early"))}))
