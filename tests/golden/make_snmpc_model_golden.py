#!/usr/bin/env python
"""Numeric pin of the SNMPC model: evaluates the reference's OWN folded model expressions, as committed in
/root/reference/acados_ocp_SNMPC.json (`model.disc_dyn_expr` json line 775, `model.con_h_expr` :745,
`model.cost_y_expr` :772, `model.cost_y_expr_e`, `model.con_h_expr_e`), on seeded random augmented states and
writes inputs + outputs to tests/golden/snmpc_model_golden.npz.

Runs ONLY in the build container (needs /root/reference).  The JSON text is parsed by casadi_print_eval.py
(CasADi is not installed and nothing is imported from the reference).  Two evaluations are stored per case:

  *_print : every numeric literal exactly as printed (6 significant digits)   -> compare at ~1e-5
  *_exact : each printed literal that is the rounding of a model constant is replaced by the full-precision
            constant from the reference's committed config (Config/EDGAR/veh_params_pred.yaml:3-10,
            pacejka_params.yaml:3-12, pred_model_dynamic_disc.py:38-46, 2 pi), AFTER asserting that the printed
            literal equals that constant rounded to 6 significant digits    -> compare at ~1e-12

The g-g-v look-up tables are not part of the JSON (CasADi prints only the interpolant's name); they are read
from the reference's Config/EDGAR/ggv.csv (SNMPC_class.py:322-335 builds them with casadi.interpolant 'linear').
"""
import json
import math
import os
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import casadi_print_eval as cpe  # noqa: E402

REF = "/root/reference"
NS, L = 10, 10
STATE = ["posx", "posy", "yaw", "vlong", "vlat", "yawrate", "delta_f", "a"]


def exact_constants():
    with open(f"{REF}/Config/EDGAR/veh_params_pred.yaml") as f:
        veh = yaml.safe_load(f)
    with open(f"{REF}/Config/EDGAR/pacejka_params.yaml") as f:
        tp = yaml.safe_load(f)["tire_params"]
    tire = {**tp["front"], **tp["rear"]}
    lf, lr, m, Iz = veh["lf"], veh["lr"], veh["m"], veh["Iz"]
    g = 9.81                                                   # pred_model_dynamic_disc.py:38
    Fz_f, Fz_r = m * lr * g / (lf + lr), m * lf * g / (lf + lr)
    cands = {
        "Fz_f": Fz_f, "Fz_r": Fz_r,
        "Fmax_f": math.sqrt(Fz_f ** 2 + (tire["Cf"] * Fz_f) ** 2), "Fmax_r": math.sqrt(Fz_r ** 2 + (tire["Cr"] * Fz_r) ** 2),
        "Df": tire["Df"], "Dr": tire["Dr"], "1/m": 1.0 / m, "1/Iz": 1.0 / Iz,
        "kaero": 0.5 * veh["ro"] * veh["S"] * veh["Cd"], "2pi": 2.0 * math.pi,
    }
    return cands


def literal_map(texts):
    """printed literal -> full-precision constant, for every literal that is the 6-significant-digit print of a constant."""
    cands = exact_constants()
    lits = set()
    for t in texts:
        lits |= {tok[1] for tok in cpe.tokenize(t) if tok[0] == "num"}
    out, names = {}, {}
    for lit in sorted(lits):
        for name, val in cands.items():
            if float(f"{val:.6g}") == float(lit) and float(lit) != val:
                out[lit] = val
                names[lit] = name
    missing = set(cands) - set(names.values())
    assert not missing, f"constants not found in the JSON print: {missing}"
    return out, names


def build(J):
    model = J["model"]
    texts = {k: model[k] for k in ("disc_dyn_expr", "con_h_expr", "con_h_expr_e", "cost_y_expr", "cost_y_expr_e")}
    progs = {k: cpe.parse(v) for k, v in texts.items()}
    lit, names = literal_map(texts.values())
    ggv = np.genfromtxt(f"{REF}/Config/EDGAR/ggv.csv", delimiter=",", skip_header=1)

    # the CasADi Function `f` (pred_model_dynamic_disc.py:186): state = the nominal symbols, body = the RHS vertcat
    ddefs = progs["disc_dyn_expr"][0]
    body = [k for k, d in sorted(ddefs.items())
            if d[0] == "call" and d[1] == "vertcat" and len(d[2]) == 8 and d[2][-1] == ("var", "jerk")]
    assert len(body) == 1
    body = body[0]

    def f_call(ev, x, u):
        x, u = np.asarray(x, float).ravel(), np.asarray(u, float).ravel()
        env = dict(ev.env)
        env.update({f"{n}_0": x[i] for i, n in enumerate(STATE)})
        env.update({"jerk": u[0], "steering_rate": u[1]})
        return ev.child(env).ev(("ref", body))

    fns = {
        "f": f_call,
        "ax_max_interpolant": lambda ev, v: cpe.interp_linear(ggv[:, 0], ggv[:, 1], float(v)),
        "ay_max_interpolant": lambda ev, v: cpe.interp_linear(ggv[:, 0], ggv[:, 3], float(v)),
    }
    # reshape targets (the print drops them): 8 (1 + ... ) -> reshape(f_disc, 8, n_samples); p[:-2] -> (n_samples, L)
    rr = {8 * NS: 8, NS * L: NS}

    def run(key, xa, u, p, exact):
        defs, final = progs[key]
        env = {f"{n}_{i}": xa[8 * i + j] for i in range(NS + 1) for j, n in enumerate(STATE)}
        env.update({"jerk": u[0], "steering_rate": u[1], "A_pce": p[:-2].reshape(-1, 1), "risk_parameter": p[-2],
                    "stop_flag": p[-1]})
        ev = cpe.Evaluator(defs, env, fns, lit if exact else None, rr)
        return np.asarray(ev.ev(final), float).ravel()

    return run, names


def main():
    with open(f"{REF}/acados_ocp_SNMPC.json") as f:
        J = json.load(f)
    run, names = build(J)
    print("literals replaced in the *_exact evaluation:", names)
    from tum_control_b200 import pce
    from tum_control_b200.config import MPC_PARAMS
    _, A, _ = pce.samples_and_A(NS, MPC_PARAMS["stds"], 2)
    rng = np.random.default_rng(20261017)
    n = 160
    lo = np.array([-600, -600, -7.0, 4.0, -1.5, -0.35, -0.15, -3.4])
    hi = np.array([600, 600, 7.0, 36.0, 1.5, 0.35, 0.15, 2.4])
    xa = np.zeros((n, 8 * (NS + 1)))
    for c in range(n):
        x0 = rng.uniform(lo, hi)
        xa[c, :8] = x0
        spread = np.array([0.5, 0.5, 0.05, 1.5, 0.4, 0.06, 0.01, 0.4])
        for i in range(1, NS + 1):
            xa[c, 8 * i:8 * i + 8] = x0 + spread * rng.standard_normal(8)
    xa[5, 3] = 0.0005; xa[5, 8 + 3] = 0.0002          # the `0.001 < vlong` branch of the slip angles
    xa[6, 2] = -0.3; xa[7, 2] = 6.9                    # yaw wrap on both sides
    u = np.stack([rng.uniform(-6, 5, n), rng.uniform(-0.3, 0.3, n)], axis=1)
    stop = (np.arange(n) % 2).astype(float)
    out = {"xa": xa, "u": u, "stop": stop, "A": A, "gamma": np.array([0.8]), "dt": np.array([J["solver_options"]["tf"] / J["dims"]["N"]])}
    for tag, exact in (("print", False), ("exact", True)):
        disc = np.zeros((n, 8 * (NS + 1))); h = np.zeros(n); he = np.zeros(n); y = np.zeros((n, 6)); ye = np.zeros((n, 4))
        for c in range(n):
            p = np.concatenate([A.ravel(), [0.8, stop[c]]])
            disc[c] = run("disc_dyn_expr", xa[c], u[c], p, exact)
            h[c] = run("con_h_expr", xa[c], u[c], p, exact)[0]
            he[c] = run("con_h_expr_e", xa[c], u[c], p, exact)[0]
            y[c] = run("cost_y_expr", xa[c], u[c], p, exact)
            ye[c] = run("cost_y_expr_e", xa[c], u[c], p, exact)
        out.update({f"disc_{tag}": disc, f"h_{tag}": h, f"he_{tag}": he, f"y_{tag}": y, f"ye_{tag}": ye})
    d = np.abs(out["disc_print"] - out["disc_exact"]).max()
    print(f"n = {n}; max |print - exact| over disc_dyn outputs: {d:.3e}")
    np.savez_compressed(os.path.join(HERE, "snmpc_model_golden.npz"), **out)
    print("snmpc_model_golden.npz", os.path.getsize(os.path.join(HERE, "snmpc_model_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
