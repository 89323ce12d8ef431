#!/usr/bin/env python
"""Generate the committed known-answer fixtures of tests/golden/.

The reference (/root/reference) stores NO golden fields for this path and cannot be executed here
(C++ on OpenFOAM), so these fixtures are evaluations of the closed forms the reference
itself ships, written down independently of the package code where the formula is short enough:

  kirsch_points.npz       Kirsch stress and displacement at fixed points; formulas re-typed here from
                          src/solids4FoamModels/functionObjects/plateHoleAnalyticalSolution/
                          plateHoleAnalyticalSolution.C:43-88 (sigma), :91-122 (D); parameters of
                          tutorials/solids/linearElasticity/plateHole/system/controlDict
                          (farFieldTractionX 1e6, holeRadius 0.5, E 200e9, nu 0.3).
  patch_test.json         the exact strain of tutorials/solids/linearElasticity/patchTest/README.md:102-113
                          (exx 2e-6, eyy 6e-6, exy 4e-6) and the Hooke stress it implies (E 200e9 nu 0.3,
                          plane strain).
  necking_bar_table.json  piece-wise linear, clamped look-ups of tutorials/solids/elastoplasticity/
                          neckingBar/constant/plasticStrainVsYieldStress.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def kirsch(points, T=1e6, a=0.5, E=200e9, nu=0.3):
    x, y = points[:, 0], points[:, 1]
    r = np.sqrt(x * x + y * y)
    th = np.arctan2(y, x)
    c2, s2 = np.cos(2 * th), np.sin(2 * th)
    srr = T / 2 * (1 - a**2 / r**2) + T / 2 * (1 + 3 * a**4 / r**4 - 4 * a**2 / r**2) * c2
    stt = T / 2 * (1 + a**2 / r**2) - T / 2 * (1 + 3 * a**4 / r**4) * c2
    srt = -T / 2 * (1 - 3 * a**4 / r**4 + 2 * a**2 / r**2) * s2
    sig = np.zeros((len(x), 6))
    for i in range(len(x)):
        R = np.array([[np.cos(th[i]), -np.sin(th[i])], [np.sin(th[i]), np.cos(th[i])]])
        S = R @ np.array([[srr[i], srt[i]], [srt[i], stt[i]]]) @ R.T
        sig[i, 0], sig[i, 1], sig[i, 3] = S[0, 0], S[0, 1], S[1, 1]
    mu = E / (2 * (1 + nu))
    kappa = 3 - 4 * nu
    D = np.zeros((len(x), 3))
    D[:, 0] = a * T / (8 * mu) * (r / a * (kappa + 1) * np.cos(th) + 2 * a / r * ((1 + kappa) * np.cos(th) + np.cos(3 * th))
                                   - 2 * a**3 / r**3 * np.cos(3 * th))
    D[:, 1] = a * T / (8 * mu) * (r / a * (kappa - 3) * np.sin(th) + 2 * a / r * ((1 - kappa) * np.sin(th) + np.sin(3 * th))
                                   - 2 * a**3 / r**3 * np.sin(3 * th))
    return sig, D


def main():
    rng = np.random.default_rng(20240601)
    r = rng.uniform(0.5, 2.0, 64)
    th = rng.uniform(0.0, np.pi / 2, 64)
    pts = np.stack([r * np.cos(th), r * np.sin(th), np.zeros(64)], axis=1)
    sig, D = kirsch(pts)
    np.savez(os.path.join(HERE, "kirsch_points.npz"), points=pts, sigma=sig, D=D)

    E, nu = 200e9, 0.3
    mu, lam = E / (2 * (1 + nu)), nu * E / ((1 + nu) * (1 - 2 * nu))
    exx, eyy, exy = 2e-6, 6e-6, 4e-6
    tr = exx + eyy
    sigma = [2 * mu * exx + lam * tr, 2 * mu * exy, 0.0, 2 * mu * eyy + lam * tr, 0.0, lam * tr]
    with open(os.path.join(HERE, "patch_test.json"), "w") as f:
        json.dump(dict(epsilon=dict(xx=exx, yy=eyy, xy=exy), sigma=sigma, E=E, nu=nu), f, indent=1)

    eps = [0.000, 0.006, 0.019, 0.038, 0.066, 0.147, 0.500, 1.000]
    sy = [0.451e9, 0.476e9, 0.525e9, 0.583e9, 0.642e9, 0.710e9, 0.777e9, 0.831e9]
    xs = [-0.5, 0.0, 0.001, 0.006, 0.01, 0.03, 0.05, 0.1, 0.2, 0.49, 0.5, 0.9, 1.0, 2.0]
    out = []
    for x in xs:
        if x <= eps[0]:
            out.append(sy[0])
        elif x >= eps[-1]:
            out.append(sy[-1])
        else:
            i = max(k for k in range(len(eps)) if eps[k] <= x)
            out.append(sy[i] + (sy[i + 1] - sy[i]) * (x - eps[i]) / (eps[i + 1] - eps[i]))
    with open(os.path.join(HERE, "necking_bar_table.json"), "w") as f:
        json.dump(dict(x=xs, sigmaY=out), f, indent=1)


if __name__ == "__main__":
    main()
