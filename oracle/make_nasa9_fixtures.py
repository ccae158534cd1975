#!/usr/bin/env python
"""TEST INFRASTRUCTURE - writes tests/golden/<cfg>_nasa9.npz: the NASA-9 polynomial coefficients (thermo/NASA9/<species>.txt, 2 x 10
numbers: 200-1000 K and 1000-6000 K, thermo/gibbs_text.txt:6-11) of every species of the BASELINE configs' networks, so that the
reverse-rate (equilibrium-constant) path can be tested on the GPU box, which has no /root/reference.  Data files only, no code.
`CH3CCH` (SNCHO_photo_network_2025) has no file of its own: it is propyne, `CH3C2H` (SURVEY.md §8c shim 7)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("VULCAN_REFERENCE", "/root/reference")
ALIAS = {"CH3CCH": "CH3C2H"}
for tag in ("HD189", "Jupiter", "Earth", "HD209S", "HD189ion", "EarthS"):
    with open(os.path.join(REPO, "tests", "golden", tag + "_network.json")) as f:
        species = json.load(f)["species"]
    coef = np.zeros((len(species), 20))
    for i, sp in enumerate(species):
        path = os.path.join(REF, "thermo", "NASA9", ALIAS.get(sp, sp) + ".txt")
        coef[i] = np.loadtxt(path).flatten()[:20]
    out = os.path.join(REPO, "tests", "golden", tag + "_nasa9.npz")
    np.savez_compressed(out, species=np.array(species), coef=coef)
    print(tag, coef.shape, out)
