#!/usr/bin/env python3
"""Extract the Hosek-Wilkie spectral sky DATASET (numbers only) from the reference checkout into data/hosek_wilkie_spectral.npz.

The dataset is the published ArHosekSkyModelData_Spectral.h of Hosek & Wilkie (2012-2013, BSD 3-clause) as carried by pbrt-v4 and
by the reference (src/lights/hosek_wilkie_data.jl:6-1610): for each of the 11 bands 320..720 nm, 2 albedos x 10 turbidities x 6
Bernstein control points x 9 coefficients (configs) and 2 x 10 x 6 radiance scalars.  The solar-disc tables (:1612-3411) are not
needed: sunsky_to_envlight bakes the in-scattered sky only and adds the sun as a separate SunLight (sun_sky.jl:358-434).
Run in the build container (needs /root/reference):  python tools/extract_hosek.py
"""
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "hosek_wilkie_spectral.npz")
src = open(os.path.join(REF, "src/lights/hosek_wilkie_data.jl")).read()


def array(name):
    m = re.search(r"const %s = Float64\[(.*?)\]" % re.escape(name), src, re.S)
    assert m, name
    return np.array([float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", m.group(1))], dtype=np.float64)


bands = [320 + 40 * i for i in range(11)]
configs = np.stack([array("_HOSEK_SPECTRAL_CONFIG_%d" % b) for b in bands])
radiances = np.stack([array("_HOSEK_SPECTRAL_RAD_%d" % b) for b in bands])
assert configs.shape == (11, 2 * 10 * 6 * 9) and radiances.shape == (11, 2 * 10 * 6), (configs.shape, radiances.shape)
np.savez_compressed(OUT, wavelengths=np.array(bands, dtype=np.float64), configs=configs, radiances=radiances)
print("wrote", OUT, configs.shape, radiances.shape)
