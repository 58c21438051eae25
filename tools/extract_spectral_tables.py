#!/usr/bin/env python3
"""Extract the standard spectral DATA tables the host side needs into an .npz.

The numbers are published physical/colourimetric data (CIE 1931 2-degree
observer at 1 nm, CIE D65 / D50 / ACES-D60 SPDs, CIE S0/S1/S2 daylight basis,
measured n/k of Cu/Au/Ag/Al, Sellmeier-tabulated glass IORs).  The reference
tabulates them in src/spectra/cie.rs and src/spectra/named_spectrum.rs; this
script parses only the numeric literals of those `const NAME: [Float; N]`
arrays (no code) so that scenes rendered here use bit-identical spectra to the
ones a shimmer host would pass across the ABI.

Run in the build container only (needs /root/reference):
    python tools/extract_spectral_tables.py
Output: shimmer_b200/data/spectra.npz (committed; the GPU box never reads
/root/reference).
"""
import re, sys, os
import numpy as np

REF = "/root/reference/src/spectra"
OUT = os.path.join(os.path.dirname(__file__), "..", "shimmer_b200", "data", "spectra.npz")

def arrays(path):
    src = open(path).read()
    out = {}
    for m in re.finditer(r"const\s+([A-Z0-9_]+)\s*:\s*\[Float;\s*([A-Z0-9_]+)\]\s*=\s*\[(.*?)\];", src, re.S):
        name, _n, body = m.group(1), m.group(2), m.group(3)
        body = re.sub(r"//.*", "", body)
        vals = [float(t) for t in re.findall(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", body)]
        out[name] = np.asarray(vals, dtype=np.float32)   # Float = f32 in the reference
    return out

def main():
    t = {}
    t.update(arrays(os.path.join(REF, "cie.rs")))
    t.update(arrays(os.path.join(REF, "named_spectrum.rs")))
    # the EWA filter's 128-entry Gaussian LUT (mipmap.rs:388-518; pbrt's MIPFilterLUT values)
    src = open(os.path.join(REF, "..", "mipmap.rs")).read()
    m = re.search(r"const MIP_FILTER_LUT: \[Float; MIP_FILTER_LUT_SIZE\] = \[(.*?)\];", src, re.S)
    t["MIP_FILTER_LUT"] = np.asarray([float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", m.group(1))], dtype=np.float32)
    assert t["MIP_FILTER_LUT"].shape == (128,)
    for k, v in sorted(t.items()):
        print(f"{k:28s} {v.shape}")
    assert t["CIE_X"].shape == (471,) and t["CIE_LAMBDA"][0] == 360 and t["CIE_LAMBDA"][-1] == 830
    np.savez_compressed(OUT, **t)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")

if __name__ == "__main__":
    main()
