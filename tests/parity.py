"""Tolerance semantics of the ydot parity check:
    |ydot_gpu - ydot_ref| <= 1e-12 * max(|ydot_ref|, sum of |terms| of that balance equation)
CUDA's pow/cbrt/cos differ from glibc's in the last bit or two, and a ydot component is a
difference of fluxes, so the comparison scale is the size of the addends, not of the result.
The un-cancelled flux arrays are compared at 1e-12 relative on their own."""
import numpy as np

RTOL = 1e-12


def ydot_scale(snap, o):
    """per-component sum of |terms| from oracle outputs `o` (arrays in reference order)."""
    Ne, Nr, Nl = int(snap["Ne"][0]), int(snap["Nr"][0]), int(snap["Nl"][0])
    area, sy = snap["ele_area"], snap["ele_Sy"]
    qs = np.abs(o["QeleSurf"]).reshape(3, Ne).sum(0) + np.abs(o["Qe2r_Surf"])
    qg = np.abs(o["QeleSub"]).reshape(3, Ne).sum(0) + np.abs(o["Qe2r_Sub"])
    qss = np.abs(snap["ele_QSS"]) / area
    qbc = np.abs(snap.get("ele_QBC", np.zeros(Ne))) / area
    s_sf = np.abs(snap["qEleNetPrep"]) + o["qEleInfil"] + o["qEleExfil"] + qs / area + o["qEs"] + qss
    s_us = (o["qEleInfil"] + np.abs(o["qEleRecharge"]) + o["qEu"] + o["qTu"]) / sy
    s_gw = (np.abs(o["qEleRecharge"]) + o["qEleExfil"] + qg / area + o["qEg"] + o["qTg"] + qbc + qss) / sy
    y = snap["y"]
    yr = y[3 * Ne:3 * Ne + Nr]
    s, w0, L = np.abs(snap["riv_bankslope"]), snap["riv_BottomWidth"], snap["riv_Length"]
    w = np.maximum(2 * yr * snap["riv_bankslope"] + w0, 0)
    SA = (np.abs(o["QrivUp"]) + np.abs(o["QrivSurf"]) + np.abs(o["QrivSub"]) + np.abs(o["QrivDown"])
          + np.abs(snap.get("riv_qBC", np.zeros(Nr)))) / L
    with np.errstate(divide="ignore", invalid="ignore"):
        s_riv = np.where(s < 0.05e-6, SA / np.maximum(w, 1e-300), (w + np.sqrt(w * w + 4 * s * SA)) / (2 * s))
    s_lake = np.zeros(Nl)
    if Nl:
        ln = snap["ele_lakenabr"].reshape(3, Ne)
        fus = np.maximum(snap["fu_Sub"], 1e-300)
        for l in range(Nl):
            msk = ln == (l + 1)
            bs = np.abs(o["QeleSurf"].reshape(3, Ne)[msk]).sum()
            bg = (np.abs(o["QeleSub"].reshape(3, Ne)) / fus[None, :])[msk].sum()
            s_lake[l] = (abs(o["qLakePrcp"][l]) + abs(o["qLakeEvap"][l])
                         + (abs(o["QLakeRivIn"][l]) + abs(o["QLakeRivOut"][l]) + bs + bg) / o["y2LakeArea"][l])
    return np.concatenate([s_sf, s_us, s_gw, s_riv, s_lake])


def mismatches(got, ref, scale=None, rtol=RTOL):
    """indices where |got-ref| exceeds rtol * max(|ref|, scale)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    sc = np.abs(ref) if scale is None else np.maximum(np.abs(ref), scale)
    bad = ~(np.abs(got - ref) <= rtol * sc)
    bad &= ~((got == ref) | (np.isnan(got) & np.isnan(ref)))
    return np.nonzero(bad)[0]
