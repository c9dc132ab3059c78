"""Shared helpers of the parity tests: scene packs, the oracle, and the comparison metrics of
BASELINE.json (hit ids bit-exact; <= 1 LSB on >= 99.9 % of pixels; PSNR >= 50 dB)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENES = ["cornellBox", "texture", "transform", "dragon_superlow", "dragon_low", "dragon_full", "outScene"]


def load_scene(name):
    from rayhs_b200 import Scene

    return Scene.from_pack(os.path.join(GOLDEN, name + ".pack"))


def oracle_for(scene):
    from oracle.orc import OracleScene

    return OracleScene(scene.raw)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def compare_images(gpu_u8, ref_u8):
    diff = np.abs(gpu_u8.astype(np.int32) - ref_u8.astype(np.int32)).max(axis=-1)
    return dict(exact=float(np.mean(diff == 0)), within1=float(np.mean(diff <= 1)), maxdiff=int(diff.max()),
                psnr=psnr(gpu_u8, ref_u8))


def assert_parity(gpu_u8, ref_u8, what=""):
    m = compare_images(gpu_u8, ref_u8)
    assert m["within1"] >= 0.999, (what, m)   # BASELINE.json: <= 1 LSB per channel on >= 99.9 % of pixels
    assert m["psnr"] >= 50.0, (what, m)       # BASELINE.json: PSNR >= 50 dB
    return m
