"""Writes tests/golden/thickness_small.npz: small volumes with every intermediate of the Thickness path as the CPU
oracle (oracle/lt_oracle.c) computes them today.  The reference itself cannot run here (Java, un-vendored jar, no
JVM), so these are NOT reference outputs: they freeze the oracle -- which tests/test_oracle_golden.py pins to the
reference's own known-answer tests -- so that neither the oracle nor the CUDA path can drift unnoticed.

    python tests/golden/make_golden.py        # regenerate (only after a deliberate oracle change)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lt_oracle as o          # noqa: E402
from tests import shapes                   # noqa: E402


def cases():
    yield "blobs_a", shapes.random_blobs((14, 17, 19), 21)
    yield "blobs_b", shapes.random_blobs((9, 24, 16), 22)
    plate = np.zeros((12, 12, 12), np.uint8); plate[3:8] = 255
    yield "plate", plate
    rod = np.zeros((16, 15, 15), np.uint8)
    z, y, x = np.indices(rod.shape); rod[(y - 7) ** 2 + (x - 7) ** 2 <= 16] = 255
    yield "rod", rod
    yield "solid", np.full((4, 5, 6), 255, np.uint8)
    yield "empty", np.zeros((4, 5, 6), np.uint8)


def main():
    out = {}
    for name, vol in cases():
        out[f"{name}/vol"] = vol
        for inverse in (False, True):
            tag = f"{name}/{'sp' if inverse else 'th'}"
            r = o.process_image(vol, inverse=inverse, mask=True, pixel_width=0.5, taps=True)
            out[f"{tag}/sq"] = r["sq"].astype(np.uint32)
            out[f"{tag}/ridge"] = np.where(r["ridge"] > 0, r["sq"], 0).astype(np.uint32)
            out[f"{tag}/rsq"] = r["rsq"].astype(np.uint32)
            out[f"{tag}/map"] = r["map"].astype(np.float32)
            st = r["stats"]
            out[f"{tag}/stats"] = np.array([st.count, st.mean, st.sd, st.min, st.max], np.float64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "thickness_small.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
